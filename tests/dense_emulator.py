"""fp32 CPU emulation of the engine's *dilated dense* evaluation plan (TEST INFRASTRUCTURE ONLY).

The reference evaluates the CNN once per site on a 401-wide window (src/app/hifimeth/mod_batch.cpp:66-75).  Neighbouring
sites' windows overlap almost entirely, so the engine evaluates every conv layer ONCE per strand position as a dilated
convolution (dilation 2^(l-1) for layer l) and reads each site's answer from row s = o - 201 of the final map.  The
only site-specific values are the first and last output of every layer (they see the zero padding instead of the
neighbouring data); they depend on the position alone, so they are dense maps too:

    Y_l[i]  = relu(b_l + sum_j W_l[j] . Y_{l-1}[i + j * 2^(l-1)])                    l = 1..6 (Y_0 = X, features)
    F_l[s]  = first output of layer l of the site with window start s + 1           (tap 0 sees the left pad)
    G_l[s]  = last output of layer l of that site                                    (last tap may see the right pad)
    T7_v[s], T8_w[s], fc1, fc2: the site-level tail, also evaluated per row s.

Site-level output v of layer l is  Y_l[s + off_l + v * 2^l]  with off_l = -(2^l - 2), except v = 0 (F_l) and
v = Lout_l - 1 (G_l).  bn0 is folded into conv1 (the pad columns that bn0 never touches are corrected in F_1 / G_1).

This file builds that plan from the ONNX weights and runs it with numpy so that tests can check (a) the plan against
the per-site oracle (oracle/cnn_oracle.py) and (b) the CUDA kernels against the plan, layer by layer.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

HALO_L = 208  # rows before strand position 0 (>= 201, multiple of 16)
HALO_R = 200  # rows after strand position L-1 that must hold zero features


@dataclass
class Term:
    src: str      # input map name
    shift: int    # row shift: out[r] += in[r + shift] @ w
    w: np.ndarray  # [Cin, Cout] float32


@dataclass
class Op:
    out: str
    cout: int
    bias: np.ndarray
    terms: list = field(default_factory=list)
    relu: bool = True


def conv_lengths(k1: int):
    lens = [401]
    for l in range(8):
        k = k1 if l == 0 else 3
        lens.append((lens[-1] + 2 - k) // 2 + 1)
    return lens  # lens[l] = site-level length of layer l's output (lens[0] = 401)


def build_plan(w):
    """w: oracle.cnn_oracle.CnnWeights -> list[Op] (topological order).  Final op writes 'LOGIT' [rows, 2]."""
    g, b, m, v = (np.asarray(t, np.float64) for t in w.bn0)
    scale = g / np.sqrt(v + w.bn_eps)
    shift = b - m * scale
    convs = [(np.asarray(cw, np.float64), np.asarray(cb, np.float64)) for cw, cb in w.convs]
    k1 = convs[0][0].shape[2]
    lens = conv_lengths(k1)
    ops = []
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)

    # ---- layer 1 on the raw features, bn0 folded -------------------------------------------------------------
    W1, b1 = convs[0]
    W1s = W1 * scale[None, :, None]
    b1f = b1 + (W1 * shift[None, :, None]).sum(axis=(1, 2))
    ops.append(Op("Y1", 128, f32(b1f), [Term("X", j, f32(W1s[:, :, j].T)) for j in range(k1)]))
    # F1: tap 0 is the left pad column: contributes exactly 0 (bn0 is applied before padding)
    ops.append(Op("F1", 128, f32(b1f - (W1[:, :, 0] * shift[None, :]).sum(1)),
                  [Term("X", j, f32(W1s[:, :, j].T)) for j in range(1, k1)]))
    # G1: last output t = lens[1]-1 starts at row s + 2t; its last tap is the right pad column
    t_last = lens[1] - 1
    assert 2 * t_last + k1 - 1 == 402, "conv1 geometry: the last tap of the last output must be the pad column"
    ops.append(Op("G1", 128, f32(b1f - (W1[:, :, k1 - 1] * shift[None, :]).sum(1)),
                  [Term("X", 2 * t_last + j, f32(W1s[:, :, j].T)) for j in range(k1 - 1)]))

    def off(l):
        return -((1 << l) - 2)

    def src(l, q):
        """Site-level element q of layer l's output -> (map, shift) relative to row s; None = zero pad."""
        n = lens[l]
        if q < 0 or q >= n:
            return None
        if l <= 6:
            if q == 0:
                return (f"F{l}", 0)
            if q == n - 1:
                return (f"G{l}", 0)
            return (f"Y{l}", off(l) + q * (1 << l))
        return (f"T{l}_{q}", 0)

    for l in range(2, 9):
        W, bb = convs[l - 1]
        cout = W.shape[0]
        taps = [f32(W[:, :, j].T) for j in range(3)]
        if l <= 6:
            d = 1 << (l - 1)
            ops.append(Op(f"Y{l}", cout, f32(bb), [Term(f"Y{l-1}", j * d, taps[j]) for j in range(3)]))
            outs = [(f"F{l}", 0), (f"G{l}", lens[l] - 1)]
        else:
            outs = [(f"T{l}_{v}", v) for v in range(lens[l])]
        for name, vv in outs:
            terms = []
            for j in range(3):
                s_ = src(l - 1, 2 * vv - 1 + j)
                if s_ is not None:
                    terms.append(Term(s_[0], s_[1], taps[j]))
            ops.append(Op(name, cout, f32(bb), terms))
    # ---- head: flatten is channel-major (index c*2 + t) ------------------------------------------------------
    assert lens[8] == 2
    fw1, fb1 = np.asarray(w.fcs[0][0], np.float64), np.asarray(w.fcs[0][1], np.float64)
    fw2, fb2 = np.asarray(w.fcs[1][0], np.float64), np.asarray(w.fcs[1][1], np.float64)
    ops.append(Op("H", fw1.shape[0], f32(fb1), [Term("T8_0", 0, f32(fw1[:, 0::2].T)), Term("T8_1", 0, f32(fw1[:, 1::2].T))]))
    ops.append(Op("LOGIT", 2, f32(fb2), [Term("H", 0, f32(fw2.T))], relu=False))
    return ops


def run_plan(ops, X: np.ndarray, stop_at: str = None):
    """X [rows, 8] f32 (strand features placed at rows HALO_L .. HALO_L+L-1, zeros elsewhere).
    Rows past the end of an input are read as zeros (they only feed rows nobody needs)."""
    maps = {"X": np.asarray(X, np.float32)}
    rows = X.shape[0]
    for op in ops:
        acc = np.tile(op.bias.astype(np.float32), (rows, 1))
        for t in op.terms:
            a = maps[t.src]
            sh = np.zeros_like(a)
            if t.shift >= 0:
                sh[:rows - t.shift] = a[t.shift:]
            else:
                sh[-t.shift:] = a[:rows + t.shift]
            acc += sh @ t.w
        maps[op.out] = np.maximum(acc, 0) if op.relu else acc
        if stop_at == op.out:
            break
    return maps


def strand_features(fwd_codes, fi, fp, ri, rp, strand: int):
    """[L, 8] f32 features of one strand in strand coordinates (eval_kmer_features.cpp:42-64)."""
    from hifimeth_b200.synth import codev1_decode_table

    lut = (codev1_decode_table().astype(np.float32) / np.float32(952)).astype(np.float32)
    L = len(fwd_codes)
    X = np.zeros((L, 8), np.float32)
    if strand == 0:
        seq = np.asarray(fwd_codes)
        own_i, own_p, opp_i, opp_p = fi, fp, ri[::-1], rp[::-1]
    else:
        seq = 3 - np.asarray(fwd_codes)[::-1]
        own_i, own_p, opp_i, opp_p = ri, rp, fi[::-1], fp[::-1]
    ok = (seq >= 0) & (seq < 4)
    X[np.nonzero(ok)[0], seq[ok]] = 1.0
    X[:, 4], X[:, 5], X[:, 6], X[:, 7] = lut[own_i], lut[own_p], lut[opp_i], lut[opp_p]
    return X


def place(Xs: np.ndarray):
    """Strand features [L, 8] -> track rows [HALO_L + L + HALO_R rounded up to 128 (+ slack), 8]."""
    L = Xs.shape[0]
    rows = ((HALO_L + L + HALO_R + 127) // 128) * 128
    X = np.zeros((rows, 8), np.float32)
    X[HALO_L:HALO_L + L] = Xs
    return X


def site_row(o: int) -> int:
    """Track row that holds the answer of the site whose centre is strand offset o."""
    return HALO_L + o - 201
