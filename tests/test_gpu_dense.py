"""GPU unit tests of dense_gemm_kernel (the one tcgen05 kernel of the CNN path) through hm_debug_dense_op, and of the
whole dense plan against its numpy restatement (tests/dense_emulator.py), which test_dense_plan.py ties to the oracle.

Arithmetic under test: bf16 hi/lo split operands, three tensor-core products per term, fp32 accumulation.  Expected
error vs exact arithmetic is ~2^-16 relative per product; tolerance below is 2e-4 of the row's magnitude."""
import numpy as np
import pytest

from hifimeth_b200 import engine as hme

pytestmark = pytest.mark.gpu


def _ref(srcs, terms, bias, rows, relu=True):
    acc = np.tile(np.asarray(bias, np.float64), (rows, 1))
    for si, sh, w in terms:
        acc += srcs[si][sh:sh + rows].astype(np.float64) @ np.asarray(w, np.float64)
    return np.maximum(acc, 0) if relu else acc


def _check(got, want, scale_rows):
    err = np.abs(got - want) / (scale_rows[:, None] + 1e-6)
    assert err.max() < 2e-4, (err.max(), np.unravel_index(err.argmax(), err.shape))


def _case(rng, rows, cin, cout, shifts, n_src=1, src_of=None):
    rows_alloc = rows + max(shifts) + 8
    srcs = [rng.standard_normal((rows_alloc, cin)).astype(np.float32) * rng.uniform(0.1, 3.0) for _ in range(n_src)]
    terms = [(src_of[k] if src_of else 0, sh, (rng.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)) for k, sh in enumerate(shifts)]
    bias = rng.standard_normal(cout).astype(np.float32)
    want = _ref(srcs, terms, bias, rows)
    scale = sum(np.abs(srcs[si][sh:sh + rows]).astype(np.float64) @ np.abs(w).astype(np.float64) for si, sh, w in terms).max(axis=1)
    return srcs, terms, bias, want, scale


@pytest.mark.parametrize("rows,cin,cout,shifts", [
    (128, 16, 64, [0]),             # one tile, one k-step
    (128, 128, 128, [0]),           # full K loop, ring wraps once
    (128 * 3, 128, 128, [0, 2, 4]),  # conv2 shape: three taps as shifted views of one staged segment
    (128 * 5, 128, 96, [0, 8, 16]),  # conv4 shape
    (128 * 2, 96, 96, [0, 32, 64]),  # conv6 shape
    (128 * 2, 96, 64, [2, 66, 130]),  # conv7 middle outputs: wide segment
    (128 * 400, 128, 128, [0, 4, 8]),  # many tiles per CTA: TMEM double buffering and ring phases
])
def test_dense_op_shapes(lib_built, rows, cin, cout, shifts):
    rng = np.random.default_rng(rows + cin + cout)
    srcs, terms, bias, want, scale = _case(rng, rows, cin, cout, shifts)
    got = hme.debug_dense_op(srcs, terms, bias, rows)
    _check(got, want, scale)


def test_dense_op_two_sources(lib_built):
    rng = np.random.default_rng(7)
    srcs, terms, bias, want, scale = _case(rng, 256, 128, 128, [0, 2], n_src=2, src_of=[0, 1])
    _check(hme.debug_dense_op(srcs, terms, bias, 256), want, scale)
    srcs, terms, bias, want, scale = _case(rng, 256, 128, 96, [370, 378, 0], n_src=2, src_of=[0, 0, 1])  # G4 of the k=11 plan
    _check(hme.debug_dense_op(srcs, terms, bias, 256), want, scale)


@pytest.mark.parametrize("taps,shift", [(11, 0), (13, 0), (11, 392), (13, 390)])
def test_dense_op_conv1_form(lib_built, taps, shift):
    rng = np.random.default_rng(taps + shift)
    rows = 384
    rows_alloc = rows + shift + 32
    x = rng.random((rows_alloc, 8)).astype(np.float32)
    w = (rng.standard_normal((taps, 8, 128)) / 9).astype(np.float32)
    bias = rng.standard_normal(128).astype(np.float32)
    acc = np.tile(bias.astype(np.float64), (rows, 1))
    scale = np.zeros(rows)
    for j in range(taps):
        acc += x[shift + j:shift + j + rows].astype(np.float64) @ w[j].astype(np.float64)
        scale = np.maximum(scale, (np.abs(x[shift + j:shift + j + rows]).astype(np.float64) @ np.abs(w[j]).astype(np.float64)).max(axis=1))
    got = hme.debug_dense_op([x], [(0, shift, w)], bias, rows, conv1_taps=taps)
    _check(got, np.maximum(acc, 0), scale * taps)


def test_dense_op_head_form(lib_built):
    rng = np.random.default_rng(11)
    rows = 256
    srcs = [np.abs(rng.standard_normal((rows, 64))).astype(np.float32) for _ in range(2)]
    terms = [(k, 0, (rng.standard_normal((64, 256)) / 8).astype(np.float32)) for k in range(2)]
    bias = rng.standard_normal(256).astype(np.float32)
    w2 = (rng.standard_normal((2, 256)) / 16).astype(np.float32)
    b2 = rng.standard_normal(2).astype(np.float32)
    h = _ref(srcs, terms, bias, rows)
    want = h @ w2.astype(np.float64).T + b2
    got = hme.debug_dense_op(srcs, terms, bias, rows, w2=w2, b2=b2)
    assert np.abs(got - want).max() < 2e-4 * np.abs(h).sum(axis=1).max()


# ---- compact ops: rows picked through a site-row index (cp.async gather producer) ----------------------------------------

def _gather_ref(srcs, terms, gmask, grows, bias, rows):
    acc = np.tile(np.asarray(bias, np.float64), (rows, 1))
    scale = np.zeros((rows, len(bias)))
    for k, (si, sh, w) in enumerate(terms):
        a = srcs[si][grows + sh] if (gmask >> k) & 1 else srcs[si][sh:sh + rows]
        acc += a.astype(np.float64) @ np.asarray(w, np.float64)
        scale += np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64)
    return np.maximum(acc, 0), scale.max(axis=1)


@pytest.mark.parametrize("cin,cout,shifts,gmask,n_src", [
    (96, 64, [2, 66, 130], 0b111, 1),   # T7_1: three gathered views of Y6
    (128, 128, [0, 2], 0b10, 2),        # F2: compact F1 (direct) + Y1 gathered at +2
    (128, 96, [370, 378, 0], 0b011, 2),  # G4 (k = 11): two gathered Y3 rows + compact G3
    (64, 64, [0, 0, 0], 0, 3),          # T8_1: all direct
])
def test_compact_op_gather(lib_built, cin, cout, shifts, gmask, n_src):
    rng = np.random.default_rng(cin * 7 + cout + gmask)
    rows, rows_alloc = 128 * 9, 6000
    srcs = [rng.standard_normal((rows_alloc, cin)).astype(np.float32) for _ in range(n_src)]
    src_of = [0, 0, 1] if (n_src == 2 and len(shifts) == 3) else ([0, 1] if n_src == 2 else ([0, 1, 2] if n_src == 3 else [0, 0, 0]))
    if n_src == 2 and len(shifts) == 2:
        src_of = [1, 0]  # direct term reads the compact map (src 1), gathered term the dense map (src 0)
    terms = [(src_of[k], sh, (rng.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)) for k, sh in enumerate(shifts)]
    bias = rng.standard_normal(cout).astype(np.float32)
    grows = np.sort(rng.integers(0, rows_alloc - 400, size=rows)).astype(np.uint32)
    want, scale = _gather_ref(srcs, terms, gmask, grows, bias, rows)
    got = hme.debug_dense_op(srcs, terms, bias, rows, gather_rows=grows, gather_mask=gmask)
    _check(got, want, scale)


@pytest.mark.parametrize("taps,shift", [(11, 0), (11, 392), (13, 390)])
def test_compact_op_conv1_gather(lib_built, taps, shift):
    rng = np.random.default_rng(taps * 3 + shift)
    rows, rows_alloc = 128 * 5, 4000
    x = rng.random((rows_alloc, 8)).astype(np.float32)
    w = (rng.standard_normal((taps, 8, 128)) / 9).astype(np.float32)
    bias = rng.standard_normal(128).astype(np.float32)
    grows = rng.integers(0, rows_alloc - shift - 16, size=rows).astype(np.uint32)
    acc = np.tile(bias.astype(np.float64), (rows, 1))
    scale = np.zeros(rows)
    for j in range(taps):
        a = x[grows + shift + j]
        acc += a.astype(np.float64) @ w[j].astype(np.float64)
        scale += (np.abs(a).astype(np.float64) @ np.abs(w[j]).astype(np.float64)).max(axis=1)
    got = hme.debug_dense_op([x], [(0, shift, w)], bias, rows, conv1_taps=taps, gather_rows=grows, gather_mask=1)
    _check(got, np.maximum(acc, 0), scale)


# ---- CTA-pair (cta_group::2) form: M = 256 per MMA, half of every weight tile per CTA --------------------------------------

@pytest.mark.parametrize("rows,cin,cout,shifts", [
    (256, 128, 128, [0]),
    (128 * 3, 128, 128, [0, 2, 4]),     # odd tile count: the peer's last tile lies in the slack rows
    (128 * 8, 128, 96, [0, 8, 16]),
    (128 * 600, 128, 128, [0, 4, 8]),   # several tile pairs per cluster: ring phases, TMEM double buffering, relay
])
def test_dense_op_cta_pair(lib_built, rows, cin, cout, shifts, monkeypatch):
    monkeypatch.setenv("HM_DENSE_2CTA", "1")
    rng = np.random.default_rng(rows + cout)
    rows_alloc = rows + max(shifts) + 128 + 512
    srcs = [rng.standard_normal((rows_alloc, cin)).astype(np.float32)]
    terms = [(0, sh, (rng.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)) for sh in shifts]
    bias = rng.standard_normal(cout).astype(np.float32)
    want = _ref(srcs, terms, bias, rows)
    scale = sum(np.abs(srcs[0][sh:sh + rows]).astype(np.float64) @ np.abs(w).astype(np.float64) for _, sh, w in terms).max(axis=1)
    got = hme.debug_dense_op(srcs, terms, bias, rows)
    _check(got, want, scale)
