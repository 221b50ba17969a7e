"""oracle/hmoracle.py -- ctypes front-end to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * ``O``   = oracle/libhm_oracle.so          our plain-C restatement (oracle/hm_oracle.c)
  * ``REF`` = oracle/_ref/libhifimeth_ref.so  the reference's own sources compiled here (oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


def build(quiet: bool = True) -> None:
    """Compile the restatement (always) and, when /root/reference exists, the reference library."""
    subprocess.run(["make", "-C", str(HERE)], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class _Oracle:
    def __init__(self):
        path = HERE / "libhm_oracle.so"
        if not path.exists():
            build()
        self.lib = L = C.CDLL(str(path))
        L.hmo_codev1_decode.restype = C.c_int
        L.hmo_codev1_encode.restype = C.c_int
        L.hmo_decode_plane.argtypes = [_u8p, C.c_int64, C.POINTER(C.c_uint16)]
        L.hmo_decode_seq.argtypes = [_u8p, C.c_int, C.c_int, _u8p, _u8p]
        L.hmo_scan_sites.argtypes = [_u8p, C.c_int, C.c_int, _i32p]
        L.hmo_site_features.argtypes = [_u8p, _u8p, C.c_int, _u8p, _u8p, _u8p, _u8p, C.c_int, _f32p]
        L.hmo_logits_to_prob.argtypes = [C.c_float, C.c_float]
        L.hmo_logits_to_prob.restype = C.c_float
        L.hmo_prob_to_ml.argtypes = [C.c_float]
        L.hmo_read_calls.argtypes = [_u8p, C.c_int, C.c_int, _i32p, _u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hmo_build_mm.argtypes = [_u8p, _i32p, C.c_int, _i32p, C.c_int, C.c_char_p]
        L.hmo_build_mod_record.argtypes = [_u8p, C.c_size_t, C.c_int, _i32p, _u8p, C.c_int, _i32p, _u8p, C.c_int, _u8p]
        L.hmo_build_mod_record.restype = C.c_size_t
        L.hmo_ml_histogram.argtypes = [_u8p, _u8p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), C.c_int64, C.POINTER(C.c_uint64)]
        L.hmo_ml_threshold.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.hmo_ml_threshold.restype = C.c_int

    def ml_histogram(self, ml, ctx, read_flag, call_read) -> np.ndarray:
        """[3, 256] u64 histograms of the ML bytes per context over reads without flag 0x900 (pileup.cpp:237-272)."""
        ml = np.ascontiguousarray(ml, np.uint8)
        ctx = np.ascontiguousarray(ctx, np.uint8)
        fl = np.ascontiguousarray(read_flag, np.uint16)
        rd = np.ascontiguousarray(call_read, np.uint32)
        bins = np.zeros((3, 256), np.uint64)
        self.lib.hmo_ml_histogram(_p(ml, _u8p), _p(ctx, _u8p), _p(fl, C.POINTER(C.c_uint16)), _p(rd, C.POINTER(C.c_uint32)), len(ml),
                                  _p(bins, C.POINTER(C.c_uint64)))
        return bins

    def ml_threshold(self, bins):
        b = np.ascontiguousarray(bins, np.uint64)
        n = C.c_uint64(0)
        return int(self.lib.hmo_ml_threshold(_p(b, C.POINTER(C.c_uint64)), C.byref(n))), int(n.value)

    # -- per-read primitives ---------------------------------------------------------------------------------
    def decode_seq(self, seq4: np.ndarray, l: int, flag: int):
        fwd = np.empty(l, np.uint8)
        rev = np.empty(l, np.uint8)
        seq4 = np.ascontiguousarray(seq4)
        ok = self.lib.hmo_decode_seq(_p(seq4, _u8p), l, int(flag), _p(fwd, _u8p), _p(rev, _u8p))
        return bool(ok), fwd, rev

    def decode_plane(self, codes: np.ndarray) -> np.ndarray:
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        out = np.empty(codes.size, np.uint16)
        self.lib.hmo_decode_plane(_p(codes, _u8p), codes.size, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def scan_sites(self, fwd_qs: np.ndarray, ctx: int) -> np.ndarray:
        out = np.empty(len(fwd_qs) + 1, np.int32)
        n = self.lib.hmo_scan_sites(_p(fwd_qs, _u8p), len(fwd_qs), ctx, _p(out, _i32p))
        return out[:n].copy()

    def site_features(self, fwd, rev, fi, fp, ri, rp, off: int):
        f = np.empty((401, 8), np.float32)
        s = self.lib.hmo_site_features(_p(fwd, _u8p), _p(rev, _u8p), len(fwd), _p(fi, _u8p), _p(fp, _u8p),
                                       _p(ri, _u8p), _p(rp, _u8p), int(off), _p(f, _f32p))
        return s, f

    def read_calls(self, fwd_qs: np.ndarray, ctx_mask: int):
        l = len(fwd_qs)
        qoff = np.empty(l + 1, np.int32)
        ctx = np.empty(l + 1, np.uint8)
        nf, nr = C.c_int(), C.c_int()
        n = self.lib.hmo_read_calls(_p(fwd_qs, _u8p), l, ctx_mask, _p(qoff, _i32p), _p(ctx, _u8p), C.byref(nf), C.byref(nr))
        return qoff[:n].copy(), ctx[:n].copy(), nf.value, nr.value

    def build_mm(self, fwd_qs, fwd_qoff, rev_qoff) -> bytes:
        fq = np.ascontiguousarray(fwd_qoff, np.int32)
        rq = np.ascontiguousarray(rev_qoff, np.int32)
        buf = C.create_string_buffer(16 + 11 * (len(fq) + len(rq)))
        n = self.lib.hmo_build_mm(_p(fwd_qs, _u8p), _p(fq, _i32p), len(fq), _p(rq, _i32p), len(rq), buf)
        return buf.raw[:n]

    def build_mod_record(self, body: bytes, keep: bool, fq, fml, rq, rml) -> bytes:
        fq = np.ascontiguousarray(fq, np.int32); rq = np.ascontiguousarray(rq, np.int32)
        fml = np.ascontiguousarray(fml, np.uint8); rml = np.ascontiguousarray(rml, np.uint8)
        src = np.frombuffer(body, np.uint8)
        out = np.empty(len(body) + 64 + 12 * (len(fq) + len(rq)), np.uint8)
        n = self.lib.hmo_build_mod_record(_p(src, _u8p), len(body), int(keep), _p(fq, _i32p), _p(fml, _u8p), len(fq),
                                          _p(rq, _i32p), _p(rml, _u8p), len(rq), _p(out, _u8p))
        return out[:n].tobytes()

    # -- batch-level pipeline (the thing the GPU engine is compared with) --------------------------------------
    def batch_sites(self, batch, ctx_mask: int = 7):
        """Per read: dict(valid, fwd, rev, qoff, ctx, n_fwd, n_rev) in the engine's output order."""
        out = []
        for r in range(batch.n_reads):
            b0, b1 = int(batch.base_off[r]), int(batch.base_off[r + 1])
            s0 = int(batch.seq_off[r])
            l = b1 - b0
            ok, fwd, rev = self.decode_seq(batch.seq4[s0:s0 + (l + 1) // 2], l, int(batch.flag[r]))
            if not batch.valid[r] or not ok:
                out.append(dict(valid=False, fwd=fwd, rev=rev, qoff=np.zeros(0, np.int32), ctx=np.zeros(0, np.uint8), n_fwd=0, n_rev=0))
                continue
            qoff, ctx, nf, nr = self.read_calls(fwd, ctx_mask)
            out.append(dict(valid=True, fwd=fwd, rev=rev, qoff=qoff, ctx=ctx, n_fwd=nf, n_rev=nr))
        return out

    def batch_features(self, batch, sites, r: int, idx=None) -> np.ndarray:
        """[n,401,8] f32 features of read r's calls (all, or the given indices into its call list)."""
        b0, b1 = int(batch.base_off[r]), int(batch.base_off[r + 1])
        s = sites[r]
        qoff = s["qoff"] if idx is None else s["qoff"][idx]
        f = np.empty((len(qoff), 401, 8), np.float32)
        fi, fp, ri, rp = (np.ascontiguousarray(a[b0:b1]) for a in (batch.fi, batch.fp, batch.ri, batch.rp))
        for k, o in enumerate(qoff):
            _, f[k] = self.site_features(s["fwd"], s["rev"], fi, fp, ri, rp, int(o))
        return f

    def batch_call(self, batch, models, ctx_mask: int = 7, chunk: int = 2048):
        """Full CPU pipeline: returns per read dict(+ logits [n,2], prob [n], ml [n])."""
        from . import cnn_oracle

        sites = self.batch_sites(batch, ctx_mask)
        for r, s in enumerate(sites):
            n = len(s["qoff"])
            s["logits"] = np.zeros((n, 2), np.float32)
            if n == 0:
                s["prob"] = np.zeros(0, np.float32); s["ml"] = np.zeros(0, np.uint8)
                continue
            for c in range(3):
                sel = np.nonzero(s["ctx"] == c)[0]
                for i in range(0, len(sel), chunk):
                    ii = sel[i:i + chunk]
                    s["logits"][ii] = cnn_oracle.forward_logits(models[c], self.batch_features(batch, sites, r, ii))
            s["prob"], s["ml"] = cnn_oracle.logits_to_prob_ml(s["logits"])
        return sites


class _Ref:
    """The reference's own compiled code.  ``available`` is False when the library was never built."""

    def __init__(self):
        path = HERE / "_ref" / "libhifimeth_ref.so"
        if not path.exists() and os.path.isdir("/root/reference/src"):
            build()
        self.available = path.exists()
        if not self.available:
            return
        self.lib = L = C.CDLL(str(path))
        L.ref_query_decode.argtypes = [_u8p, C.c_size_t, _u8p, _u8p, _i32p, _i32p, _i32p, _i32p]
        L.ref_extract_sites.argtypes = [_u8p, C.c_size_t, C.c_int, _i32p, C.c_int]
        L.ref_extract_features.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i32p, _i32p]
        L.ref_build_mod_bam.argtypes = [_u8p, C.c_size_t, C.c_int, _i32p, _u8p, C.c_int, _i32p, _u8p, C.c_int,
                                        _u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ref_parse_mods.argtypes = [_u8p, C.c_size_t, _i32p, _u8p, _u8p, C.c_int]
        self.has_pileup = hasattr(L, "hmref_pileup_thresholds")
        if self.has_pileup:
            L.hmref_pileup_thresholds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.hmref_pileup_thresholds.restype = None

    def pileup_thresholds(self, cpg, chg, chh):
        """The reference's own s_resolve_scaled_prob_threshold (src/app/hifimeth/pileup.cpp:355-436, compiled through
        oracle/ref_pileup.cpp) on three 256-bin histograms -> (cpg, chg, chh) thresholds."""
        a, b, c = (np.ascontiguousarray(x, np.uint64) for x in (cpg, chg, chh))
        out = np.zeros(3, np.uint8)
        self.lib.hmref_pileup_thresholds(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data)
        return tuple(int(x) for x in out)

    def query_decode(self, body: bytes, l: int):
        src = np.frombuffer(body, np.uint8)
        fwd = np.empty(l, np.uint8); rev = np.empty(l, np.uint8)
        k = [np.empty(l, np.int32) for _ in range(4)]
        ok = self.lib.ref_query_decode(_p(src, _u8p), len(body), _p(fwd, _u8p), _p(rev, _u8p), *[_p(a, _i32p) for a in k])
        return bool(ok), fwd, rev, k

    def extract_sites(self, body: bytes, ctx: int, l: int):
        src = np.frombuffer(body, np.uint8)
        out = np.empty(l + 1, np.int32)
        n = self.lib.ref_extract_sites(_p(src, _u8p), len(body), ctx, _p(out, _i32p), l + 1)
        return None if n < 0 else out[:n].copy()

    def extract_features(self, body: bytes, ctx: int, first: int, count: int):
        src = np.frombuffer(body, np.uint8)
        f = np.empty((count, 401, 8), np.float32)
        off = np.empty(count, np.int32); st = np.empty(count, np.int32)
        n = self.lib.ref_extract_features(_p(src, _u8p), len(body), ctx, 401, 8, first, count, _p(f, _f32p), _p(off, _i32p), _p(st, _i32p))
        if n < 0:
            return None
        return f[:n], off[:n], st[:n]

    def build_mod_bam(self, body: bytes, keep: bool, fq, fml, rq, rml) -> bytes:
        fq = np.ascontiguousarray(fq, np.int32); rq = np.ascontiguousarray(rq, np.int32)
        fml = np.ascontiguousarray(fml, np.uint8); rml = np.ascontiguousarray(rml, np.uint8)
        src = np.frombuffer(body, np.uint8)
        cap = len(body) + 64 + 12 * (len(fq) + len(rq))
        out = np.empty(cap, np.uint8)
        n = C.c_size_t()
        rc = self.lib.ref_build_mod_bam(_p(src, _u8p), len(body), int(keep), _p(fq, _i32p), _p(fml, _u8p), len(fq),
                                        _p(rq, _i32p), _p(rml, _u8p), len(rq), _p(out, _u8p), cap, C.byref(n))
        assert rc == 0
        return out[:n.value].tobytes()

    def parse_mods(self, body: bytes, cap: int):
        src = np.frombuffer(body, np.uint8)
        q = np.empty(cap, np.int32); s = np.empty(cap, np.uint8); p = np.empty(cap, np.uint8)
        n = self.lib.ref_parse_mods(_p(src, _u8p), len(body), _p(q, _i32p), _p(s, _u8p), _p(p, _u8p), cap)
        return q[:n].copy(), s[:n].copy(), p[:n].copy()


_O = None
_R = None


def oracle() -> _Oracle:
    global _O
    if _O is None:
        _O = _Oracle()
    return _O


def ref() -> _Ref:
    global _R
    if _R is None:
        _R = _Ref()
    return _R
