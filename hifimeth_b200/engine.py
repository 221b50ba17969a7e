"""ctypes binding of the C ABI (include/hm_engine.h) plus a thin host-side mirror of the reference's per-read
interface (EvalKmerFeaturesGenerator + ModBatch, src/app/hifimeth/eval_kmer_features.hpp:13-49 and
src/app/hifimeth/mod_batch.hpp:12-43) for tests and bench.py.

There is no CPU fallback: importing works anywhere (the library is cross-compiled), creating an Engine without a
B200 raises HmError.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["HM_ENGINE_LIB"]) if os.environ.get("HM_ENGINE_LIB") else PKG / "libhm_engine.so"  # override: A/B runs of two builds
DEFAULT_MODEL_DIR = PKG.parent / "models"

HM_CTX_CPG, HM_CTX_CHG, HM_CTX_CHH = 1, 2, 4
HM_CNN_TENSOR, HM_CNN_FP32_SIMT = 0, 1
HM_SUBMIT_SKIP_H2D, HM_SUBMIT_SKIP_D2H, HM_SUBMIT_MM_TEXT, HM_SUBMIT_ML_HIST, HM_SUBMIT_READ_STATS = 1, 2, 4, 8, 16

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)


class HmError(RuntimeError):
    pass


class hm_config(C.Structure):
    _fields_ = [("model_dir", C.c_char_p), ("ctx_mask", C.c_int32), ("min_read_len", C.c_int32), ("device", C.c_int32),
                ("n_slots", C.c_int32), ("max_reads", C.c_uint32), ("max_bases", C.c_uint32), ("cnn_mode", C.c_int32),
                ("keep_debug", C.c_int32)]


class hm_read_batch(C.Structure):
    _fields_ = [("max_reads", C.c_uint32), ("max_bases", C.c_uint32), ("base_off", _u32p), ("seq_off", _u32p), ("seq4", _u8p),
                ("flag", _u16p), ("valid", _u8p), ("fi", _u8p), ("fp", _u8p), ("ri", _u8p), ("rp", _u8p)]


class hm_read_stats(C.Structure):
    _fields_ = [("sum", C.c_uint64 * 4), ("max", C.c_uint32 * 4)]


class hm_call_batch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_calls", C.c_uint32), ("call_off", _u32p), ("n_fwd", _u32p), ("qoff", _i32p),
                ("ml", _u8p), ("n_sites", C.c_uint64 * 3), ("mm_text", _u8p), ("mm_off", _u32p), ("mm_fwd_len", _u32p), ("ml_hist", _u32p),
                ("read_stats", C.POINTER(hm_read_stats))]


class hm_timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("decode_ms", C.c_float), ("scan_ms", C.c_float), ("cnn_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("top_kernel_ms", C.c_float), ("top_kernel_launches", C.c_uint32),
                ("executed_flops", C.c_double)]


ABI_SYMBOLS = ["hm_engine_create", "hm_engine_destroy", "hm_last_error", "hm_version", "hm_batch_acquire", "hm_batch_submit",
               "hm_batch_collect", "hm_batch_timing", "hm_model_weights", "hm_codev1_encode", "hm_codev1_decode", "hm_pack_record", "hm_pack_records",
               "hm_mod_record_bound", "hm_build_mod_record", "hm_build_mod_record_mm", "hm_parse_mod_record", "hm_ml_threshold", "hm_call_main", "hm_call_fast_exit", "hm_bam_copy", "hm_deflate_block", "hm_inflate_block", "hm_crc32_bytes", "hm_debug_dump_decode", "hm_debug_dump_ctx",
               "hm_debug_dump_features", "hm_debug_dump_logits", "hm_debug_dump_xmap", "hm_debug_dump_acts", "hm_debug_dense_op", "hm_debug_last_op_ms", "hm_microbench"]

_lib = None


def load_library() -> C.CDLL:
    """Loads the in-tree C-ABI library; raises HmError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise HmError(f"{LIB_PATH} is missing: run `python -m hifimeth_b200.build` (or __graft_entry__.build())")
    L = C.CDLL(str(LIB_PATH))
    L.hm_engine_create.argtypes = [C.POINTER(hm_config), C.POINTER(C.c_void_p)]
    L.hm_engine_destroy.argtypes = [C.c_void_p]
    L.hm_engine_destroy.restype = None
    L.hm_last_error.argtypes = [C.c_void_p]
    L.hm_last_error.restype = C.c_char_p
    L.hm_version.restype = C.c_char_p
    L.hm_model_weights.argtypes = [C.c_char_p, _f32p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
    L.hm_batch_acquire.argtypes = [C.c_void_p, C.c_int, C.POINTER(hm_read_batch)]
    L.hm_batch_submit.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32]
    L.hm_batch_collect.argtypes = [C.c_void_p, C.c_int, C.POINTER(hm_call_batch)]
    L.hm_batch_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(hm_timing)]
    L.hm_codev1_encode.argtypes = [C.c_uint32]
    L.hm_codev1_encode.restype = C.c_uint8
    L.hm_codev1_decode.argtypes = [C.c_uint8]
    L.hm_codev1_decode.restype = C.c_uint16
    L.hm_pack_record.argtypes = [C.POINTER(hm_read_batch), _u32p, _u8p, C.c_size_t, C.c_int32]
    L.hm_mod_record_bound.argtypes = [C.c_size_t, C.c_uint32]
    L.hm_mod_record_bound.restype = C.c_size_t
    L.hm_build_mod_record.argtypes = [_u8p, C.c_size_t, C.c_int, _i32p, _u8p, C.c_uint32, _i32p, _u8p, C.c_uint32, _u8p,
                                      C.POINTER(C.c_size_t)]
    L.hm_build_mod_record_mm.argtypes = [_u8p, C.c_size_t, C.c_int, _u8p, C.c_uint32, _u8p, C.c_uint32, _u8p, C.c_uint32, C.c_uint32, _u8p,
                                         C.POINTER(C.c_size_t)]
    L.hm_ml_threshold.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.hm_ml_threshold.restype = C.c_uint8
    L.hm_parse_mod_record.argtypes = [_u8p, C.c_size_t, _i32p, _u8p, _u8p, C.c_char_p, C.c_uint32, _u32p]
    L.hm_pack_records.argtypes = [C.POINTER(hm_read_batch), C.c_uint32, C.POINTER(_u8p), C.POINTER(C.c_size_t), C.c_int32, C.c_int,
                                  _i32p, _u32p]
    L.hm_call_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.hm_bam_copy.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.hm_deflate_block.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.hm_deflate_block.restype = C.c_size_t
    L.hm_inflate_block.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.hm_inflate_block.restype = C.c_int
    L.hm_crc32_bytes.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
    L.hm_crc32_bytes.restype = C.c_uint32
    L.hm_debug_dump_decode.argtypes = [C.c_void_p, C.c_int, _u16p, _u16p, _u16p, _u16p, _u8p, _u8p]
    L.hm_debug_dump_ctx.argtypes = [C.c_void_p, C.c_int, _u8p]
    L.hm_debug_dump_features.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, _f32p]
    L.hm_debug_dump_logits.argtypes = [C.c_void_p, C.c_int, _f32p]
    L.hm_debug_dump_xmap.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, _f32p]
    L.hm_debug_dump_acts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, _f32p, C.c_size_t, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32)]
    L.hm_debug_dense_op.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(_f32p), C.c_int, _i32p, _i32p,
                                    _f32p, _f32p, C.c_int, _f32p, _f32p, _u32p, C.c_uint32, _f32p]
    L.hm_debug_last_op_ms.restype = C.c_float
    L.hm_microbench.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint32, C.c_int, _f32p, C.POINTER(C.c_double),
                                C.POINTER(C.c_double)]
    _lib = L
    return L


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


@dataclass
class CallBatch:
    n_reads: int
    n_calls: int
    call_off: np.ndarray
    n_fwd: np.ndarray
    qoff: np.ndarray
    ml: np.ndarray
    n_sites: tuple
    mm_text: np.ndarray = None     # HM_SUBMIT_MM_TEXT: device-built MM skip-count text
    mm_off: np.ndarray = None
    mm_fwd_len: np.ndarray = None
    ml_hist: np.ndarray = None     # HM_SUBMIT_ML_HIST: [3, 256] ML histograms per context (CpG, CHG, CHH)
    stats_sum: np.ndarray = None   # HM_SUBMIT_READ_STATS: [n_reads, 4] u64 sums of the decoded frames (fi, fp, ri, rp)
    stats_max: np.ndarray = None   # [n_reads, 4] u32 maxima

    def read_mm(self, r: int):
        """(fwd text, rev text) of read r: the ",d,d,..." runs that follow "C+m" and "G-m" in its MM tag."""
        a, b, nf = int(self.mm_off[r]), int(self.mm_off[r + 1]), int(self.mm_fwd_len[r])
        return self.mm_text[a:a + nf], self.mm_text[a + nf:b]

    def read_calls(self, r: int):
        """(fwd_qoff, fwd_ml, rev_qoff, rev_ml) of read r: the argument lists of build_one_mod_bam."""
        a, b = int(self.call_off[r]), int(self.call_off[r + 1])
        nf = int(self.n_fwd[r])
        return self.qoff[a:a + nf], self.ml[a:a + nf], self.qoff[a + nf:b], self.ml[a + nf:b]


class Engine:
    """One engine per GPU.  Mirrors ModModels + the per-thread ModBatch set of the reference worker."""

    def __init__(self, model_dir=None, ctx_mask: int = 7, min_read_len: int = 1000, device: int = 0, n_slots: int = 2,
                 max_reads: int = 4096, max_bases: int = 1 << 26, cnn_mode: int = HM_CNN_TENSOR, keep_debug: bool = False):
        self.lib = load_library()
        self._model_dir = str(model_dir or DEFAULT_MODEL_DIR).encode()
        cfg = hm_config(self._model_dir, ctx_mask, min_read_len, device, n_slots, max_reads, max_bases, cnn_mode, int(keep_debug))
        h = C.c_void_p()
        rc = self.lib.hm_engine_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise HmError(f"hm_engine_create failed ({rc}): {self.lib.hm_last_error(None).decode()}")
        self.h = h
        self.n_slots = n_slots
        self.min_read_len = min_read_len

    def close(self):
        if getattr(self, "h", None):
            self.lib.hm_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise HmError(f"{what} failed ({rc}): {self.lib.hm_last_error(self.h).decode()}")

    # -- staging ----------------------------------------------------------------------------------------------
    def acquire(self, slot: int) -> hm_read_batch:
        b = hm_read_batch()
        self._check(self.lib.hm_batch_acquire(self.h, slot, C.byref(b)), "hm_batch_acquire")
        return b

    def stage(self, slot: int, batch) -> int:
        """Copies a synth.ReadBatch (host numpy) into the slot's pinned staging buffers; returns n_reads."""
        b = self.acquire(slot)
        n, nb = batch.n_reads, batch.n_bases
        if n > b.max_reads or nb > b.max_bases:
            raise HmError(f"batch ({n} reads, {nb} bases) exceeds slot capacity ({b.max_reads}, {b.max_bases})")
        _view(b.base_off, n + 1, np.uint32)[:] = batch.base_off
        _view(b.seq_off, n + 1, np.uint32)[:] = batch.seq_off
        _view(b.seq4, len(batch.seq4), np.uint8)[:] = batch.seq4
        _view(b.flag, n, np.uint16)[:] = batch.flag
        _view(b.valid, n, np.uint8)[:] = batch.valid
        for name in ("fi", "fp", "ri", "rp"):
            _view(getattr(b, name), nb, np.uint8)[:] = getattr(batch, name)
        return n

    def stage_records(self, slot: int, bodies) -> int:
        """Packs BAM record bodies with hm_pack_record (the reference's EvalKmerFeaturesGenerator::init rules)."""
        b = self.acquire(slot)
        n = C.c_uint32(0)
        for body in bodies:
            src = np.frombuffer(body, np.uint8)
            self._check(self.lib.hm_pack_record(C.byref(b), C.byref(n), src.ctypes.data_as(_u8p), len(body), self.min_read_len),
                        "hm_pack_record")
        return n.value

    def submit(self, slot: int, n_reads: int, flags: int = 0):
        self._check(self.lib.hm_batch_submit(self.h, slot, n_reads, flags), "hm_batch_submit")

    def collect(self, slot: int, copy: bool = True) -> CallBatch:
        c = hm_call_batch()
        self._check(self.lib.hm_batch_collect(self.h, slot, C.byref(c)), "hm_batch_collect")
        f = (lambda a: a.copy()) if copy else (lambda a: a)
        out = CallBatch(c.n_reads, c.n_calls, f(_view(c.call_off, c.n_reads + 1, np.uint32)), f(_view(c.n_fwd, c.n_reads, np.uint32)),
                        f(_view(c.qoff, c.n_calls, np.int32)), f(_view(c.ml, c.n_calls, np.uint8)), tuple(int(x) for x in c.n_sites))
        if c.mm_off:
            out.mm_off = f(_view(c.mm_off, c.n_reads + 1, np.uint32))
            out.mm_fwd_len = f(_view(c.mm_fwd_len, c.n_reads, np.uint32))
            out.mm_text = f(_view(c.mm_text, int(out.mm_off[-1]), np.uint8))
        if c.ml_hist:
            out.ml_hist = f(_view(c.ml_hist, 3 * 256, np.uint32)).reshape(3, 256)
        if c.read_stats:
            raw = np.ctypeslib.as_array(C.cast(c.read_stats, _u8p), shape=(max(c.n_reads, 1) * C.sizeof(hm_read_stats),))[:c.n_reads * C.sizeof(hm_read_stats)]
            rec = raw.view(np.dtype([("sum", "<u8", 4), ("max", "<u4", 4)]))
            out.stats_sum, out.stats_max = rec["sum"].copy(), rec["max"].copy()
        return out

    def timing(self, slot: int) -> hm_timing:
        t = hm_timing()
        self._check(self.lib.hm_batch_timing(self.h, slot, C.byref(t)), "hm_batch_timing")
        return t

    def call(self, batch, slot: int = 0, flags: int = 0) -> CallBatch:
        """Public one-shot API: host batch in, host calls out (H2D + kernels + D2H)."""
        n = self.stage(slot, batch)
        self.submit(slot, n, flags)
        return self.collect(slot)

    # -- validation hooks ---------------------------------------------------------------------------------------
    def dump_decode(self, slot: int, n_bases: int):
        k = [np.empty(n_bases, np.uint16) for _ in range(4)]
        fwd = np.empty(n_bases, np.uint8)
        rev = np.empty(n_bases, np.uint8)
        self._check(self.lib.hm_debug_dump_decode(self.h, slot, *[a.ctypes.data_as(_u16p) for a in k], fwd.ctypes.data_as(_u8p),
                                                  rev.ctypes.data_as(_u8p)), "hm_debug_dump_decode")
        return dict(fi=k[0], fp=k[1], ri=k[2], rp=k[3], fwd_qs=fwd, rev_qs=rev)

    def dump_ctx(self, slot: int, n_calls: int) -> np.ndarray:
        out = np.empty(max(n_calls, 1), np.uint8)
        self._check(self.lib.hm_debug_dump_ctx(self.h, slot, out.ctypes.data_as(_u8p)), "hm_debug_dump_ctx")
        return out[:n_calls]

    def dump_features(self, slot: int, first: int, count: int) -> np.ndarray:
        out = np.empty((count, 401, 8), np.float32)
        self._check(self.lib.hm_debug_dump_features(self.h, slot, first, count, out.ctypes.data_as(_f32p)), "hm_debug_dump_features")
        return out

    def dump_xmap(self, slot: int, first: int, count: int) -> np.ndarray:
        """[count,401,8] windows as the product CNN path holds them (X map, bf16 hi + lo re-summed)."""
        out = np.empty((count, 401, 8), np.float32)
        self._check(self.lib.hm_debug_dump_xmap(self.h, slot, C.c_uint32(first), C.c_uint32(count), out.ctypes.data_as(_f32p)), "hm_debug_dump_xmap")
        return out

    def dump_acts(self, slot: int, ctx: int, layer: int, first: int, count: int) -> np.ndarray:
        """[count, n_l, C_l] per-site output of conv layer 1..8 as assembled from the product path's maps (NaN = never stored)."""
        out = np.empty(count * 197 * 128, np.float32)
        npos, ch = C.c_int32(), C.c_int32()
        self._check(self.lib.hm_debug_dump_acts(self.h, slot, ctx, layer, C.c_uint32(first), C.c_uint32(count), out.ctypes.data_as(_f32p),
                                                C.c_size_t(out.size), C.byref(npos), C.byref(ch)), "hm_debug_dump_acts")
        return out[:count * npos.value * ch.value].reshape(count, npos.value, ch.value).copy()

    def dump_logits(self, slot: int, n_calls: int) -> np.ndarray:
        out = np.empty((max(n_calls, 1), 2), np.float32)
        self._check(self.lib.hm_debug_dump_logits(self.h, slot, out.ctypes.data_as(_f32p)), "hm_debug_dump_logits")
        return out[:n_calls]

    def microbench(self, slot: int, name: str, n_sites: int = 0, iters: int = 10):
        ms = C.c_float()
        by = C.c_double()
        fl = C.c_double()
        self._check(self.lib.hm_microbench(self.h, slot, name.encode(), n_sites, iters, C.byref(ms), C.byref(by), C.byref(fl)),
                    "hm_microbench")
        return ms.value, by.value, fl.value


def model_weights(path) -> tuple:
    """hm_model_weights: (flattened parameters in graph order, conv1 kernel size) of an .onnx or .pt model file."""
    L = load_library()
    n, k = C.c_size_t(), C.c_int32()
    if L.hm_model_weights(str(path).encode(), None, 0, C.byref(n), C.byref(k)) != 0:
        raise HmError(L.hm_last_error(None).decode())
    out = np.empty(n.value, np.float32)
    if L.hm_model_weights(str(path).encode(), out.ctypes.data_as(_f32p), n.value, C.byref(n), C.byref(k)) != 0:
        raise HmError(L.hm_last_error(None).decode())
    return out, k.value


def build_mod_record(body: bytes, keep_kinetics: bool, fwd_qoff, fwd_ml, rev_qoff, rev_ml) -> bytes:
    """hm_build_mod_record: the reference's build_one_mod_bam on a record body (host helper of the ABI)."""
    L = load_library()
    fq = np.ascontiguousarray(fwd_qoff, np.int32)
    rq = np.ascontiguousarray(rev_qoff, np.int32)
    fm = np.ascontiguousarray(fwd_ml, np.uint8)
    rm = np.ascontiguousarray(rev_ml, np.uint8)
    src = np.frombuffer(body, np.uint8)
    out = np.empty(L.hm_mod_record_bound(len(body), len(fq) + len(rq)), np.uint8)
    n = C.c_size_t()
    rc = L.hm_build_mod_record(src.ctypes.data_as(_u8p), len(body), int(keep_kinetics), fq.ctypes.data_as(_i32p), fm.ctypes.data_as(_u8p),
                               len(fq), rq.ctypes.data_as(_i32p), rm.ctypes.data_as(_u8p), len(rq), out.ctypes.data_as(_u8p), C.byref(n))
    if rc != 0:
        raise HmError(f"hm_build_mod_record failed ({rc})")
    return out[:n.value].tobytes()


def ml_threshold(bins):
    """hm_ml_threshold: (threshold, samples in range) for one context's 256-bin ML histogram."""
    L = load_library()
    b = np.ascontiguousarray(bins, np.uint64)
    assert b.shape == (256,)
    n = C.c_uint64(0)
    t = L.hm_ml_threshold(b.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(n))
    return int(t), int(n.value)


def parse_mod_record(body: bytes):
    """hm_parse_mod_record: (qoff i32[], strand u8[], prob u8[], codes bytes) of a record's MM/ML tags."""
    L = load_library()
    src = np.frombuffer(body, np.uint8)
    n = C.c_uint32(0)
    rc = L.hm_parse_mod_record(src.ctypes.data_as(_u8p), len(body), None, None, None, None, 0, C.byref(n))
    if rc != 0:
        raise HmError(f"hm_parse_mod_record failed ({rc})")
    q, s, p = np.zeros(n.value, np.int32), np.zeros(n.value, np.uint8), np.zeros(n.value, np.uint8)
    codes = C.create_string_buffer(max(n.value, 1))
    rc = L.hm_parse_mod_record(src.ctypes.data_as(_u8p), len(body), q.ctypes.data_as(_i32p), s.ctypes.data_as(_u8p), p.ctypes.data_as(_u8p),
                               codes, n.value, C.byref(n))
    if rc != 0:
        raise HmError(f"hm_parse_mod_record failed ({rc})")
    return q, s, p, codes.raw[:n.value]


def pack_records_host(bodies, min_read_len: int = 1000, max_bases: int = None, threads: int = 0):
    """hm_pack_record into plain numpy buffers (no GPU needed): returns a synth.ReadBatch.  threads > 0 packs the whole list
    with one hm_pack_records call instead (the `call` driver's path); records it refuses (-1) are left out."""
    from .synth import ReadBatch

    L = load_library()
    n_max = len(bodies)
    nb_max = max_bases or sum(len(b) for b in bodies)
    arr = dict(base_off=np.zeros(n_max + 1, np.uint32), seq_off=np.zeros(n_max + 1, np.uint32), seq4=np.zeros(nb_max // 2 + n_max + 16, np.uint8),
               flag=np.zeros(n_max, np.uint16), valid=np.zeros(n_max, np.uint8), fi=np.zeros(nb_max, np.uint8), fp=np.zeros(nb_max, np.uint8),
               ri=np.zeros(nb_max, np.uint8), rp=np.zeros(nb_max, np.uint8))
    b = hm_read_batch(n_max, nb_max, arr["base_off"].ctypes.data_as(_u32p), arr["seq_off"].ctypes.data_as(_u32p), arr["seq4"].ctypes.data_as(_u8p),
                      arr["flag"].ctypes.data_as(_u16p), arr["valid"].ctypes.data_as(_u8p), arr["fi"].ctypes.data_as(_u8p),
                      arr["fp"].ctypes.data_as(_u8p), arr["ri"].ctypes.data_as(_u8p), arr["rp"].ctypes.data_as(_u8p))
    n = C.c_uint32(0)
    if threads > 0:
        keep = [np.frombuffer(body, np.uint8) for body in bodies]
        ptrs = (_u8p * len(keep))(*[k.ctypes.data_as(_u8p) for k in keep])
        lens = (C.c_size_t * len(keep))(*[len(body) for body in bodies])
        idx = np.full(len(keep), -2, np.int32)
        rc = L.hm_pack_records(C.byref(b), len(keep), ptrs, lens, min_read_len, threads, idx.ctypes.data_as(_i32p), C.byref(n))
        if rc != 0:
            raise HmError(f"hm_pack_records failed ({rc})")
        assert (idx[idx >= 0] == np.arange(n.value)).all()
    else:
        for body in bodies:
            src = np.frombuffer(body, np.uint8)
            rc = L.hm_pack_record(C.byref(b), C.byref(n), src.ctypes.data_as(_u8p), len(body), min_read_len)
            if rc != 0:
                raise HmError(f"hm_pack_record failed ({rc})")
    nb = int(arr["base_off"][n.value])
    ns = int(arr["seq_off"][n.value])
    return ReadBatch(n.value, arr["base_off"][:n.value + 1], arr["seq_off"][:n.value + 1], arr["seq4"][:ns], arr["flag"][:n.value],
                     arr["valid"][:n.value], arr["fi"][:nb], arr["fp"][:nb], arr["ri"][:nb], arr["rp"][:nb])


def debug_dense_op(srcs, terms, bias, rows: int, conv1_taps: int = 0, w2=None, b2=None, gather_rows=None, gather_mask: int = 0,
                   device: int = 0) -> np.ndarray:
    """hm_debug_dense_op: one op of the tensor-core dense plan on caller data (unit test of dense_gemm_kernel).

    srcs: list of [rows_alloc, cin] f32 maps; terms: list of (src index, row shift, W) with W [cin, cout]
    (conv1 form: one term, W [taps, 8, cout]); returns [rows, cout] f32, or [rows, 2] when w2/b2 select the head form.
    gather_rows [rows] u32 + gather_mask: terms whose bit is set read row gather_rows[r] + shift (compact ops)."""
    L = load_library()
    srcs = [np.ascontiguousarray(a, np.float32) for a in srcs]
    rows_alloc, cin = srcs[0].shape
    w = np.ascontiguousarray(np.stack([np.asarray(t[2], np.float32) for t in terms]))
    cout = w.shape[-1]
    bias = np.ascontiguousarray(bias, np.float32)
    tsrc = np.array([t[0] for t in terms], np.int32)
    tsh = np.array([t[1] for t in terms], np.int32)
    ptrs = (_f32p * len(srcs))(*[a.ctypes.data_as(_f32p) for a in srcs])
    head = w2 is not None
    gr = np.ascontiguousarray(gather_rows, np.uint32) if gather_rows is not None else None
    out = np.empty((rows, 2 if head else cout), np.float32)
    if head:
        w2 = np.ascontiguousarray(w2, np.float32)
        b2 = np.ascontiguousarray(b2, np.float32)
    rc = L.hm_debug_dense_op(device, rows, rows_alloc, cin, cout, len(srcs), ptrs, len(terms), tsrc.ctypes.data_as(_i32p),
                             tsh.ctypes.data_as(_i32p), w.ctypes.data_as(_f32p), bias.ctypes.data_as(_f32p), conv1_taps,
                             w2.ctypes.data_as(_f32p) if head else None, b2.ctypes.data_as(_f32p) if head else None,
                             gr.ctypes.data_as(_u32p) if gr is not None else None, gather_mask, out.ctypes.data_as(_f32p))
    if rc != 0:
        raise HmError(f"hm_debug_dense_op failed ({rc}): {L.hm_last_error(None).decode()}")
    return out


def build_mod_record_mm(body: bytes, keep_kinetics: bool, mm_fwd, mm_rev, ml, n_fwd: int, n_rev: int) -> bytes:
    """hm_build_mod_record_mm: the record from device-built MM text (no per-base loop on the host)."""
    L = load_library()
    src = np.frombuffer(body, np.uint8)
    as_u8 = lambda a: np.frombuffer(a, np.uint8) if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a, np.uint8)
    f, r, m = as_u8(mm_fwd), as_u8(mm_rev), as_u8(ml)
    out = np.empty(len(body) + 64 + len(f) + len(r) + len(m), np.uint8)
    n = C.c_size_t()
    rc = L.hm_build_mod_record_mm(src.ctypes.data_as(_u8p), len(body), int(keep_kinetics), f.ctypes.data_as(_u8p), len(f), r.ctypes.data_as(_u8p),
                                  len(r), m.ctypes.data_as(_u8p), n_fwd, n_rev, out.ctypes.data_as(_u8p), C.byref(n))
    if rc != 0:
        raise HmError(f"hm_build_mod_record_mm failed ({rc})")
    return out[:n.value].tobytes()
