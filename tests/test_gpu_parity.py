"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): decoded kinetics, base codes, site lists, call order and feature tensors BIT-EXACT;
per-site probability within 1e-3 absolute; ML byte within +-1.
"""
import os
import numpy as np
import pytest

from hifimeth_b200 import engine as hme
from hifimeth_b200 import synth
from oracle import cnn_oracle, hmoracle
from oracle.make_golden import golden_bodies, golden_reads

from conftest import ROOT

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-3  # absolute, north_star
ML_TOL = 1       # bytes, north_star

MODES = [pytest.param(hme.HM_CNN_TENSOR, id="tensor"), pytest.param(hme.HM_CNN_FP32_SIMT, id="fp32")]


@pytest.fixture(scope="module")
def models():
    return cnn_oracle.load_models(ROOT / "models")


@pytest.fixture(scope="module", params=MODES)
def eng(request, lib_built):
    e = hme.Engine(max_reads=64, max_bases=1 << 20, cnn_mode=request.param, keep_debug=True)
    yield e
    e.close()


def _check_batch(eng, batch, models, ctx_mask=7, feature_samples=64, check_cnn=True):
    O = hmoracle.oracle()
    want = O.batch_call(batch, models, ctx_mask) if check_cnn else O.batch_sites(batch, ctx_mask)
    got = eng.call(batch, slot=0)
    assert got.n_reads == batch.n_reads
    # ---- decode: bit-exact -----------------------------------------------------------------------------------
    dec = eng.dump_decode(0, batch.n_bases)
    for name in ("fi", "fp", "ri", "rp"):
        assert (dec[name] == O.decode_plane(getattr(batch, name))).all(), name
    for r in range(batch.n_reads):
        a, b = int(batch.base_off[r]), int(batch.base_off[r + 1])
        assert (dec["fwd_qs"][a:b] == want[r]["fwd"]).all() and (dec["rev_qs"][a:b] == want[r]["rev"]).all(), r
    # ---- site lists and order: bit-exact -----------------------------------------------------------------------
    n_calls = sum(len(w["qoff"]) for w in want)
    assert got.n_calls == n_calls
    ctx = eng.dump_ctx(0, got.n_calls)
    off = 0
    for r, w in enumerate(want):
        assert int(got.call_off[r]) == off
        n = len(w["qoff"])
        assert int(got.n_fwd[r]) == w["n_fwd"]
        assert (got.qoff[off:off + n] == w["qoff"]).all(), r
        assert (ctx[off:off + n] == w["ctx"]).all(), r
        off += n
    assert int(got.call_off[batch.n_reads]) == n_calls
    wctx = np.concatenate([w["ctx"] for w in want]) if want else np.zeros(0, np.uint8)
    assert got.n_sites == tuple(int((wctx == c).sum()) for c in range(3))
    if n_calls == 0:
        return got, want
    # ---- feature tensors: bit-exact ------------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    picks = np.unique(np.r_[0, n_calls - 1, rng.integers(0, n_calls, size=min(feature_samples, n_calls))])
    read_of = np.searchsorted(got.call_off, picks, side="right") - 1
    for k, r in zip(picks, read_of):
        f_gpu = eng.dump_features(0, int(k), 1)[0]
        f_cpu = O.batch_features(batch, want, int(r), np.array([int(k - got.call_off[r])]))[0]
        assert (f_gpu.view(np.uint32) == f_cpu.view(np.uint32)).all(), (k, r)
    first = int(picks[len(picks) // 2])
    cnt = min(5, n_calls - first)
    blk = eng.dump_features(0, first, cnt)
    for j in range(cnt):
        assert (blk[j].view(np.uint32) == eng.dump_features(0, first + j, 1)[0].view(np.uint32)).all()
    if not check_cnn:
        return got, want
    # ---- CNN: probability within 1e-3, ML within +-1 -----------------------------------------------------------------
    logits = eng.dump_logits(0, got.n_calls)
    w_logits = np.concatenate([w["logits"] for w in want])
    w_prob = np.concatenate([w["prob"] for w in want])
    w_ml = np.concatenate([w["ml"] for w in want])
    p_gpu, _ = cnn_oracle.logits_to_prob_ml(logits)
    assert np.abs(p_gpu - w_prob).max() <= PROB_TOL, (np.abs(p_gpu - w_prob).max(), np.abs(logits - w_logits).max())
    assert np.abs(got.ml.astype(np.int32) - w_ml.astype(np.int32)).max() <= ML_TOL
    # the byte the engine wrote is the truncation of ITS OWN probability (mod_batch.cpp:59), up to expf ulps
    assert np.abs(got.ml.astype(np.int32) - np.minimum(255, (255 * p_gpu).astype(np.int32))).max() <= 1
    return got, want


def test_golden_reads_all_contexts(eng, models, golden):
    """The committed golden reads (flag 0x10, N bases, B:S kinetics, short read, kinetics-less read)."""
    reads = golden_reads()
    batch = hme.pack_records_host(golden_bodies(reads), min_read_len=1000)
    got, want = _check_batch(eng, batch, models)
    # site lists against the REFERENCE-generated golden vectors directly
    ctx = eng.dump_ctx(0, got.n_calls)
    for i in range(batch.n_reads):
        a, b = int(got.call_off[i]), int(got.call_off[i + 1])
        if not batch.valid[i]:
            assert a == b
            continue
        for c in range(3):
            assert (np.sort(got.qoff[a:b][ctx[a:b] == c]) == golden[f"sites{i}_{c}"]).all(), (i, c)
    # logits against the golden logits (reference TorchScript for CpG/CHH, fp32 ONNX forward for CHG)
    logits = eng.dump_logits(0, got.n_calls)
    n = 0
    for i in range(batch.n_reads):
        a, b = int(got.call_off[i]), int(got.call_off[i + 1])
        if not batch.valid[i]:
            continue  # below -l: the reference's scan still lists sites, the worker never calls them (mod_main.cpp:189-192)
        for c in range(3):
            if f"sel{i}_{c}" not in golden:
                continue
            sites = golden[f"sites{i}_{c}"][golden[f"sel{i}_{c}"]]
            key = "ptlogits" if f"ptlogits{i}_{c}" in golden else "logits"
            for s, lg in zip(sites, golden[f"{key}{i}_{c}"]):
                k = a + int(np.nonzero(got.qoff[a:b] == s)[0][0])
                p_g, _ = cnn_oracle.logits_to_prob_ml(logits[k:k + 1])
                p_w, ml_w = cnn_oracle.logits_to_prob_ml(lg[None])
                assert abs(float(p_g[0]) - float(p_w[0])) <= PROB_TOL
                assert abs(int(got.ml[k]) - int(ml_w[0])) <= ML_TOL
                n += 1
    assert n > 100


@pytest.mark.parametrize("cnn_mode", MODES)
@pytest.mark.parametrize("ctx_mask", [1, 2, 4, 5])
def test_context_masks(lib_built, models, ctx_mask, cnn_mode):
    batch, _ = synth.make_reads(3, (1000, 1800), seed=11 + ctx_mask, flag_rev_every=2)
    e2 = hme.Engine(ctx_mask=ctx_mask, max_reads=8, max_bases=1 << 16, cnn_mode=cnn_mode, keep_debug=True)
    try:
        _check_batch(e2, batch, models, ctx_mask=ctx_mask, feature_samples=8)
    finally:
        e2.close()


def test_ragged_and_edge_reads(eng, models):
    """Reads of length exactly -l, reads that are all one base, sites in the first / last 200 bases (clipped windows)."""
    rng = np.random.default_rng(3)
    reads = []
    for i, l in enumerate([1000, 1001, 1399, 2048, 1024, 3000]):
        seq = rng.integers(0, 4, size=l, dtype=np.uint8)
        if i == 3:
            seq[:] = 1  # poly-C: every position but the last two is a CHH site
        if i == 4:
            seq[:] = np.tile(np.array([1, 2], np.uint8), l // 2)  # CGCG...: CpG on every C
        k = rng.integers(0, 256, size=(4, l), dtype=np.uint8)
        reads.append(dict(name=f"edge/{i}", seq=seq, fi=k[0], ri=k[1], fp=k[2], rp=k[3], flag=4 | (16 if i % 2 else 0), np=5, rq=0.99, zm=i))
    _check_batch(eng, synth.soa_from_reads(reads), models, feature_samples=96)


def test_empty_and_passthrough_batches(eng, models):
    empty = synth.soa_from_reads([])
    got = eng.call(empty, slot=1)
    assert got.n_reads == 0 and got.n_calls == 0
    batch, _ = synth.make_reads(4, 400, seed=2)  # all shorter than -l: no calls, nothing launched on the CNN
    got = eng.call(batch, slot=1)
    assert got.n_calls == 0 and (got.call_off == 0).all() and got.n_sites == (0, 0, 0)


def test_slots_are_independent_and_rerun_is_idempotent(eng, models):
    a, _ = synth.make_reads(2, 1500, seed=21)
    b, _ = synth.make_reads(3, 1200, seed=22, flag_rev_every=1)
    na = eng.stage(0, a)
    nb = eng.stage(1, b)
    eng.submit(0, na)
    eng.submit(1, nb)
    ra, rb = eng.collect(0), eng.collect(1)
    O = hmoracle.oracle()
    for got, batch in ((ra, a), (rb, b)):
        want = O.batch_sites(batch, 7)
        assert (got.qoff == np.concatenate([w["qoff"] for w in want])).all()
    # re-run on resident inputs (kernel-only path of bench.py) reproduces the same bytes
    eng.submit(0, na, hme.HM_SUBMIT_SKIP_H2D)
    ra2 = eng.collect(0)
    assert (ra2.qoff == ra.qoff).all() and (ra2.ml == ra.ml).all()
    t = eng.timing(0)
    assert t.kernel_launches > 0 and t.total_ms > 0 and t.h2d_bytes == 0


def test_mod_records_round_trip(eng, models):
    """A8 through the ABI's host helper: engine calls -> MM/ML/MN record == the oracle's record, byte for byte."""
    batch, reads = synth.make_reads(3, (1000, 1600), seed=31, flag_rev_every=2)
    got = eng.call(batch, slot=0)
    O = hmoracle.oracle()
    for r, rd in enumerate(reads):
        body = synth.record_body(rd)
        fq, fml, rq, rml = got.read_calls(r)
        rec = hme.build_mod_record(body, False, fq, fml, rq, rml)
        assert rec == O.build_mod_record(body, False, fq, fml, rq, rml)
        assert b"MMZC+m," in rec and b"fiBC" not in rec


def test_microbench_kernels_run(eng):
    batch, _ = synth.make_reads(4, 2000, seed=41)
    eng.call(batch, slot=0)
    for name in ("decode", "scan", "gather", "cnn", "mm"):
        ms, by, fl = eng.microbench(0, name, 0, 2)
        assert ms > 0 and (by > 0 or fl > 0)


def test_mm_text_built_on_device(eng, models):
    """Row N1: the MM skip-count text comes from the device; records assembled from it equal the oracle's build_one_mod_bam
    restatement byte for byte (which test_oracle.py pins to the reference's own compiled build_mod_bam.cpp)."""
    reads = golden_reads()
    bodies = golden_bodies(reads)
    batch = hme.pack_records_host(bodies, min_read_len=1000)
    got = eng.call(batch, slot=0, flags=hme.HM_SUBMIT_MM_TEXT)
    assert got.mm_off is not None and int(got.mm_off[0]) == 0
    O = hmoracle.oracle()
    n_with = 0
    for r, body in enumerate(bodies):
        fq, fml, rq, rml = got.read_calls(r)
        mm_f, mm_r = got.read_mm(r)
        a, b = int(got.call_off[r]), int(got.call_off[r + 1])
        rec = hme.build_mod_record_mm(body, False, mm_f, mm_r, got.ml[a:b], len(fq), len(rq))
        assert rec == O.build_mod_record(body, False, fq, fml, rq, rml), r
        assert rec == hme.build_mod_record(body, False, fq, fml, rq, rml), r
        n_with += len(fq) + len(rq) > 0
    assert n_with >= 4
    # a context subset skips C's and G's between calls: counts above 9 exercise multi-digit text
    e2 = hme.Engine(ctx_mask=1, max_reads=16, max_bases=1 << 18, keep_debug=True)
    try:
        b2, rd = synth.make_reads(3, (1500, 2500), seed=91, flag_rev_every=2)
        g2 = e2.call(b2, flags=hme.HM_SUBMIT_MM_TEXT)
        for r, rr in enumerate(rd):
            body = synth.record_body(rr)
            fq, fml, rq, rml = g2.read_calls(r)
            mm_f, mm_r = g2.read_mm(r)
            a, b = int(g2.call_off[r]), int(g2.call_off[r + 1])
            assert hme.build_mod_record_mm(body, True, mm_f, mm_r, g2.ml[a:b], len(fq), len(rq)) == O.build_mod_record(body, True, fq, fml, rq, rml)
        assert any(bytes(g2.read_mm(r)[0]).count(b",1") for r in range(3))
    finally:
        e2.close()
    # no calls at all: empty text
    short, _ = synth.make_reads(2, 400, seed=3)
    g3 = eng.call(short, slot=1, flags=hme.HM_SUBMIT_MM_TEXT)
    assert g3.n_calls == 0 and int(g3.mm_off[-1]) == 0


def test_call_cli_end_to_end(lib_built, models, tmp_path):
    """`hifimeth-b200 call in.bam out.bam`: input order kept, every record written, tags stripped, MM/ML/MN equal to what the
    oracle's build_one_mod_bam restatement makes of the engine's calls, ML within +-1 of the CPU oracle's, @PG line added."""
    import subprocess

    reads = golden_reads()
    _, more = synth.make_reads(9, (900, 2600), seed=404, flag_rev_every=4)
    all_reads = reads + more
    bodies = golden_bodies(reads) + [synth.record_body(r) for r in more]
    src, dst = tmp_path / "in.bam", tmp_path / "mod.bam"
    synth.write_bam(src, bodies)
    exe = hme.PKG / "bin" / "hifimeth-b200"
    # small batches (-b 4, 8 kb of bases) force several batches, a batch cut by bases, and the two-slot pipeline; two workers on
    # the same device finish batches out of order, the writer must restore the input order
    r = subprocess.run([str(exe), "call", "-b", "4", "--max-bases", "8192", "-t", "3", "--devices", "0,0", str(src), str(dst)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "on 2 worker(s)" in r.stderr and "CpG scaled probability threshold (pileup rule): 128" in r.stderr
    text, _, got = synth.read_bam(dst)
    assert "@PG\tID:hifimeth\tPN:hifimeth\tVN:1.1.0\tCL:" in text and text.startswith("@HD")
    assert len(got) == len(bodies)
    # expected records: engine calls through the Python mirror + the oracle's record builder
    batch = hme.pack_records_host(bodies, min_read_len=1000)
    eng = hme.Engine(max_reads=64, max_bases=1 << 20)
    try:
        calls = eng.call(batch)
    finally:
        eng.close()
    O = hmoracle.oracle()
    want = O.batch_call(batch, models)
    n_called = 0
    for i, body in enumerate(bodies):
        fq, fml, rq, rml = calls.read_calls(i)
        assert got[i] == O.build_mod_record(body, False, fq, fml, rq, rml), i
        if len(fq) + len(rq):
            n_called += 1
            ml_cli = np.frombuffer(got[i][got[i].index(b"MLBC") + 8:][:len(fq) + len(rq)], np.uint8)
            assert np.abs(ml_cli.astype(int) - want[i]["ml"].astype(int)).max() <= ML_TOL
            assert b"fiBC" not in got[i] and b"MMZC+m" in got[i]
    assert n_called >= 10
    # -k keeps the kinetics, -c cpg restricts the contexts
    r = subprocess.run([str(exe), "call", "-k", "-c", "cpg", str(src), str(dst)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    _, _, got_k = synth.read_bam(dst)
    assert len(got_k) == len(bodies) and any(b"fiBC" in g for g in got_k)
    e1 = hme.Engine(ctx_mask=1, max_reads=64, max_bases=1 << 20)
    try:
        c1 = e1.call(batch)
    finally:
        e1.close()
    for i, body in enumerate(bodies):
        fq, fml, rq, rml = c1.read_calls(i)
        assert got_k[i] == O.build_mod_record(body, True, fq, fml, rq, rml), i


def test_ml_histogram_on_device(eng, models):
    """Row N3: HM_SUBMIT_ML_HIST returns the per-context histograms of the batch's ML bytes over reads without flag 0x900, equal
    to the oracle's accumulation over the engine's own calls; without the flag the pointer is NULL."""
    batch, reads = synth.make_reads(12, (1000, 5000), seed=909, flag_rev_every=3)
    batch.flag[2] |= 0x100   # secondary and supplementary records are left out of the histograms (pileup.cpp:237)
    batch.flag[7] |= 0x800
    got = eng.call(batch, flags=hme.HM_SUBMIT_ML_HIST)
    assert got.ml_hist is not None and got.ml_hist.shape == (3, 256)
    ctx = eng.dump_ctx(0, got.n_calls)
    call_read = np.repeat(np.arange(got.n_reads, dtype=np.uint32), np.diff(got.call_off.astype(np.int64)))
    want = hmoracle.oracle().ml_histogram(got.ml, ctx, batch.flag, call_read)
    assert (got.ml_hist.astype(np.uint64) == want).all()
    keep = ~np.isin(call_read, [2, 7])
    assert int(got.ml_hist.sum()) == int(keep.sum()) < got.n_calls
    for c in range(3):
        assert (got.ml_hist[c] == np.bincount(got.ml[keep & (ctx == c)], minlength=256)).all()
    assert eng.call(batch, slot=1).ml_hist is None


def test_sub_batches_and_compact_groups_do_not_change_results(lib_built, monkeypatch):
    """A batch cut into many sub-batches (HM_DENSE_ROWS small: dense maps of 16 Ki rows, compact buffers of 8 Ki sites, so the
    dense chains run per sub-batch and the compact chain per GROUP of sub-batches, several groups per context) gives
    bit-identical calls and logits to the same batch evaluated in one piece."""
    batch, _ = synth.make_reads(40, (1000, 6000), seed=31337, flag_rev_every=3, short_every=9)
    whole = hme.Engine(max_reads=64, max_bases=1 << 19, keep_debug=True)
    try:
        a = whole.call(batch)
        la = whole.dump_logits(0, a.n_calls)
    finally:
        whole.close()
    monkeypatch.setenv("HM_DENSE_ROWS", "16384")
    cut = hme.Engine(max_reads=64, max_bases=1 << 19, keep_debug=True)
    try:
        b = cut.call(batch)
        lb = cut.dump_logits(0, b.n_calls)
        launches = cut.timing(0).kernel_launches
    finally:
        cut.close()
    assert a.n_calls == b.n_calls > 30000
    assert (a.call_off == b.call_off).all() and (a.qoff == b.qoff).all()
    assert (la.view(np.uint32) == lb.view(np.uint32)).all()
    assert (a.ml == b.ml).all()
    assert launches > 300  # really many sub-batches


def test_full_size_config1_batch(lib_built, models):
    """BASELINE.json configs[1] at full size (1 000 reads x 15 kb, all contexts, ~5.86 M sites, 15 sub-batches, several compact
    groups per context): site lists, call order and contexts bit-exact against the oracle for EVERY read; per-read offsets
    consistent; probabilities within 1e-3 and ML within +-1 on a random sample of 1 500 sites per context; the MM text of a sample
    of reads equal to the oracle's; ML histograms summing to the calls."""
    O = hmoracle.oracle()
    batch, _ = synth.make_reads(1000, 15000, seed=20261)
    eng = hme.Engine(n_slots=1, max_reads=1000, max_bases=batch.n_bases + 1024, keep_debug=True)
    try:
        got = eng.call(batch, flags=hme.HM_SUBMIT_MM_TEXT | hme.HM_SUBMIT_ML_HIST)
        ctx = eng.dump_ctx(0, got.n_calls)
        logits = eng.dump_logits(0, got.n_calls)
        launches = eng.timing(0).kernel_launches
    finally:
        eng.close()
    want = O.batch_sites(batch, 7)
    assert got.n_calls == sum(len(w["qoff"]) for w in want) > 5_800_000
    assert launches > 250  # 15 sub-batches x 3 contexts of dense chains (5 launches each) + F1, G1 and one chain kernel per compact group
    off = 0
    for r, w in enumerate(want):
        n = len(w["qoff"])
        assert int(got.call_off[r]) == off and int(got.n_fwd[r]) == w["n_fwd"], r
        assert (got.qoff[off:off + n] == w["qoff"]).all() and (ctx[off:off + n] == w["ctx"]).all(), r
        off += n
    assert got.n_sites == tuple(int((ctx == c).sum()) for c in range(3))
    assert int(got.ml_hist.sum()) == got.n_calls
    for c in range(3):
        assert (got.ml_hist[c] == np.bincount(got.ml[ctx == c], minlength=256)).all()
    # MM text of a few reads against the oracle's build_mm (C+m,...;G-m,...;)
    for r in (0, 499, 999):
        a, nf = int(got.call_off[r]), int(got.n_fwd[r])
        b = int(got.call_off[r + 1])
        fwd_txt, rev_txt = got.read_mm(r)
        mm = O.build_mm(want[r]["fwd"], got.qoff[a:a + nf], got.qoff[a + nf:b])
        assert mm == b"C+m" + fwd_txt.tobytes() + b";G-m" + rev_txt.tobytes() + b";", r
    # CNN on a random sample per context
    rng = np.random.default_rng(11)
    p_gpu, _ = cnn_oracle.logits_to_prob_ml(logits)
    for c in range(3):
        idx = np.sort(rng.choice(np.nonzero(ctx == c)[0], size=1500, replace=False))
        read_of = np.searchsorted(got.call_off, idx, side="right") - 1
        feats = np.concatenate([O.batch_features(batch, want, int(r), np.array([int(k - got.call_off[r])])) for k, r in zip(idx, read_of)])
        lg = cnn_oracle.forward_logits(models[c], feats)
        p_cpu, ml_cpu = cnn_oracle.logits_to_prob_ml(lg)
        assert np.abs(p_gpu[idx] - p_cpu).max() <= PROB_TOL, (c, float(np.abs(p_gpu[idx] - p_cpu).max()))
        assert np.abs(got.ml[idx].astype(np.int32) - ml_cpu.astype(np.int32)).max() <= ML_TOL, c


def test_call_cli_two_devices(lib_built, tmp_path):
    """`--devices 0,1`: one engine per GPU inside one process (per-device kernel attributes, per-thread current device), batches
    dealt by the host work queue; the output equals the single-device run record for record.  On a one-GPU box the second worker
    runs on device 0 as well (`--devices 0,0`): the same queue, two engines, batches finishing out of order."""
    import subprocess

    try:
        n_gpu = len([l for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if l.startswith("GPU ")])
    except OSError:
        n_gpu = 0
    two_devs = "0,1" if n_gpu >= 2 else "0,0"
    _, reads = synth.make_reads(24, (1000, 4000), seed=77, flag_rev_every=5)
    bodies = [synth.record_body(r) for r in reads]
    src, one, two = tmp_path / "in.bam", tmp_path / "one.bam", tmp_path / "two.bam"
    synth.write_bam(src, bodies)
    exe = hme.PKG / "bin" / "hifimeth-b200"
    for devs, dst in (("0", one), (two_devs, two)):
        r = subprocess.run([str(exe), "call", "-b", "3", "--devices", devs, str(src), str(dst)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert ("on 2 worker(s)" in r.stderr) == (devs != "0")
    assert synth.read_bam(one)[2] == synth.read_bam(two)[2]


def test_read_stats_diagnostics(eng):
    """Row A9 (diagnostics only): per-read sum / max of the decoded frames of fi, fp, ri, rp from the warp-reduction kernel equal
    the oracle's decode table applied on the host; asking for them changes nothing else; without the flag the pointer is NULL."""
    O = hmoracle.oracle()
    batch, _ = synth.make_reads(9, (300, 7000), seed=4242, flag_rev_every=2, uniform_codes=True)
    plain = eng.call(batch, slot=1)
    got = eng.call(batch, flags=hme.HM_SUBMIT_READ_STATS)
    assert plain.stats_sum is None and got.stats_sum.shape == (batch.n_reads, 4)
    assert (got.qoff == plain.qoff).all() and (got.ml == plain.ml).all()
    frames = [O.decode_plane(getattr(batch, k)).astype(np.uint64) for k in ("fi", "fp", "ri", "rp")]
    for r in range(batch.n_reads):
        a, b = int(batch.base_off[r]), int(batch.base_off[r + 1])
        for k in range(4):
            assert int(got.stats_sum[r, k]) == int(frames[k][a:b].sum()), (r, k)
            assert int(got.stats_max[r, k]) == int(frames[k][a:b].max()), (r, k)
    ms, by, _ = eng.microbench(0, "stats", 0, 3)
    assert ms > 0 and by == 8.0 * batch.n_bases


def test_product_feature_map_and_layer_activations_direct(lib_built, models):
    """the maps the PRODUCT path computes on, compared with the oracle directly instead of through |dp| <= 1e-3.

    X map (track_features_kernel, bf16 hi + lo): one-hot columns and out-of-read rows exact, kinetics within 2^-16 relative of the
    oracle's fp32 features; both strands, flag 0x10, reads of length exactly -l and -l + 1.
    Layers 1..8 (Y_l dense maps, F_l / G_l / T7 / T8 compact maps, the scatter copies of Y1): against
    cnn_oracle.forward_logits(return_acts=True) within 1e-3 of the layer's scale -- a wrong halo row, tap or strand flip is O(1)."""
    O = hmoracle.oracle()
    _, reads = synth.make_reads(5, (1000, 2600), seed=4242, flag_rev_every=2)
    _, r1000 = synth.make_reads(1, 1000, seed=4243)
    _, r1001 = synth.make_reads(1, 1001, seed=4244)
    batch = hme.pack_records_host([synth.record_body(r) for r in reads + r1000 + r1001], min_read_len=1000)
    assert list(batch.valid) == [1] * 7
    eng = hme.Engine(max_reads=16, max_bases=1 << 16, keep_debug=True)
    try:
        got = eng.call(batch)
        want = O.batch_call(batch, models)
        ctx = eng.dump_ctx(0, got.n_calls)
        n_calls = got.n_calls
        rng = np.random.default_rng(9)
        # every read contributes its first, last and a few random calls (first / last sites have clipped windows)
        picks = []
        for r in range(batch.n_reads):
            a, b = int(got.call_off[r]), int(got.call_off[r + 1])
            picks += [a, a + 1, b - 1, b - 2] + [int(x) for x in rng.integers(a, b, 6)]
        picks = sorted(set(picks))
        read_of = np.searchsorted(got.call_off, picks, side="right") - 1
        feats = {}
        for k, r in zip(picks, read_of):
            f_cpu = O.batch_features(batch, want, int(r), np.array([int(k - got.call_off[r])]))[0]
            feats[k] = f_cpu
            x = eng.dump_xmap(0, int(k), 1)[0]
            assert (x[:, :4] == f_cpu[:, :4]).all(), (k, "one-hot")
            outside = ~(f_cpu[:, :4].any(axis=1))  # rows beyond the read carry no base
            assert (x[outside] == 0).all() and (f_cpu[outside] == 0).all(), (k, "rows outside the read")
            assert (x[200, :4] == np.array([0, 1, 0, 0], np.float32)).all()  # centre row is the called C on its own strand
            err = np.abs(x[:, 4:].astype(np.float64) - f_cpu[:, 4:]) 
            assert (err <= np.abs(f_cpu[:, 4:]) * 2.0 ** -16).all(), (k, float(err.max()))
        # ---- per-layer activations --------------------------------------------------------------------------------------
        worst = {}
        for c in range(3):
            ks = [k for k in picks if ctx[k] == c]
            assert len(ks) >= 6, c
            f = np.stack([feats[k] for k in ks])
            _, acts = cnn_oracle.forward_logits(models[c], f, return_acts=True)
            for layer in range(1, 9):
                ref = np.transpose(acts[layer], (0, 2, 1))  # [B, C, n] -> [B, n, C]
                scale = max(1.0, float(np.abs(ref).max()))
                seen = 0
                for j, k in enumerate(ks):
                    a = eng.dump_acts(0, c, layer, int(k), 1)[0]
                    assert a.shape == ref[j].shape, (c, layer, a.shape, ref[j].shape)
                    have = ~np.isnan(a)
                    if layer >= 2:
                        assert have.all(), (c, layer)
                    else:
                        assert have[0].all() and have[-1].all() and have.any(axis=1).sum() >= 4  # F1, G1, the rows scattered for F2 / G2
                    d = np.abs(a[have] - ref[j][have]).max()
                    worst[(c, layer)] = max(worst.get((c, layer), 0.0), float(d) / scale)
                    seen += int(have.sum())
                assert seen > 0
        print("max |d act| / layer scale:", {k: f"{v:.1e}" for k, v in worst.items()})
        assert max(worst.values()) <= 1e-3, worst
        # the dump leaves the engine usable and the results unchanged
        again = eng.call(batch)
        assert (again.ml == got.ml).all() and (again.qoff == got.qoff).all()
    finally:
        eng.close()


def test_torchscript_weights_give_the_same_calls(lib_built, monkeypatch):
    """north_star: "the same models/{CpG,CHG,CHH}.onnx/.pt weights".  An engine created from the .pt exports (HM_MODEL_FORMAT=pt; the
    files the reference's app-gpu binary loads, 5mc_call_gpu.cpp:48) calls CHH bit-identically to the .onnx engine (the two files hold
    the same bits) and CpG within float noise (weights differ by <= 2.4e-7).  CHG.pt is another checkpoint and is not compared."""
    batch, _ = synth.make_reads(6, (1000, 3000), seed=31337, flag_rev_every=3)
    a = hme.Engine(ctx_mask=5, max_reads=16, max_bases=1 << 16, keep_debug=True)
    try:
        ra = a.call(batch)
        la = a.dump_logits(0, ra.n_calls)
        ctx = a.dump_ctx(0, ra.n_calls)
    finally:
        a.close()
    monkeypatch.setenv("HM_MODEL_FORMAT", "pt")
    b = hme.Engine(ctx_mask=5, max_reads=16, max_bases=1 << 16, keep_debug=True)
    try:
        rb = b.call(batch)
        lb = b.dump_logits(0, rb.n_calls)
    finally:
        b.close()
    assert ra.n_calls == rb.n_calls and (ra.qoff == rb.qoff).all()
    chh, cpg = ctx == 2, ctx == 0
    assert chh.sum() > 1000 and cpg.sum() > 100
    assert (la[chh].view(np.uint32) == lb[chh].view(np.uint32)).all() and (ra.ml[chh] == rb.ml[chh]).all()
    assert np.abs(la[cpg] - lb[cpg]).max() < 1e-4 and np.abs(ra.ml[cpg].astype(int) - rb.ml[cpg].astype(int)).max() <= 1


def test_chain_kernel_and_op_by_op_compact_ops_agree(lib_built, tmp_path):
    """The compact chain of a site tile runs in one kernel with its maps in tensor memory (site_chain.cuh); HM_NO_CHAIN=1 launches
    the same ops one by one through HBM.  Same split-precision products, different fp32 accumulation order only: logits agree to
    float noise, site lists exactly, ML bytes within the +-1 of a truncation boundary."""
    import subprocess
    import sys

    code = (
        "import numpy as np, sys\n"
        "from hifimeth_b200 import engine as hme, synth\n"
        "batch, _ = synth.make_reads(5, (1500, 4000), seed=4242, flag_rev_every=2)\n"
        "e = hme.Engine(max_reads=16, max_bases=1 << 16, keep_debug=True)\n"
        "r = e.call(batch)\n"
        "np.savez(sys.argv[1], qoff=r.qoff, ml=r.ml, logits=e.dump_logits(0, r.n_calls), launches=e.timing(0).kernel_launches)\n"
        "e.close()\n"
    )
    out = {}
    for name, env in (("chain", {}), ("ops", {"HM_NO_CHAIN": "1"})):
        path = tmp_path / f"{name}.npz"
        full = dict(os.environ, **env)
        full.pop("HM_NO_CHAIN", None) if not env else None
        subprocess.run([sys.executable, "-c", code, str(path)], check=True, env=full, cwd=str(ROOT), timeout=300)
        out[name] = np.load(path)
    a, b = out["chain"], out["ops"]
    assert (a["qoff"] == b["qoff"]).all() and len(a["qoff"]) > 3000
    assert int(a["launches"]) < int(b["launches"])  # one chain launch per context instead of 17 compact ops
    assert np.abs(a["logits"] - b["logits"]).max() < 2e-4
    assert np.abs(a["ml"].astype(int) - b["ml"].astype(int)).max() <= 1


def test_repeated_engines_and_pipelined_slots_are_deterministic(lib_built):
    """Engines created and destroyed in a row, two staging slots in flight at a time (the way the `call` driver and bench.py's
    e2e loop use them): every batch succeeds and every run returns the same bytes.  tools/stress_engine.py is the long form
    (200 engines x 4 full-size batches on one B200: no failure, one digest)."""
    import hashlib

    batch, _ = synth.make_reads(60, (4000, 15000), seed=99, flag_rev_every=4)
    ref = None
    for _ in range(6):
        eng = hme.Engine(n_slots=2, max_reads=64, max_bases=batch.n_bases + 1024)
        try:
            digests = []
            for i in range(4):
                n = eng.stage(i & 1, batch)
                eng.submit(i & 1, n)
                if i:
                    r = eng.collect((i & 1) ^ 1)
                    digests.append(hashlib.sha1(r.ml.tobytes() + r.qoff.tobytes() + r.call_off.tobytes()).hexdigest())
            r = eng.collect(1)
            digests.append(hashlib.sha1(r.ml.tobytes() + r.qoff.tobytes() + r.call_off.tobytes()).hexdigest())
        finally:
            eng.close()
        ref = ref or digests[0]
        assert digests == [ref] * 4
