"""CPU tests of the C-ABI library: it builds for sm_100a, loads, exports every symbol include/hm_engine.h
declares, its host helpers agree with the oracle / golden vectors, and engine creation FAILS LOUDLY without a
GPU (no CPU fallback).  No device compute is exercised here."""
import ctypes as C
import re

import numpy as np
import pytest

from hifimeth_b200 import engine as hme
from hifimeth_b200 import synth
from oracle import hmoracle
from oracle.make_golden import golden_bodies, golden_reads

from conftest import ROOT


def _declared_functions():
    text = (ROOT / "include" / "hm_engine.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    lib = hme.load_library()
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hm_engine.h but not exported"
    assert sorted(hme.ABI_SYMBOLS) == declared
    assert b"sm_100a" in lib.hm_version()


def test_library_is_sm100a_only(lib_built):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", str(hme.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_engine_create_fails_loudly_without_gpu(lib_built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hme.HmError) as ei:
        hme.Engine(max_reads=4, max_bases=1 << 16)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_codev1_helpers(lib_built):
    lib = hme.load_library()
    lut = synth.codev1_decode_table()
    for c in range(256):
        assert lib.hm_codev1_decode(c) == lut[c]
    O = hmoracle.oracle()
    for s in list(range(0, 1100)) + [5000, 65535]:
        assert lib.hm_codev1_encode(s) == O.lib.hmo_codev1_encode(s)


def test_pack_record_matches_soa_and_acceptance_rules(lib_built):
    reads = golden_reads()
    bodies = golden_bodies(reads)
    packed = hme.pack_records_host(bodies, min_read_len=1000)
    want = synth.soa_from_reads(reads, min_read_len=1000)
    assert packed.n_reads == want.n_reads
    assert (packed.base_off == want.base_off).all() and (packed.seq_off == want.seq_off).all()
    assert (packed.seq4 == want.seq4).all() and (packed.flag == want.flag).all()
    # read 5 lacks rp, read 6 is shorter than -l: both passed through (src/app/hifimeth/mod_main.cpp:189-196)
    assert list(packed.valid) == [1, 1, 1, 1, 1, 0, 0, 1]
    assert (packed.valid == want.valid).all()
    for r in range(want.n_reads):
        if not want.valid[r]:
            continue
        a, b = int(want.base_off[r]), int(want.base_off[r + 1])
        for k in ("fi", "fp", "ri", "rp"):  # read 1 carries B:S raw frames that must re-encode to the same codes
            assert (getattr(packed, k)[a:b] == getattr(want, k)[a:b]).all(), (r, k)


def test_pack_record_rejects_overflow_and_garbage(lib_built):
    reads = golden_reads()
    body = golden_bodies(reads)[0]
    with pytest.raises(hme.HmError):
        hme.pack_records_host([body, body], max_bases=len(reads[0]["seq"]) + 10)
    lib = hme.load_library()
    b = hme.hm_read_batch()
    n = C.c_uint32(0)
    junk = np.zeros(8, np.uint8)
    assert lib.hm_pack_record(C.byref(b), C.byref(n), junk.ctypes.data_as(C.POINTER(C.c_uint8)), 8, 1000) != 0


def test_pack_records_batched_equals_one_by_one(lib_built):
    """hm_pack_records (the `call` driver's parallel packer) fills the staging arrays exactly like repeated hm_pack_record;
    malformed and over-long records get index -1 and take no space; a batch that does not fit is refused."""
    _, reads = synth.make_reads(23, (300, 5000), seed=77, flag_rev_every=3, short_every=5)
    bodies = golden_bodies(golden_reads()) + [synth.record_body(r, kinetics_as_u16=(i % 4 == 1)) for i, r in enumerate(reads)]
    one = hme.pack_records_host(bodies, min_read_len=1000)
    for threads in (1, 5):
        many = hme.pack_records_host(bodies, min_read_len=1000, threads=threads)
        assert many.n_reads == one.n_reads == len(bodies)
        for k in ("base_off", "seq_off", "seq4", "flag", "valid", "fi", "fp", "ri", "rp"):
            assert (getattr(many, k) == getattr(one, k)).all(), k
    junk = bytes(40)  # l_seq 0 with a name length of 0: parses as an empty record; a truncated one does not
    with_bad = [bodies[0], bodies[1][:50], bodies[2]]
    got = hme.pack_records_host(with_bad, min_read_len=1000, threads=2)
    assert got.n_reads == 2 and (got.fi == hme.pack_records_host([bodies[0], bodies[2]]).fi).all()
    lens = [len(r["seq"]) for r in golden_reads()]
    i_long, i_short = int(np.argmax(lens)), int(np.argmin(lens))
    gb = golden_bodies(golden_reads())
    got = hme.pack_records_host([gb[i_long], gb[i_short]], min_read_len=1000, threads=2, max_bases=lens[i_long] - 1)
    assert got.n_reads == 1 and got.n_bases == lens[i_short]  # the long record can never be staged: left out, not an error
    with pytest.raises(hme.HmError):
        hme.pack_records_host([bodies[0], bodies[0]], threads=2, max_bases=len(golden_reads()[0]["seq"]) + 10)
    del junk


def test_build_mod_record_vs_golden(lib_built, golden):
    for i in range(int(golden["n_reads"])):
        body = golden[f"body{i}"].tobytes()
        if not golden[f"ok{i}"]:
            assert hme.build_mod_record(body, False, [], [], [], []) == golden[f"mod{i}"].tobytes()
            continue
        fq, rq, fml, rml = (golden[f"{k}{i}"] for k in ("fq", "rq", "fml", "rml"))
        assert hme.build_mod_record(body, False, fq, fml, rq, rml) == golden[f"mod{i}"].tobytes()
        assert hme.build_mod_record(body, True, fq, fml, rq, rml) == golden[f"modkeep{i}"].tobytes()


def test_build_mod_record_rejects_bad_calls(lib_built, golden):
    body = golden["body0"].tobytes()
    fq = golden["fq0"]
    bad = fq.copy()
    bad[1], bad[0] = fq[0], fq[1]  # not ascending: the reference aborts (build_mod_bam.cpp:138), we return an error
    with pytest.raises(hme.HmError):
        hme.build_mod_record(body, False, bad, np.zeros(len(bad), np.uint8), [], [])
    with pytest.raises(hme.HmError):
        not_c = int(np.nonzero(golden["fwd0"] != 1)[0][0])  # a call on a base that is not C (build_mod_bam.cpp:139-140)
        hme.build_mod_record(body, False, [not_c], [1], [], [])


def test_build_mod_record_from_mm_text_vs_golden(lib_built, golden):
    """hm_build_mod_record_mm (row N1's host half) against the records the REFERENCE's build_mod_bam.cpp produced: the MM text
    is cut out of the reference record itself, so only the assembly (tag stripping, MM/ML/MN layout) is under test here."""
    n = 0
    for i in range(int(golden["n_reads"])):
        body = golden[f"body{i}"].tobytes()
        if not golden[f"ok{i}"]:
            assert hme.build_mod_record_mm(body, False, b"", b"", b"", 0, 0) == golden[f"mod{i}"].tobytes()
            continue
        fq, rq, fml, rml = (golden[f"{k}{i}"] for k in ("fq", "rq", "fml", "rml"))
        want = golden[f"mod{i}"].tobytes()
        if len(fq) + len(rq) == 0:
            assert hme.build_mod_record_mm(body, False, b"", b"", b"", 0, 0) == want
            continue
        mm = want[want.index(b"MMZC+m") + 3:]
        mm = mm[:mm.index(b"\0")]
        fwd_part, rev_part, tail = mm.split(b";")
        assert fwd_part.startswith(b"C+m") and rev_part.startswith(b"G-m") and tail == b""
        ml = np.concatenate([fml, rml]).astype(np.uint8)
        for keep, key in ((False, "mod"), (True, "modkeep")):
            got = hme.build_mod_record_mm(body, keep, np.frombuffer(fwd_part[3:], np.uint8), np.frombuffer(rev_part[3:], np.uint8), ml, len(fq), len(rq))
            assert got == golden[f"{key}{i}"].tobytes(), (i, keep)
        n += 1
    assert n >= 4


def test_parse_mod_record_matches_reference_parser(lib_built, golden):
    """Row N4: hm_parse_mod_record on the golden mod records (produced by the reference's build_one_mod_bam; flag 0x10 reads
    included) returns the calls they were built from, and agrees with the reference's own extract_bam_base_mods
    (oracle/_ref, src/corelib/bam_mod_parser.cpp:231-286) when that library is present."""
    try:
        R = hmoracle.ref()
    except Exception:
        R = None  # /root/reference is absent on the GPU box; the golden vectors still pin the result
    n_checked = 0
    for i in range(int(golden["n_reads"])):
        rec = golden[f"mod{i}"].tobytes()
        q, s, p, codes = hme.parse_mod_record(rec)
        if not golden[f"ok{i}"]:
            assert len(q) == 0
            continue
        fq, rq, fml, rml = (golden[f"{k}{i}"] for k in ("fq", "rq", "fml", "rml"))
        assert (q == np.concatenate([fq, rq])).all() and (p == np.concatenate([fml, rml])).all()
        assert (s[:len(fq)] == 0).all() and (s[len(fq):] == 1).all() and codes == b"m" * len(q)
        if R is not None and len(q):
            qq, ss, pp = R.parse_mods(rec, len(q) + 1)
            assert (qq == q).all() and (ss == s).all() and (pp == p).all()
        n_checked += 1
    assert n_checked >= 5


def test_parse_mod_record_general_series_and_errors(lib_built):
    """MM forms the reference parser accepts beyond what `call` writes (ChEBI numbers, several codes, '.'/'?' flags, other ML
    integer types) and the malformed ones it aborts on (HM_ERR_FORMAT here)."""
    import struct

    _, reads = synth.make_reads(1, 64, seed=9, min_read_len=10)
    rd = reads[0]
    rd["seq"][:] = np.array([1, 0, 1, 2, 3, 1, 1, 2] * 8, np.uint8)  # C A C G T C C G ...
    base = synth.record_body(dict(rd, fi=None, ri=None, fp=None, rp=None))

    def with_tags(mm: bytes, ml: bytes):
        return base + b"MMZ" + mm + b"\0" + ml

    ml3 = b"MLBC" + struct.pack("<I", 3) + bytes([10, 20, 30])
    # C's are at 0, 2, 5, 6, 8, ...: skip 1 -> position 2, skip 0 -> 5, skip 1 -> 8
    q, s, p, codes = hme.parse_mod_record(with_tags(b"C+m?,1,0,1;", ml3))
    assert list(q) == [2, 5, 8] and list(s) == [0, 0, 0] and list(p) == [10, 20, 30] and codes == b"mmm"
    q, s, p, codes = hme.parse_mod_record(with_tags(b"C+27551,1,0,1;", ml3))
    assert list(q) == [2, 5, 8] and codes == b"mmm"
    ml4 = b"MLBS" + struct.pack("<I", 4) + struct.pack("<4H", 1, 2, 3, 255)
    q, s, p, codes = hme.parse_mod_record(with_tags(b"C+mh,0,0;G-m;", ml4))  # two codes per position
    assert list(q) == [0, 0, 2, 2] and codes == b"mhmh" and list(p) == [1, 2, 3, 255]
    q, s, p, codes = hme.parse_mod_record(with_tags(b"C+m,0;G-m,1;", b"MLBC" + struct.pack("<I", 2) + bytes([7, 9])))
    assert list(q) == [0, 7] and list(s) == [0, 1]  # G's at 3, 7
    assert len(hme.parse_mod_record(base)[0]) == 0  # no tags: nothing
    for mm, ml in ((b"C+m,1,0,1", ml3),            # no terminating ';'
                   (b"X+m,1;", ml3),               # unknown base
                   (b"C*m,1;", ml3),               # unknown strand
                   (b"C+a,1;", ml3),               # code does not fit the base
                   (b"C+m,1,0,1,0;", ml3),         # more positions than ML values
                   (b"C+m,99;", ml3),              # skip count runs past the read
                   (b"C+m,1;", b"MLBS" + struct.pack("<I", 1) + struct.pack("<H", 256))):  # probability out of range
        with pytest.raises(hme.HmError):
            hme.parse_mod_record(with_tags(mm, ml))


def test_ml_threshold_rule_matches_oracle(lib_built):
    """Row N3: hm_ml_threshold against the oracle's restatement of s_resolve_scaled_prob_threshold (pileup.cpp:355-436) on
    bimodal, flat, sparse, narrow and empty histograms."""
    O = hmoracle.oracle()
    rng = np.random.default_rng(5)
    cases = []
    x = np.arange(256)
    bimodal = (40000 * np.exp(-x / 12.0) + 30000 * np.exp(-(255 - x) / 9.0) + 15).astype(np.uint64)
    cases.append(bimodal)
    cases.append(np.full(256, 100, np.uint64))                       # flat: first minimum of the range wins
    cases.append(np.zeros(256, np.uint64))                           # empty -> 128
    sparse = np.zeros(256, np.uint64); sparse[100:140] = 5000        # range narrower than 50 bins -> 128
    cases.append(sparse)
    few = np.zeros(256, np.uint64); few[30:200] = 12                 # wide enough but < 10000 samples -> 128
    cases.append(few)
    edge = bimodal.copy(); edge[:60] = 3; edge[200:] = 9             # outer bins below 10 shrink the range
    cases.append(edge)
    for _ in range(40):
        cases.append(rng.integers(0, 400, 256).astype(np.uint64) * rng.integers(0, 2, 256).astype(np.uint64))
        cases.append((rng.integers(0, 3000, 256) + 10).astype(np.uint64))
    seen = set()
    for b in cases:
        got, want = hme.ml_threshold(b), O.ml_threshold(b)
        assert got == want, (got, want)
        seen.add(got[0])
    assert 128 in seen and len(seen) > 5
    assert hme.ml_threshold(bimodal)[0] == int(np.argmin(bimodal[20:236])) + 20


def _threshold_cases():
    rng = np.random.default_rng(5)
    x = np.arange(256)
    bimodal = (40000 * np.exp(-x / 12.0) + 30000 * np.exp(-(255 - x) / 9.0) + 15).astype(np.uint64)
    cases = [bimodal, np.full(256, 100, np.uint64), np.zeros(256, np.uint64)]
    sparse = np.zeros(256, np.uint64); sparse[100:140] = 5000
    few = np.zeros(256, np.uint64); few[30:200] = 12
    edge = bimodal.copy(); edge[:60] = 3; edge[200:] = 9
    exact50 = np.zeros(256, np.uint64); exact50[60:110] = 300       # en - st == 50 exactly: the rule applies
    exact49 = np.zeros(256, np.uint64); exact49[60:109] = 300       # 49 bins: it does not
    sum9999 = np.zeros(256, np.uint64); sum9999[40:140] = 100; sum9999[40] = 99   # 9 999 samples -> 128
    sum10000 = np.zeros(256, np.uint64); sum10000[40:140] = 100                   # 10 000 samples -> first minimum
    outer = np.zeros(256, np.uint64); outer[:] = 50; outer[:20] = 10**7; outer[236:] = 10**7; outer[130] = 11
    ties = np.full(256, 77, np.uint64); ties[90] = 76; ties[150] = 76              # equal minima: the first wins
    huge = np.full(256, 2**40, np.uint64); huge[77] = 2**40 - 1
    cases += [sparse, few, edge, exact50, exact49, sum9999, sum10000, outer, ties, huge]
    for _ in range(60):
        cases.append(rng.integers(0, 400, 256).astype(np.uint64) * rng.integers(0, 2, 256).astype(np.uint64))
        cases.append((rng.integers(0, 3000, 256) + 10).astype(np.uint64))
        b = np.zeros(256, np.uint64)
        lo, hi = sorted(int(v) for v in rng.integers(0, 257, 2))
        b[lo:hi] = rng.integers(0, 25, hi - lo)
        b[lo:hi] *= np.uint64(rng.integers(1, 40))
        cases.append(b)
    return cases


def test_ml_threshold_pinned_to_reference_pileup(lib_built, capfd):
    """Row N3, pinned: hm_ml_threshold (and the oracle's restatement) against the REFERENCE'S OWN compiled
    s_resolve_scaled_prob_threshold (src/app/hifimeth/pileup.cpp:355-436; oracle/ref_pileup.cpp includes pileup.cpp into
    oracle/_ref) on degenerate, boundary and random histograms, in every context slot of the reference's signature."""
    R = hmoracle.ref()
    if not R.available or not R.has_pileup:
        pytest.skip("oracle/_ref was built without the pileup translation unit")
    O = hmoracle.oracle()
    cases = _threshold_cases()
    seen = set()
    for i in range(0, len(cases) - 2, 1):
        a, b, c = cases[i], cases[i + 1], cases[i + 2]
        want = R.pileup_thresholds(a, b, c)
        got = tuple(hme.ml_threshold(x)[0] for x in (a, b, c))
        assert got == want, (i, got, want)
        assert tuple(O.ml_threshold(x)[0] for x in (a, b, c)) == want
        seen.update(want)
    assert 128 in seen and len(seen) > 10
    capfd.readouterr()  # the reference prints its decision to stderr


def test_existing_mn_tag_keeps_its_width_like_bam_aux_update_int(lib_built):
    """A re-called BAM already carries MN (and MM / ML): build_one_mod_bam updates MN through bam_aux_update_int
    (src/corelib/build_mod_bam.cpp:240-247), which reuses the field in place and keeps its width when the value fits.  Checked
    against the reference's build_mod_bam.cpp compiled in oracle/_ref, for every integer type the old field may have."""
    import struct

    R = hmoracle.ref()
    if not R.available:
        pytest.skip("oracle/_ref not built")
    _, reads = synth.make_reads(1, 1200, seed=3)
    base = synth.record_body(reads[0])
    seq = np.asarray(reads[0]["seq"])  # the reference asserts that forward calls sit on a C and reverse calls on a G
    fq = np.nonzero(seq == 1)[0][[2, 40, 200]].astype(np.int32); fml = np.array([1, 128, 255], np.uint8)
    rq = np.nonzero(seq == 2)[0][[5, 250]].astype(np.int32); rml = np.array([7, 9], np.uint8)
    olds = [b"MNC" + struct.pack("<B", 7), b"MNc" + struct.pack("<b", 7), b"MNS" + struct.pack("<H", 99), b"MNs" + struct.pack("<h", 99),
            b"MNI" + struct.pack("<I", 123456), b"MNi" + struct.pack("<i", 123456)]
    for old in olds:
        for where in ("end", "middle"):
            body = base + old if where == "end" else base + old + b"zzZtail\0"
            want = R.build_mod_bam(body, False, fq, fml, rq, rml)
            got = hme.build_mod_record(body, False, fq, fml, rq, rml)
            assert got == want, (old, where)
            # the record still parses and MN holds l_seq
            assert b"MN" in got
    # width as a function of the value for a NEW tag: htslib compares with `<`
    for L, typ in ((254, b"C"), (255, b"S"), (65534, b"S"), (65535, b"I")):
        _, rd = synth.make_reads(1, L, seed=4)
        body = synth.record_body(rd[0])
        q = np.nonzero(np.asarray(rd[0]["seq"]) == 1)[0][:1].astype(np.int32); m = np.array([200], np.uint8)
        want = R.build_mod_bam(body, False, q, m, [], [])
        got = hme.build_mod_record(body, False, q, m, [], [])
        assert got == want and got[-(3 + {b"C": 1, b"S": 2, b"I": 4}[typ]):][:3] == b"MN" + typ, (L, typ)


def test_model_loader_onnx_and_torchscript(lib_built, tmp_path):
    """The weight loader reads both ONNX dialects and the TorchScript .pt exports (ZIP directory + the 24 stored constants):
    parameters equal the oracle's own parse of the .onnx files; CpG.pt / CHH.pt equal their .onnx (<= 2.4e-7 / bit-equal) and
    CHG.pt is the different checkpoint SURVEY.md s0.5 describes; damaged archives are errors, not crashes."""
    from oracle import cnn_oracle

    md = ROOT / "models"
    for name, k1 in (("CpG", 11), ("CHG", 11), ("CHH", 13)):
        w = cnn_oracle.CnnWeights(md / f"{name}.onnx")
        want = np.concatenate([np.ravel(t) for t in w.bn0] + [np.ravel(x) for cw, cb in w.convs for x in (cw, cb)] +
                              [np.ravel(w.fcs[0][0]), np.ravel(w.fcs[0][1]), np.ravel(w.fcs[1][0]), np.ravel(w.fcs[1][1])]).astype(np.float32)
        got, k = hme.model_weights(md / f"{name}.onnx")
        assert k == k1 and got.shape == want.shape and (got == want).all(), name
        pt, kp = hme.model_weights(md / f"{name}.pt")
        assert kp == k1 and pt.shape == want.shape
        d = float(np.abs(pt - want).max())
        if name == "CHH":
            assert d == 0.0
        elif name == "CpG":
            assert d < 1e-6
        else:
            assert d > 1e-2  # CHG.pt is another checkpoint
    raw = (md / "CpG.pt").read_bytes()
    for bad in (raw[:1000], raw[:-30], b"PK" + bytes(100), raw.replace(b"constants/7", b"constantz/7")):
        f = tmp_path / "bad.pt"
        f.write_bytes(bad)
        with pytest.raises(hme.HmError):
            hme.model_weights(f)
