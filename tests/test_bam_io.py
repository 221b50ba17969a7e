"""Row N2: the block-parallel BGZF/BAM codec of the `call` driver, on CPU (no device code involved)."""
import struct

import numpy as np
import pytest

from hifimeth_b200 import engine as hme
from hifimeth_b200 import synth


def test_bam_round_trip_parallel_codec(lib_built, tmp_path):
    _, reads = synth.make_reads(40, (300, 9000), seed=5, flag_rev_every=3)
    bodies = [synth.record_body(r, kinetics_as_u16=(i % 5 == 0)) for i, r in enumerate(reads)]
    src, dst = tmp_path / "in.bam", tmp_path / "out.bam"
    synth.write_bam(src, bodies, level=1, block=30000)  # blocks smaller than a record: records straddle blocks
    lib = hme.load_library()
    for threads, level in ((1, 1), (4, 6)):
        n = lib.hm_bam_copy(str(src).encode(), str(dst).encode(), threads, level)
        assert n == len(bodies)
        text, refs, got = synth.read_bam(dst)
        text0, refs0, want = synth.read_bam(src)
        assert text == text0 and refs == refs0 == struct.pack("<i", 0)
        assert got == want == bodies
        raw = dst.read_bytes()
        assert raw[:4] == bytes([31, 139, 8, 4]) and raw[12:14] == b"BC"          # BGZF member with the BC extra field
        assert raw[-28:] == bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])  # EOF marker


def test_bam_records_straddling_slabs(lib_built, tmp_path, monkeypatch):
    """The reader hands out records without copying them; one that straddles two (or, with tiny slabs, many) inflated slabs is
    gathered into a side buffer.  HM_BGZF_SLAB shrinks the slabs so that most records straddle."""
    _, reads = synth.make_reads(25, (200, 7000), seed=11)
    bodies = [synth.record_body(r) for r in reads]
    src, dst = tmp_path / "in.bam", tmp_path / "out.bam"
    synth.write_bam(src, bodies, level=1, block=5000)
    lib = hme.load_library()
    for slab in ("1024", "9000", "70000"):
        monkeypatch.setenv("HM_BGZF_SLAB", slab)
        assert lib.hm_bam_copy(str(src).encode(), str(dst).encode(), 3, 1) == len(bodies)
        assert synth.read_bam(dst)[2] == bodies
    # a file cut in the middle of a record is an error, not a short read
    raw = src.read_bytes()
    cut = tmp_path / "cut.bam"
    cut.write_bytes(raw[:len(raw) // 2])
    assert lib.hm_bam_copy(str(cut).encode(), str(dst).encode(), 2, 1) < 0


def test_bam_copy_rejects_garbage(lib_built, tmp_path):
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"this is not a BAM file at all" * 10)
    assert hme.load_library().hm_bam_copy(str(bad).encode(), str(tmp_path / "o.bam").encode(), 2, 1) < 0
    assert hme.load_library().hm_bam_copy(str(tmp_path / "missing.bam").encode(), str(tmp_path / "o.bam").encode(), 2, 1) < 0


def test_call_cli_usage_errors(lib_built):
    import subprocess

    exe = hme.PKG / "bin" / "hifimeth-b200"
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode != 0 and "call [OPTIONS] BAM MOD-BAM" in r.stderr
    r = subprocess.run([str(exe), "call", "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "-c <string>" in r.stderr and "-k" in r.stderr
    r = subprocess.run([str(exe), "call", "-c", "cpg,xyz", "a.bam", "b.bam"], capture_output=True, text=True)
    assert r.returncode != 0
    r = subprocess.run([str(exe), "call", "only-one-path.bam"], capture_output=True, text=True)
    assert r.returncode != 0
