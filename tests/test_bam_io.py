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


def test_bam_reader_survives_corruption(lib_built, tmp_path):
    """Fuzz of the BGZF/BAM reader: truncations and random byte damage of a valid file (headers, BSIZE / ISIZE / CRC fields,
    deflate streams, record lengths after re-compression) must end in an error code or a clean copy -- never in a crash, a hang
    or an unbounded allocation.  Runs in a child process so that a crash is a test failure, not the end of the session."""
    import subprocess
    import sys
    import zlib

    _, reads = synth.make_reads(12, (300, 4000), seed=99)
    bodies = [synth.record_body(r) for r in reads]
    good = tmp_path / "good.bam"
    synth.write_bam(good, bodies, level=1, block=20000)
    raw = bytearray(good.read_bytes())
    rng = np.random.default_rng(7)
    cases = []
    for cut in (0, 10, 17, 28, 100, len(raw) // 3, len(raw) - 29, len(raw) - 1):
        cases.append(bytes(raw[:cut]))
    for _ in range(60):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        cases.append(bytes(b))
    # damage INSIDE the BAM stream (valid BGZF around it): record lengths, header lengths
    text, refs, _ = synth.read_bam(good)
    import gzip
    stream = bytearray(gzip.decompress(bytes(raw)))
    for pos, val in ((4, 0xffffffff), (4, 0x7fffffff), (8 + len(text), 0x10000000), (8 + len(text) + 4, 0xfffffff0),
                     (8 + len(text) + 4, 5), (8 + len(text) + 4 + 16, 0x7fffffff)):
        s2 = bytearray(stream)
        s2[pos:pos + 4] = int(val).to_bytes(4, "little")
        out = bytearray()
        for off in range(0, len(s2), 0xff00):
            chunk = bytes(s2[off:off + 0xff00])
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            out += bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0]) + (len(comp) + 25).to_bytes(2, "little")
            out += comp + (zlib.crc32(chunk) & 0xffffffff).to_bytes(4, "little") + len(chunk).to_bytes(4, "little")
        cases.append(bytes(out))
    for i, c in enumerate(cases):
        (tmp_path / f"case{i}.bam").write_bytes(c)
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from hifimeth_b200 import engine as hme\n"
        "lib = hme.load_library()\n"
        "res = []\n"
        "for i in range(%d):\n"
        "    res.append(lib.hm_bam_copy((%r + '/case%%d.bam' %% i).encode(), (%r + '/out.bam').encode(), 3, 1))\n"
        "print(' '.join(map(str, res)))\n" % (str(hme.PKG.parent), len(cases), str(tmp_path), str(tmp_path)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    res = [int(x) for x in r.stdout.split()]
    assert len(res) == len(cases)
    assert all(x < 0 for x in res[:8])             # every truncation is an error
    assert all(x < 0 or x <= len(bodies) + 2 for x in res)
    assert sum(x < 0 for x in res) > len(res) // 2  # most random damage is caught (CRC / structure)


def test_bgzf_block_with_oversized_xlen_is_rejected(lib_built, tmp_path):
    """A crafted block whose XLEN is larger than BSIZE allows (the inflate length would wrap around and zlib would read far past the
    buffer) must be refused as corrupt -- ADVICE r1."""
    import zlib

    def block(payload: bytes, xlen_extra: int, bsize_delta: int = 0) -> bytes:
        co = zlib.compressobj(1, zlib.DEFLATED, -15)
        comp = co.compress(payload) + co.flush()
        extra = b"BC" + (2).to_bytes(2, "little") + b"\0\0" + b"\0" * xlen_extra
        total = 12 + len(extra) + len(comp) + 8
        extra = b"BC" + (2).to_bytes(2, "little") + (total - 1 + bsize_delta).to_bytes(2, "little") + b"\0" * xlen_extra
        return bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255]) + len(extra).to_bytes(2, "little") + extra + comp + \
            (zlib.crc32(payload) & 0xffffffff).to_bytes(4, "little") + len(payload).to_bytes(4, "little")

    lib = hme.load_library()
    out = str(tmp_path / "o.bam").encode()
    # XLEN = 200 claimed, but BSIZE says the whole block is 40 bytes
    evil = bytearray(block(b"BAM\x01" + bytes(8), 0))
    evil[10:12] = (200).to_bytes(2, "little")
    evil += bytes(400)
    bad = tmp_path / "xlen.bam"
    for bs in (27, 40, 100, 219):
        e2 = bytearray(evil)
        e2[16:18] = (bs - 1).to_bytes(2, "little")
        bad.write_bytes(bytes(e2))
        assert lib.hm_bam_copy(str(bad).encode(), out, 2, 1) < 0
    # sanity: padding inside a consistent extra field is legal BGZF and still reads
    _, reads = synth.make_reads(2, 400, seed=1)
    bodies = [synth.record_body(r) for r in reads]
    good = tmp_path / "g.bam"
    synth.write_bam(good, bodies, level=1)
    import gzip
    stream = gzip.decompress(good.read_bytes())
    padded = tmp_path / "padded.bam"
    padded.write_bytes(block(stream, 6) + block(b"", 0))
    assert lib.hm_bam_copy(str(padded).encode(), out, 2, 1) == len(bodies)


def test_reader_and_writer_pipelines_keep_file_order(lib_built, tmp_path, monkeypatch):
    """The reader is three stages deep (I/O thread, two inflate drivers working on different slabs at the same time, ordered
    hand-over) and the writer deflates whole chunks in a background thread: with slabs of a few blocks, hundreds of slabs are in
    flight over a file, and the records must still come out in file order, none lost or doubled, every time."""
    _, reads = synth.make_reads(300, (150, 900), seed=23)
    for i, r in enumerate(reads):
        r["name"] = f"order/{i}/ccs"
    bodies = [synth.record_body(r) for r in reads]
    src, dst = tmp_path / "in.bam", tmp_path / "out.bam"
    synth.write_bam(src, bodies, level=1, block=1500)
    lib = hme.load_library()
    monkeypatch.setenv("HM_BGZF_SLAB", "2048")
    for rep in range(4):
        assert lib.hm_bam_copy(str(src).encode(), str(dst).encode(), 8, 1 + rep % 2) == len(bodies)
        assert synth.read_bam(dst)[2] == bodies


def test_reader_paths_and_piecewise_writer_agree(lib_built, tmp_path, monkeypatch):
    """The mapped and the fread input paths, the read-only mode, and the writer fed with finished pieces of the stream (what `call`
    hands over: large and small pieces, each closing its last BGZF block short) all give the same records."""
    _, reads = synth.make_reads(60, (300, 12000), seed=17, flag_rev_every=4)
    bodies = [synth.record_body(r) for r in reads]
    src, dst = tmp_path / "in.bam", tmp_path / "out.bam"
    synth.write_bam(src, bodies, level=6, block=40000)
    lib = hme.load_library()
    for no_mmap in ("0", "1"):
        monkeypatch.setenv("HM_NO_MMAP", no_mmap)
        assert lib.hm_bam_copy(str(src).encode(), None, 3, -1) == len(bodies)          # read only
        for level in (101, 106, 1):
            assert lib.hm_bam_copy(str(src).encode(), str(dst).encode(), 3, level) == len(bodies)
            text, refs, got = synth.read_bam(dst)
            assert got == bodies and text == synth.read_bam(src)[0]
            assert dst.read_bytes()[-28:-12] == bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0])
    # a mapped file that ends inside a block is an error on this path too
    cut = tmp_path / "cut.bam"
    cut.write_bytes(src.read_bytes()[:-40])
    monkeypatch.setenv("HM_NO_MMAP", "0")
    assert lib.hm_bam_copy(str(cut).encode(), None, 2, -1) < 0


def test_zlib_only_switch_gives_the_same_records(lib_built, tmp_path):
    """HM_ZLIB_ONLY=1 (read once per process, hence the subprocess): every block through zlib -- the path that also arbitrates
    blocks the built-in inflater rejects.  Same records, and a level-1 file any inflater reads."""
    import os
    import subprocess
    import sys

    _, reads = synth.make_reads(30, (300, 9000), seed=23)
    bodies = [synth.record_body(r) for r in reads]
    src, dst = tmp_path / "in.bam", tmp_path / "out.bam"
    synth.write_bam(src, bodies, level=6)
    code = ("import sys; from hifimeth_b200 import engine as e; L = e.load_library(); "
            f"sys.exit(0 if L.hm_bam_copy({str(src)!r}.encode(), {str(dst)!r}.encode(), 3, 1) == {len(bodies)} else 1)")
    env = dict(os.environ, HM_ZLIB_ONLY="1", PYTHONPATH=str(hme.PKG.parent))
    assert subprocess.run([sys.executable, "-c", code], env=env).returncode == 0
    assert synth.read_bam(dst)[2] == bodies
