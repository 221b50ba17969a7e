"""The raw-DEFLATE block codec of the BGZF reader / writer (hifimeth_b200/csrc/fast_deflate.cpp) against zlib, which is what htslib --
the reference's reader and writer, src/corelib/sam_batch.hpp:38-54, src/app/hifimeth/mod_main.cpp:353-362 -- uses: everything
hm_deflate_block writes must inflate to the input with zlib, and hm_inflate_block must read whatever zlib writes at any level and
strategy (stored, fixed and dynamic blocks, long distances), and reject damaged streams without touching memory outside its buffers."""
import ctypes as C
import zlib

import numpy as np
import pytest

from hifimeth_b200 import engine as E


@pytest.fixture(scope="module")
def lib(lib_built):
    return E.load_library()


def deflate_block(lib, data: bytes) -> bytes:
    out = C.create_string_buffer(len(data) + 64)
    n = lib.hm_deflate_block(data, len(data), out, len(data) + 64)
    assert n > 0
    return out.raw[:n]


def inflate_block(lib, comp: bytes, n_out: int):
    """(ok, bytes) with guard bytes on both sides of the output checked."""
    guard = 64
    buf = (C.c_uint8 * (n_out + 2 * guard))(*([0xA5] * (n_out + 2 * guard)))
    ok = lib.hm_inflate_block(comp, len(comp), C.byref(buf, guard), n_out)
    raw = bytes(buf)
    assert raw[:guard] == b"\xa5" * guard and raw[guard + n_out:] == b"\xa5" * guard, "wrote outside the output buffer"
    return bool(ok), raw[guard:guard + n_out]


def payloads():
    rng = np.random.default_rng(7)
    yield "empty", b""
    yield "one byte", b"x"
    yield "two bytes", b"ab"
    yield "three equal", b"aaa"
    yield "four equal", b"aaaa"
    yield "run 258", b"q" * 259
    yield "run 259", b"q" * 260
    yield "run 600", b"z" * 600
    yield "all zero 65535", bytes(65535)
    yield "all zero 0xff00", bytes(0xff00)
    yield "random 0xff00", rng.integers(0, 256, 0xff00, dtype=np.uint8).tobytes()
    yield "random 65535", rng.integers(0, 256, 65535, dtype=np.uint8).tobytes()
    yield "two symbols", rng.integers(0, 2, 40000, dtype=np.uint8).tobytes()
    yield "nibbles", rng.integers(0, 16, 50000, dtype=np.uint8).tobytes()
    # a skewed distribution deep enough that an unlimited Huffman code would exceed 15 bits (Fibonacci-like frequencies)
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    deep = np.concatenate([np.full(f, i, dtype=np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(deep)
    yield "deep code", deep[:65535].tobytes()
    # kinetics-like: skewed codes, runs of a constant quality, packed bases
    kin = np.minimum(rng.geometric(0.08, 30000), 255).astype(np.uint8).tobytes()
    yield "record-like", kin + b"\x28" * 15000 + rng.integers(0, 256, 7500, dtype=np.uint8).tobytes() + b"MM:Z:C+m?,1,2,3;" * 100
    runs = np.repeat(rng.integers(0, 5, 3000, dtype=np.uint8), rng.integers(1, 40, 3000))
    yield "many runs", runs[:65000].tobytes()
    yield "text", (b"@PG\tID:hifimeth\tPN:hifimeth\tVN:1.0\n" * 1500)[:60000]
    for n in (1, 2, 3, 5, 17, 255, 256, 257, 1000):
        yield f"random {n}", rng.integers(0, 256, n, dtype=np.uint8).tobytes()


@pytest.mark.parametrize("name,data", list(payloads()), ids=[n for n, _ in payloads()])
def test_deflate_block_is_read_by_zlib_and_by_the_own_inflater(lib, name, data):
    comp = deflate_block(lib, data)
    assert len(comp) <= len(data) + 64
    d = zlib.decompressobj(-15)
    assert d.decompress(comp) + d.flush() == data
    assert d.eof and d.unused_data == b""
    ok, back = inflate_block(lib, comp, len(data))
    assert ok and back == data


def test_deflate_block_is_not_larger_than_zlib_rle_by_much(lib):
    rng = np.random.default_rng(3)
    kin = np.minimum(rng.geometric(0.08, 40000), 255).astype(np.uint8).tobytes() + b"\x28" * 20000
    c = zlib.compressobj(1, zlib.DEFLATED, -15, 8, zlib.Z_RLE)
    ref = c.compress(kin) + c.flush()
    assert len(deflate_block(lib, kin)) <= len(ref) * 1.01 + 16


@pytest.mark.parametrize("name,data", list(payloads()), ids=[n for n, _ in payloads()])
def test_inflate_block_reads_every_zlib_setting(lib, name, data):
    for level in (0, 1, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED):
            c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            comp = c.compress(data) + c.flush()
            ok, back = inflate_block(lib, comp, len(data))
            assert ok and back == data, (level, strategy)


def test_inflate_block_reads_multi_block_streams_with_sync_flushes(lib):
    rng = np.random.default_rng(11)
    parts = [rng.integers(0, 7, 9000, dtype=np.uint8).tobytes(), b"", bytes(5000), rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(), b"abcabcabc" * 2000]
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = b""
    for i, p in enumerate(parts):
        comp += c.compress(p) + c.flush(zlib.Z_FULL_FLUSH if i % 2 else zlib.Z_SYNC_FLUSH)  # empty stored blocks in between
    comp += c.flush()
    data = b"".join(parts)
    ok, back = inflate_block(lib, comp, len(data))
    assert ok and back == data


def test_inflate_block_rejects_wrong_sizes_truncation_and_garbage(lib):
    rng = np.random.default_rng(5)
    data = (np.minimum(rng.geometric(0.1, 30000), 255).astype(np.uint8).tobytes() + b"abcdefgh" * 500)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    assert inflate_block(lib, comp, len(data))[0]
    assert not inflate_block(lib, comp, len(data) - 1)[0]      # stream longer than the block says
    assert not inflate_block(lib, comp, len(data) + 1)[0]      # stream shorter
    for cut in (0, 1, 2, 5, len(comp) // 2, len(comp) - 1):
        assert not inflate_block(lib, comp[:cut], len(data))[0]
    assert not inflate_block(lib, b"\x07" + comp[1:], len(data))[0]  # block type 3
    # a stored block whose LEN / NLEN do not match, one that is longer than the input
    assert not inflate_block(lib, b"\x01\x05\x00\x00\x00hello", 5)[0]
    assert not inflate_block(lib, b"\x01\x05\x00\xfa\xffhell", 5)[0]
    assert inflate_block(lib, b"\x01\x05\x00\xfa\xffhello", 5) == (True, b"hello")
    # a match that reaches in front of the output: fixed block, length 3 distance 1 as the first symbol
    bits = "1" + "01" + "0000001" + "00000" + "0000000"   # BFINAL, BTYPE = 01 (LSB first: 1, then 10), length code 257, distance code 0, EOB
    v = int(bits[::-1], 2)
    assert not inflate_block(lib, v.to_bytes((len(bits) + 7) // 8, "little"), 3)[0]


def test_inflate_block_survives_bit_flips(lib):
    """Damaged streams: any answer is fine when the bytes happen to decode, but never a write outside the buffer (guards in
    inflate_block) and never a crash; a stream that zlib reads to the same size must give zlib's bytes."""
    rng = np.random.default_rng(9)
    data = np.minimum(rng.geometric(0.05, 20000), 255).astype(np.uint8).tobytes() + b"\x28" * 3000 + b"ACGT" * 700
    for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_RLE), (9, zlib.Z_FIXED)):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        comp = bytearray(c.compress(data) + c.flush())
        for _ in range(300):
            bad = bytearray(comp)
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
            ok, back = inflate_block(lib, bytes(bad), len(data))
            try:
                d = zlib.decompressobj(-15)
                ref = d.decompress(bytes(bad)) + d.flush()
                ref_ok = d.eof and len(ref) == len(data)
            except zlib.error:
                ref_ok = False
            if ok:
                assert ref_ok and back == ref


def test_crc32_matches_zlib_for_every_length_alignment_and_start_value(lib):
    rng = np.random.default_rng(21)
    data = rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    buf = C.create_string_buffer(data, len(data))
    base = C.addressof(buf)
    for n in list(range(0, 300)) + [511, 512, 513, 4095, 4096, 4097, 0xff00, 65535, 70000 - 3]:
        for off in (0, 1, 3):
            for start in (0, 0x12345678, 0xffffffff):
                assert lib.hm_crc32_bytes(start, base + off, n) == zlib.crc32(data[off:off + n], start), (n, off, start)
    # continuing a CRC piecewise is the same as one call
    c = 0
    for lo, hi in ((0, 1000), (1000, 1017), (1017, 40000), (40000, 70000)):
        c = lib.hm_crc32_bytes(c, base + lo, hi - lo)
    assert c == zlib.crc32(data)
