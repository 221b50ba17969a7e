"""codec_bench.py -- single-thread MB/s and output size of the BGZF block codec (hm_deflate_block / hm_inflate_block,
hifimeth_b200/csrc/fast_deflate.cpp) against zlib on the payloads of a BAM file.   python tools/codec_bench.py in.bam [--blocks 600]"""
import argparse, ctypes as C, os, struct, sys, time, zlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hifimeth_b200 import engine as E

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bam")
    ap.add_argument("--blocks", type=int, default=600)
    a = ap.parse_args()
    L = E.load_library()
    raw = open(a.bam, "rb").read()
    off, comp_in = 0, []
    while off < len(raw) and len(comp_in) < a.blocks:
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        bs = struct.unpack_from("<H", raw, off + 16)[0] + 1
        comp_in.append(raw[off + 12 + xlen:off + bs - 8])
        off += bs
    t = time.time(); plain = [zlib.decompress(b, -15) for b in comp_in]; dz = time.time() - t
    tot = sum(map(len, plain))
    outs = [C.create_string_buffer(len(p)) for p in plain]
    t = time.time()
    for b, p, o in zip(comp_in, plain, outs):
        assert L.hm_inflate_block(b, len(b), o, len(p))
    do = time.time() - t
    assert all(o.raw == p for o, p in zip(outs, plain))
    print(f"inflate of the file's blocks: zlib {tot / dz / 1e6:7.1f} MB/s, own {tot / do / 1e6:7.1f} MB/s")
    for name, lvl, strat in (("zlib level 1", 1, 0), ("zlib level 1 Z_RLE", 1, zlib.Z_RLE), ("zlib level 6", 6, 0)):
        t = time.time(); n = 0
        for p in plain:
            c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat); n += len(c.compress(p) + c.flush())
        dt = time.time() - t
        print(f"deflate {name:20s}: {tot / dt / 1e6:7.1f} MB/s, ratio {n / tot:.4f}")
    bufs = [C.create_string_buffer(len(p) + 64) for p in plain]
    t = time.time(); sizes = [L.hm_deflate_block(p, len(p), o, len(p) + 64) for p, o in zip(plain, bufs)]; dt = time.time() - t
    print(f"deflate {'own (level 1)':20s}: {tot / dt / 1e6:7.1f} MB/s, ratio {sum(sizes) / tot:.4f}")
    own = [o.raw[:n] for o, n in zip(bufs, sizes)]
    t = time.time()
    for b, p, o in zip(own, plain, outs):
        assert L.hm_inflate_block(b, len(b), o, len(p))
    d2 = time.time() - t
    t = time.time(); back = [zlib.decompress(b, -15) for b in own]; d3 = time.time() - t
    assert back == plain
    print(f"inflate of own output: zlib {tot / d3 / 1e6:7.1f} MB/s, own {tot / d2 / 1e6:7.1f} MB/s")

if __name__ == "__main__":
    main()
