"""bam_copy_bench.py -- host-only throughput of the BGZF/BAM reader + writer (hm_bam_copy: inflate -> records -> deflate), the part of
`hifimeth-b200 call` that bounds it when a GPU has fewer than ~8 host cores.  No GPU needed.
  python tools/bam_copy_bench.py in.bam [--threads 8] [--levels 1,6] [--lib path/to/libhm_engine.so]"""
import argparse, ctypes as C, os, sys, time

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bam")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--levels", default="1,6", help="comma separated BGZF levels of the copy; -1 = read only (inflate + record framing, no writer)")
    ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hifimeth_b200", "libhm_engine.so"))
    ap.add_argument("--out", default="/tmp/bam_copy_bench.out.bam")
    a = ap.parse_args()
    L = C.CDLL(a.lib)
    L.hm_bam_copy.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.hm_bam_copy.restype = C.c_long
    size = os.path.getsize(a.bam)
    for lvl in [int(x) for x in a.levels.split(",")]:
        best = None
        for _ in range(2):
            t = time.time()
            n = L.hm_bam_copy(a.bam.encode(), a.out.encode(), a.threads, lvl)
            if lvl < 0 and os.path.exists(a.out):
                os.unlink(a.out)
            dt = time.time() - t
            best = dt if best is None else min(best, dt)
        print(f"level {lvl}: {n} records, {best:.3f} s, {n / best:9.0f} records/s, in {size / 1e6:.1f} MB, out {(os.path.getsize(a.out) if os.path.exists(a.out) else 0) / 1e6:.1f} MB, "
              f"{a.threads} threads ({os.path.basename(os.path.realpath(a.lib))})")

if __name__ == "__main__":
    main()
