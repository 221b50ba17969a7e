"""Microbenchmark of dense_gemm_kernel through hm_debug_dense_op: device time per launch for plan-shaped ops, with the
experiment switches of DenseOp::variant (HM_DENSE_VARIANT: 1 = hi*hi pass only, 2 = no MMA, 4 = no epilogue stores).
Run on a B200:  python tools/dense_microbench.py"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hifimeth_b200 import engine as hme  # noqa: E402

ROWS = 128 * 148 * int(os.environ.get('HM_TILES_PER_CTA', '16'))


def run(cin, cout, shifts, variant=0, head=False, conv1=0):
    os.environ["HM_DENSE_VARIANT"] = str(variant)
    os.environ["HM_DENSE_REPS"] = os.environ.get("HM_DENSE_REPS", "400")  # long enough for clocks to ramp; last 10 timed
    rng = np.random.default_rng(0)
    rows_alloc = ROWS + max(shifts) + 32 + 640
    x = rng.standard_normal((rows_alloc, cin)).astype(np.float32)
    if conv1:
        terms = [(0, shifts[0], (rng.standard_normal((conv1, 8, cout)) / 9).astype(np.float32))]
    else:
        terms = [(0, sh, (rng.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)) for sh in shifts]
    bias = np.zeros(cout, np.float32)
    kw = dict(w2=np.zeros((2, cout), np.float32), b2=np.zeros(2, np.float32)) if head else {}
    hme.debug_dense_op([x], terms, bias, ROWS, conv1_taps=conv1, **kw)
    ms = hme.load_library().hm_debug_last_op_ms()
    tiles_per_cta = ROWS // 128 / 148
    n_mma = (cin // 16 if not conv1 else (conv1 + 1) // 2) * len(terms) * (1 if variant & 1 else 3) * (0 if variant & 2 else 1)
    cyc_tile = ms * 1e-3 * 1.92e9 / tiles_per_cta
    macs = ROWS * (cin if not conv1 else 16 * ((conv1 + 1) // 2)) * cout * len(terms) * (1 if variant & 1 else 3)
    print(f"cin {cin:3d} cout {cout:3d} terms {len(shifts)} variant {variant}: {ms*1e3:8.1f} us  {cyc_tile:8.0f} cyc/tile  "
          f"{cyc_tile / max(n_mma, 1):6.1f} cyc/MMA  {2 * macs / ms / 1e9:7.1f} TFLOP/s", flush=True)


def sweep_ring():
    for ring in (2, 3, 4, 6, 8):
        os.environ["HM_DENSE_RING"] = str(ring)
        for v in (6, 2, 0):
            print("ring", ring, end="  ")
            run(128, 128, [0], v)
    os.environ.pop("HM_DENSE_RING")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "overlap":
        for v in (0, 8, 10, 2, 6, 4):
            run(128, 128, [0, 2, 4], v)
        for v in (0, 8, 10):
            run(128, 128, [0, 2, 4], v)
        for v in (0, 8, 10):
            run(128, 128, [0], v)
        run(96, 96, [0, 16, 32], 0)
        run(64, 64, [0, 0], 0)
        run(8, 128, [0], 0, conv1=11)
        for v in (0, 8, 10):
            run(96, 96, [0, 16, 32], v)
        for v in (0, 8, 10):
            run(64, 64, [0, 0], v)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "conv1":
        # where conv1-form tiles spend their time: 0 = all, 2 = no MMA, 4 = no epilogue stores, 8 = no epilogue work
        for v in (0, 2, 4, 8, 10):
            run(8, 128, [0], v, conv1=11)
        for v in (0, 4, 8):
            run(96, 96, [0, 32, 64], v)
        os.environ["HM_DENSE_2CTA"] = "1"
        for v in (0, 4, 8):
            run(96, 96, [0, 32, 64], v)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pair":
        for two in (0, 1):
            if two:
                os.environ["HM_DENSE_2CTA"] = "1"
            print("CTA pair" if two else "single CTA")
            run(128, 128, [0, 2, 4], 0)
            run(128, 128, [0], 0)
            run(128, 96, [0, 8, 16], 0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        for v in (0, 8, 10):
            run(128, 128, [0, 2, 4], v)
        for v in (0, 8, 10):
            run(128, 128, [0], v)
        run(96, 96, [0, 16, 32], 0)
        run(64, 64, [0, 0], 0)
        run(8, 128, [0], 0, conv1=11)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ring":
        sweep_ring()
        sys.exit(0)
    for v in (0, 1, 2, 4, 6):
        run(128, 128, [0, 2, 4], v)
    run(128, 128, [0], 0)
    run(128, 128, [0], 1)
    run(128, 256, [0], 0)
    run(128, 256, [0, 2, 4], 0, head=True)
    run(128, 64, [0, 2, 4], 0)
    run(96, 96, [0, 16, 32], 0)
    run(64, 64, [0, 0], 0)
    run(8, 128, [0], 0, conv1=11)
    run(8, 128, [0], 6, conv1=11)
