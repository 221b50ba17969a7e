"""Repeated engine creations and pipelined two-slot batches on one GPU: every run must succeed and return the same bytes
(the path is deterministic).  Catches rare launch failures and run-to-run differences.
    python tools/stress_engine.py [--iters 25] [--reads 300]"""
import argparse
import hashlib
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hifimeth_b200 import engine as hme, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=25)
    ap.add_argument("--reads", type=int, default=300)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    batch, _ = synth.make_reads(a.reads, 15000, 20261)
    ref, bad, t0 = None, 0, time.time()
    for it in range(a.iters):
        try:
            eng = hme.Engine(n_slots=2, max_reads=a.reads, max_bases=batch.n_bases + 1024)
            digests = []
            for i in range(a.steps):
                slot = i & 1
                n = eng.stage(slot, batch)
                eng.submit(slot, n)
                if i:
                    r = eng.collect(slot ^ 1)
                    digests.append(hashlib.sha1(r.ml.tobytes() + r.qoff.tobytes()).hexdigest())
            r = eng.collect((a.steps - 1) & 1)
            digests.append(hashlib.sha1(r.ml.tobytes() + r.qoff.tobytes()).hexdigest())
            eng.close()
        except Exception as ex:  # noqa: BLE001
            bad += 1
            print(f"iter {it}: FAILED {ex!r}", flush=True)
            continue
        if ref is None:
            ref = digests[0]
        if any(d != ref for d in digests):
            bad += 1
            print(f"iter {it}: results differ between runs: {digests} vs {ref}", flush=True)
    print(f"{a.iters} iterations x {a.steps} batches, {bad} bad, {time.time() - t0:.0f} s, digest {ref}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
