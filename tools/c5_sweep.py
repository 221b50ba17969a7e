"""BASELINE.json configs[4]: kernel microbench vs batch size -- decode / scan / gather HBM GB/s and the CNN's algorithmic
TFLOP/s for site batches of ~1 k ... ~1 M sites cut from 20 kb reads (0.39 sites per base: 1 read = ~7.8 k sites).
One JSON line per batch size.  Run on a B200:  python tools/c5_sweep.py [--reads 1 4 16 64 128]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hifimeth_b200 import engine as hme, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, nargs="*", default=[0, 1, 4, 16, 64, 128], help="reads per batch; 0 = one 2.6 kb read (~1 k sites)")
    ap.add_argument("--len", type=int, default=20000)
    a = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = float(peaks.get("hbm_gbs", 6554.2))
    bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1385.3)))
    for n_reads in a.reads:
        length = a.len if n_reads > 0 else 2600
        batch, _ = synth.make_reads(max(n_reads, 1), length, 20264)
        eng = hme.Engine(n_slots=1, max_reads=max(n_reads, 1), max_bases=batch.n_bases + 1024)
        calls = eng.call(batch)
        row = {"reads": max(n_reads, 1), "read_len": length, "sites": calls.n_calls}
        for name in ("decode", "scan", "mm", "gather", "cnn"):
            n_sites = min(calls.n_calls, 1 << 18) if name == "gather" else 0
            for _ in range(2):  # first round warms clocks and caches
                ms, by, fl = eng.microbench(0, name, n_sites, 10)
            if name == "cnn":
                row["cnn_ms"] = ms
                row["cnn_sites_per_s"] = calls.n_calls / ms * 1e3
                row["cnn_algorithmic_TFLOPs"] = fl / ms / 1e9
                row["cnn_frac_of_measured_bf16_peak"] = fl / ms / 1e9 / bf16
            else:
                row[name + "_us"] = ms * 1e3
                row[name + "_GBs"] = by / ms / 1e6
                row[name + "_frac_of_measured_hbm_peak"] = by / ms / 1e6 / hbm
        print(json.dumps(row), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
