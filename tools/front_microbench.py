"""BASELINE.json config 5: HBM GB/s of the integer/byte kernels (decode, scan, gather, MM text) on configs[1]-shaped reads.
Run on a B200:  python tools/front_microbench.py"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hifimeth_b200 import engine as hme, synth  # noqa: E402

peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6554.2) \
    if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0
batch, _ = synth.make_reads(1000, 15000, 20261)
eng = hme.Engine(n_slots=1, max_reads=1000, max_bases=batch.n_bases + 1024)
eng.call(batch)
for name, n_sites in (("decode", 0), ("scan", 0), ("mm", 0), ("stats", 0), ("gather", 1 << 14), ("gather", 1 << 18)):
    for _ in range(2):
        ms, by, fl = eng.microbench(0, name, n_sites, 20)
    print(json.dumps({"kernel": name, "n_sites": n_sites, "ms_per_launch": ms, "algorithmic_bytes": by, "GB_per_s": by / ms / 1e6,
                      "frac_of_measured_hbm_peak": by / ms / 1e6 / peak}))
eng.close()
