"""Condenses an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv`) of `bench.py --reads R --steps 1 --warmup 3` into
profiles/traffic.json, the record bench.py reads at run time for `roofline.traffic` (per-step DRAM bytes of the CNN kernel
family, scaled by reads) -- so the number in the bench line is tied to a committed profile, not to a literal in the code.

    python tools/traffic_from_ncu.py profiles/r2_launches_xxx.csv --reads 128 [--steps-in-capture 4] > profiles/traffic.json
"""
import argparse
import csv
import json
import re
import sys
from collections import defaultdict


def load(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--reads", type=int, required=True)
    ap.add_argument("--steps-in-capture", type=int, default=0, help="full steps the capture covers; 0 = count the decode_kernel launches (one per step)")
    a = ap.parse_args()
    per = defaultdict(lambda: defaultdict(float))   # kernel -> metric -> sum
    launches = defaultdict(set)
    rows = load(a.csv)
    # whole steps only: a step starts with its decode_kernel launch; launches after the last decode_kernel belong to a step the
    # capture (-c N) cut short
    starts = sorted({int(r["ID"]) for r in rows if "decode_kernel" in r["Kernel Name"] and "unpack" not in r["Kernel Name"]})
    lo, hi = (starts[0], starts[-1]) if len(starts) >= 2 else (0, 1 << 60)
    n_steps = max(1, len(starts) - 1) if len(starts) >= 2 else 1
    for r in rows:
        if not lo <= int(r["ID"]) < hi:
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<.*", "", name).replace("void ", "").split("::")[-1]
        v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)  # -> ms
        per[name][m] += v
        launches[name].add(r["ID"])
    steps = a.steps_in_capture or n_steps
    fam = [k for k in per if k.startswith("dense_") or k.startswith("site_chain")]
    out = {"source": a.csv, "reads": a.reads, "steps_in_capture": steps, "kernels": {}}
    tot_t = sum(per[k]["gpu__time_duration.sum"] for k in per)
    for k in sorted(per, key=lambda k: -per[k]["gpu__time_duration.sum"]):
        out["kernels"][k] = {"launches_per_step": len(launches[k]) / steps, "ms_per_step": per[k]["gpu__time_duration.sum"] / steps,
                             "share": per[k]["gpu__time_duration.sum"] / tot_t if tot_t else None,
                             "dram_read_bytes_per_step": per[k]["dram__bytes_read.sum"] / steps,
                             "dram_write_bytes_per_step": per[k]["dram__bytes_write.sum"] / steps}
    out["cnn_family"] = fam
    out["cnn_dram_bytes_per_step"] = sum(per[k]["dram__bytes_read.sum"] + per[k]["dram__bytes_write.sum"] for k in fam) / steps
    out["all_dram_bytes_per_step"] = sum(per[k]["dram__bytes_read.sum"] + per[k]["dram__bytes_write.sum"] for k in per) / steps
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
