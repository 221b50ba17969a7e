#!/usr/bin/env python
"""bench.py -- headline measurement of the read-level `hifimeth call` hot path (BASELINE.json).

Metric: CpG+CHG+CHH sites/s.  Workload at every N: BASELINE.json configs[1] per GPU -- 1,000 synthetic HiFi reads x 15 kb
with fi/ri/fp/rp kinetics, all three contexts in one pass (5.8 M sites, 15 Mbase).  A "step" is one pass of the whole hot
path over that batch through the C ABI: kinetics decode -> site scan -> feature maps -> CNN (tcgen05) -> ML bytes.

  value   sites/s with the batch already resident in HBM (HM_SUBMIT_SKIP_H2D | SKIP_D2H), device time from the CUDA
          events the library records on the stream it launches on, max over ranks.
  e2e     the same through hm_batch_submit / hm_batch_collect with HOST buffers: pinned staging -> H2D -> kernels -> D2H of
          call_off / n_fwd / qoff / ML, wall clock between device synchronisations.
  roofline   the CNN's tensor-core kernel family (dense_gemm_kernel): algorithmic FLOPs of the step (22 297 600 per
          CpG/CHG site, 22 881 280 per CHH site) / summed device time of its launches in the step.
  cpu_baseline / --impl reference   the reference-faithful CPU pipeline on this host's cores: the reference's own feature
          code (oracle/_ref, compiled from /root/reference) or its C restatement + fp32 torch forward of the ONNX weights.

Reads are independent, so N GPUs = N engines on N batches with no collective ("scaling": "weak").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FLOP_K11 = 22_297_600.0
FLOP_K13 = 22_881_280.0
N_READS, READ_LEN, SEED = 1000, 15000, 20261


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1385.3), d.get("hbm_gbs", 6554.2), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def physical_cores() -> int:
    """Physical cores by the /proc/cpuinfo rule of src/corelib/get_core_count.cpp:68-115 (unique physical id x core id)."""
    try:
        seen, phys, core = set(), None, None
        for line in open("/proc/cpuinfo"):
            if line.startswith("physical id"):
                phys = line.split(":")[1].strip()
            elif line.startswith("core id"):
                core = line.split(":")[1].strip()
            elif not line.strip():
                if phys is not None and core is not None:
                    seen.add((phys, core))
                phys = core = None
        n = len(seen)
    except OSError:
        n = 0
    n = n or (os.cpu_count() or 1)
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    return max(1, n)


def cpu_pipeline(n_reads: int, threads: int, site_batch: int = 512):
    """The reference-faithful CPU path on a bounded sample of the workload.  Returns (sites, reads, seconds, kind)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    from hifimeth_b200 import synth
    from oracle import cnn_oracle, hmoracle

    torch.set_num_threads(threads)
    models = cnn_oracle.load_models(ROOT / "models")
    batch, reads = synth.make_reads(n_reads, READ_LEN, SEED)
    O = hmoracle.oracle()
    R = hmoracle.ref()
    bodies = [synth.record_body(r) for r in reads] if R.available else None
    kind = "reference feature code (oracle/_ref) + fp32 torch forward of the ONNX weights" if R.available else "C restatement (oracle/) + fp32 torch forward"
    t0 = time.perf_counter()
    sites = O.batch_sites(batch, 7)

    def feats(r):
        out = []
        for c in range(3):
            if R.available:
                n = int((sites[r]["ctx"] == c).sum())
                f = R.extract_features(bodies[r], c, 0, n)[0] if n else np.zeros((0, 401, 8), np.float32)
            else:
                f = O.batch_features(batch, sites, r, np.nonzero(sites[r]["ctx"] == c)[0])
            out.append(f)
        return out

    n_sites = 0
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        for per_ctx in ex.map(feats, range(n_reads)):
            for c, f in enumerate(per_ctx):
                for i in range(0, len(f), site_batch):
                    lg = cnn_oracle.forward_logits(models[c], f[i:i + site_batch])
                    cnn_oracle.logits_to_prob_ml(lg)
                n_sites += len(f)
    return n_sites, n_reads, time.perf_counter() - t0, kind


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = physical_cores()
    n_reads = int(os.environ.get("HM_CPU_SAMPLE_READS", "16"))
    for _ in range(min(args.warmup, 1)):
        cpu_pipeline(1, threads)
    vals, reads_s, secs = [], [], []
    for _ in range(args.steps):
        s, r, dt, kind = cpu_pipeline(n_reads, threads)
        vals.append(s / dt)
        reads_s.append(r / dt)
        secs.append(dt)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "CpG+CHG+CHH sites/sec", "value": v, "unit": "sites/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: {N_READS} reads x {READ_LEN} b, CpG+CHG+CHH; each step a bounded sample of {n_reads} reads"},
            "reads_per_s": float(np.mean(reads_s)),
            "cpu_baseline": {"value": v, "unit": "sites/s", "cores": threads, "kind": "port", "sample": f"{n_reads} reads x {READ_LEN} b per step; {kind}; site batch 512"},
            "e2e": {"value": v, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=N_READS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cnn-mode", type=int, default=int(os.environ.get("HM_CNN_MODE", "0")))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    from hifimeth_b200 import engine as hme
    from hifimeth_b200 import synth

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch, _ = synth.make_reads(args.reads, READ_LEN, SEED + rank)
    eng = hme.Engine(device=local, n_slots=1, max_reads=args.reads, max_bases=batch.n_bases + 1024, cnn_mode=args.cnn_mode)
    n = eng.stage(0, batch)

    # ---- e2e: host buffers, H2D + kernels + D2H every step ---------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.submit(0, n)
        res = eng.collect(0, copy=False)
    sites_step = res.n_calls
    n_sites = res.n_sites
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        eng.submit(0, n)
        eng.collect(0, copy=False)
        t = eng.timing(0)
        h2d, d2h = t.h2d_bytes, t.d2h_bytes
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- kernel-only: inputs resident in HBM -------------------------------------------------------------------------------------
    flags = hme.HM_SUBMIT_SKIP_H2D | hme.HM_SUBMIT_SKIP_D2H
    for _ in range(args.warmup):
        eng.submit(0, n, flags)
        eng.collect(0, copy=False)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    dev_ms = top_ms = 0.0
    launches = top_launches = 0
    stage_ms = np.zeros(4)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.submit(0, n, flags)
        eng.collect(0, copy=False)
        t = eng.timing(0)
        dev_ms += t.total_ms
        top_ms += t.top_kernel_ms
        launches += t.kernel_launches
        top_launches += t.top_kernel_launches
        stage_ms += np.array([t.decode_ms, t.scan_ms, t.cnn_ms, t.d2h_ms])
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()

    times = torch.tensor([dev_ms / 1e3, e2e_s, top_ms / 1e3, wall_s], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(sites_step), float(launches), float(args.reads)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_s, e2e_s, top_s, wall_s = (float(x) for x in times.tolist())
    sites_all, launches_all, reads_all = (float(x) for x in tot.tolist())

    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        flop_step = FLOP_K11 * (n_sites[0] + n_sites[1]) + FLOP_K13 * n_sites[2]
        achieved = flop_step * args.steps / max(top_s, 1e-9) / 1e12 if top_s > 0 else 0.0
        line = {
            "metric": "CpG+CHG+CHH sites/sec", "value": sites_all * args.steps / dev_s, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (hi/lo split bf16 operands, fp32 accumulate)" if args.cnn_mode == 0 else "f32", "data": "synthetic",
            "config": {"workload": f"configs[1] per GPU: {args.reads} reads x {READ_LEN} b with fi/ri/fp/rp, CpG+CHG+CHH in one pass",
                       "sites_per_step_per_gpu": sites_step, "sites_by_context": list(n_sites), "cnn_mode": "tensor" if args.cnn_mode == 0 else "fp32_simt",
                       "l2": "working set per step (activation maps, GBs) exceeds the 126 MB L2; no explicit flush"},
            "reads_per_s": reads_all * args.steps / dev_s,
            "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "stage_ms_per_step": {k: float(v) / args.steps for k, v in zip(("decode", "scan", "cnn", "d2h"), stage_ms)},
            "e2e": {"value": sites_all * args.steps / e2e_s, "unit": "sites/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         # DRAM read+write of the kernel family per step: ncu, profiles/r1_launches_v9_final.csv (51.3 GB
                         # for a 128-read step, scaled by reads; 58.2 GB before conv1 + conv2 were fused)
                         "traffic": 51.3e9 * args.reads / 128.0, "traffic_unit": "bytes per step",
                         "kernel": "dense_gemm2_kernel + dense_fused12_kernel + dense_gemm_kernel (every op of the dense plan)", "launches_per_step": top_launches // max(args.steps, 1), "peak_source": peak_src,
                         "note": "achieved = algorithmic FLOPs of the per-site network / summed device time of the kernel family; the dense plan "
                                 "shares conv work between overlapping windows, so executed FLOPs are lower (DESIGN.md)"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is a rank-0, N = 1 measurement
            threads = physical_cores()
            nr = int(os.environ.get("HM_CPU_SAMPLE_READS", "32"))
            s, r, dt, kind = cpu_pipeline(nr, threads)
            line["cpu_baseline"] = {"value": s / dt, "unit": "sites/s", "cores": threads, "kind": "port",
                                    "sample": f"{nr} reads x {READ_LEN} b ({s} sites, {dt:.1f} s); {kind}; site batch 512"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
