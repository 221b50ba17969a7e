#!/usr/bin/env python
"""bench.py -- headline measurement of the read-level `hifimeth call` hot path (BASELINE.json).

Metric: CpG+CHG+CHH sites/s.  Workload at every N: BASELINE.json configs[1] per GPU -- 1,000 synthetic HiFi reads x 15 kb
with fi/ri/fp/rp kinetics, all three contexts in one pass (5.8 M sites, 15 Mbase).  A "step" is one pass of the whole hot
path over that batch through the C ABI: kinetics decode -> site scan -> feature maps -> CNN (tcgen05) -> ML bytes.

  value   sites/s with the batch already resident in HBM (HM_SUBMIT_SKIP_H2D | SKIP_D2H), device time from the CUDA
          events the library records on the stream it launches on, max over ranks.
  e2e     the same through hm_batch_submit / hm_batch_collect with HOST buffers: pinned staging -> H2D -> kernels -> D2H of
          call_off / n_fwd / qoff / ML, wall clock between device synchronisations.
  roofline   the CNN's tensor-core kernel family: `achieved` / `frac` = ALGORITHMIC FLOPs of the step (22 297 600 per
          CpG/CHG site, 22 881 280 per CHH site) / device time of the dense plan; next to it what the silicon
          did: `executed_tflops` (FLOPs the launches issued, counted by the engine) and `executed_frac` of the sustained bf16
          peak, `traffic` (DRAM bytes of the family per step from the committed ncu launch list named in `traffic_source`) and
          `hbm_frac` = traffic / time / measured HBM peak, `algorithmic_bytes` (what has to cross HBM: the H2D + D2H bytes).
  cpu_baseline / --impl reference   the reference-faithful CPU pipeline on this host's cores: the reference's own feature
          code (oracle/_ref, compiled from /root/reference) or its C restatement + fp32 torch forward of the ONNX weights;
          site batch 512 (its best case) as the headline, the reference's default 32 and configs[0] (CpG only) as variants.
  gpu_library_baseline / --impl torchscript-gpu   the reference's app-gpu design re-hosted on this GPU: torch.jit.load of
          models/*.pt on cuda at batch 4096 (src/app-gpu/hifimeth-gpu/5mc_call_gpu.cpp:194-218,309-334), CNN forward only,
          fp32 and TF32 -- the library kernels the hand-written path has to beat.
  queue   the PRODUCT's multi-GPU path: ONE process, `hifimeth-b200 call --devices 0..N-1` on a synthetic BAM of
          configs[2] shape (20 kb reads), wall clock from process start to output closed, with the CLI's phase timers.

Reads are independent, so N GPUs = N engines with no collective ("scaling": "weak").  Under torchrun every rank times its
own batch (`value`, device time, max over ranks = `kernel_scaling`); then rank 0 alone runs the one-process host work queue
over all N devices and reports it as `queue` -- and, for N > 1, as `e2e` (BAM in -> mod BAM out is the call a user makes).
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
        self.index, self.rows, self.proc, self.t_mark = index, [], None, None

    def mark(self):
        """The timed region starts now: only samples taken from here on count (nvidia-smi needs seconds to start on an 8-GPU box,
        so it is launched before the warm-up steps)."""
        self.t_mark = time.perf_counter()

    def start(self):
        """In-process NVML polling (50 ms) when nvidia_ml_py is importable: it answers from the first poll, where a fresh nvidia-smi
        needed longer than a whole 8-GPU bench run before its first line.  nvidia-smi -lms otherwise."""
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons_fn(h)
            self.nvml_stop = threading.Event()

            def poll():
                # NVML reason bits: 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown, 0x40 hw_thermal_slowdown
                while not self.nvml_stop.is_set():
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        r = int(reasons_fn(h))
                        act = lambda bit: "Active" if r & bit else "Not Active"
                        self.rows.append((time.perf_counter(), [str(sm), str(mx), "", act(0x8), act(0x40), act(0x20), act(0x4)]))
                    except Exception:
                        pass
                    self.nvml_stop.wait(0.05)

            threading.Thread(target=poll, daemon=True).start()
            self.source = "nvml"
            return
        except Exception:
            self.nvml_stop = None
        try:
            self.source = "nvidia-smi"
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if getattr(self, "nvml_stop", None) is not None:
            self.nvml_stop.set()
        if self.proc:
            self.proc.terminate()
        timed = [r for t, r in self.rows if self.t_mark is None or t >= self.t_mark]
        window = "timed region"
        if not timed:  # a timed region shorter than one sampling period: the warm-up steps ran the same kernels
            timed, window = [r for _, r in self.rows], "warm-up + timed region"
        self.rows = timed
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "window": window, "source": getattr(self, "source", None)}


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


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def traffic_record():
    """profiles/traffic.json: DRAM bytes per step of the CNN kernel family, condensed from a committed ncu launch list by
    tools/traffic_from_ncu.py.  None when absent."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(p.read_text())
        return float(d["cnn_dram_bytes_per_step"]), int(d["reads"]), str(d["source"])
    except (OSError, KeyError, ValueError):
        return None


def gpu_library_baseline(device: int, feats_by_ctx, iters: int = 8):
    """torch.jit.load(models/*.pt).cuda() forward at batch 4096 on real feature windows: what the reference's app-gpu binary
    runs per batch (5mc_call_gpu.cpp:194-218), minus its CPU feature extraction and PCIe copies.  Sites/s, fp32 and TF32."""
    import torch

    out = {"kind": "torch.jit.load(models/{CpG,CHG,CHH}.pt).cuda(), batch 4096, CNN forward only (features resident in HBM, "
                   "no feature extraction, no H2D); cuDNN/cuBLAS kernels of torch " + torch.__version__,
           "batch": 4096, "unit": "sites/s"}
    dev = torch.device("cuda", device)
    for mode in ("fp32", "tf32"):
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        tot_sites, tot_s, per_ctx = 0, 0.0, {}
        for name, f in feats_by_ctx.items():
            m = torch.jit.load(str(ROOT / "models" / f"{name}.pt"), map_location=dev).eval()
            x = torch.from_numpy(f).to(dev)
            with torch.no_grad():
                for _ in range(3):
                    m(x)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    m(x)
                e1.record()
                torch.cuda.synchronize(dev)
            dt = e0.elapsed_time(e1) / 1e3
            per_ctx[name] = len(f) * iters / dt
            tot_sites += len(f) * iters
            tot_s += dt
        out[mode] = {"value": tot_sites / tot_s, "by_context": per_ctx}
    torch.backends.cudnn.allow_tf32 = True
    out["value"] = out["tf32"]["value"]
    out["note"] = "value = TF32 (torch's default for cuDNN convolutions); the site mix is 1:1:1 here, CHH-heavy in the workload"
    return out


def write_queue_bam(path, n_reads: int, read_len: int, seed: int) -> dict:
    """Synthetic BAM of configs[2] shape for the queue bench.  500 distinct reads are generated and BGZF-compressed once; that
    run of blocks is written n_reads / 500 times (BGZF members are independent, and the BAM stream is the concatenation of
    their payloads), so a multi-GB input costs seconds to make.  Says so in the returned description."""
    import struct
    import zlib

    from hifimeth_b200 import synth

    def bgzf(data: bytes, level: int = 1) -> bytes:
        out = bytearray()
        for off in range(0, len(data), 0xff00):
            chunk = data[off:off + 0xff00]
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            out += struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, 6) + b"BC" + struct.pack("<HH", 2, len(comp) + 25)
            out += comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
        return bytes(out)

    distinct = min(500, n_reads)
    _, reads = synth.make_reads(distinct, read_len, seed)
    body = bytearray()
    for r in reads:
        b = synth.record_body(r)
        body += struct.pack("<i", len(b)) + b
    text = b"@HD\tVN:1.6\tSO:unknown\n@RG\tID:synth\tPL:PACBIO\n"
    head = bgzf(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 0))
    run = bgzf(bytes(body))
    reps = max(1, n_reads // distinct)
    with open(path, "wb") as f:
        f.write(head)
        for _ in range(reps):
            f.write(run)
        f.write(bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0]))
    return {"reads": reps * distinct, "read_len": read_len, "bam_bytes": len(head) + reps * len(run) + 28, "raw_bytes": reps * len(body),
            "note": f"{distinct} distinct reads x {reps} (the same BGZF blocks repeated), input level 1"}


def queue_bench(n_gpus: int, reads_per_gpu: int, read_len: int = 20000, level: int = 6, repeat: int = 1) -> dict:
    """The product's multi-GPU path: ONE process, one reader, one worker per device, one ordered writer
    (hifimeth_b200/csrc/call_main.cpp; reference shape: src/app/hifimeth/mod_main.cpp:330-388).  Wall clock of the process."""
    import re
    import tempfile

    from hifimeth_b200 import build as hmbuild

    held = 0
    if not os.environ.get("HM_QUEUE_COLD"):
        try:
            import torch

            for i in range(min(n_gpus, torch.cuda.device_count())):
                torch.zeros(1, device=f"cuda:{i}")
                held += 1
            torch.cuda.synchronize()
        except Exception:
            pass
    tmp = Path(tempfile.mkdtemp(prefix="hm_queue_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    src, dst = tmp / "in.bam", tmp / "mod.bam"
    try:
        t0 = time.perf_counter()
        info = write_queue_bam(src, reads_per_gpu * n_gpus, read_len, SEED + 7)
        gen_s = time.perf_counter() - t0
        cmd = [str(hmbuild.EXE), "call", "--level", str(level), "--devices", ",".join(str(i) for i in range(n_gpus)), str(src), str(dst)]
        best = None
        for _ in range(repeat):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                return {"error": r.stderr.strip()[-400:]}
            if best is None or wall < best[0]:
                best = (wall, r.stderr)
        wall, log = best
        m = re.search(r"CpG (\d+), CHG (\d+), CHH (\d+)", log)
        sites = sum(int(x) for x in m.groups())
        ph = re.search(r"read\+inflate ([\d.]+), engine create ([\d.]+), pack ([\d.]+), submit ([\d.]+), collect\(wait\) ([\d.]+), "
                       r"assemble ([\d.]+), write\+deflate ([\d.]+); (\d+) batches of <= (\d+) bases on (\d+) worker\(s\), (\d+) host threads", log)
        tl = re.search(r"engines ready ([\d.]+), input inflated ([\d.]+), last batch collected ([\d.]+), engines destroyed ([\d.]+), output closed ([\d.]+)", log)
        out = {"value": sites / wall, "unit": "sites/s", "reads_per_s": info["reads"] / wall, "wall_s": wall, "sites": sites, "n_gpus": n_gpus,
               "workload": f"configs[2] shape: {info['reads']} reads x {read_len} b BAM -> mod BAM, one process, --devices 0..{n_gpus - 1}, output BGZF level {level}",
               "input": info, "out_bam_bytes": dst.stat().st_size, "gen_s": gen_s, "host_threads": os.cpu_count(), "cmd": " ".join(cmd[1:-2]),
               "contexts_held_by_parent": held,
               "contexts_note": "this process keeps a CUDA context open on every device while the CLI runs: a fresh process on an idle "
                                "box otherwise pays ~2 s of GPU re-initialisation PER DEVICE before its first kernel (measured: engines ready "
                                "at 5.1 s cold against 1.3 s with the contexts held, 2 GPUs) -- start-up cost of the box, not of the path"}
        if ph:
            keys = ("read_inflate", "engine_create", "pack", "submit", "collect_wait", "assemble", "write_deflate")
            out["phase_s_summed_per_role"] = {k: float(v) for k, v in zip(keys, ph.groups()[:7])}
            out["batches"], out["max_bases_per_batch"] = int(ph.group(8)), int(ph.group(9))
            out["host_threads"] = int(ph.group(11))
        if tl:
            ready, inflated, collected, destroyed, closed = (float(x) for x in tl.groups())
            out["timeline_s"] = {"engines_ready": ready, "input_inflated": inflated, "last_batch_collected": collected,
                                 "engines_destroyed": destroyed, "output_closed": closed}
            if collected > ready:
                out["steady_sites_per_s"] = sites / (collected - ready)
            # the phase that ends last names the limiter
            # (the executable leaves through _exit without waiting for the engines' teardown: `destroyed` is then 0)
            lim = "reader (BGZF inflate + framing)" if inflated >= collected - max(0.05, 0.1 * (collected - ready)) else \
                  "writer (record assembly hand-over + BGZF deflate)" if closed - max(destroyed, collected) > 0.25 * wall else \
                  "engine creation (CUDA context + pinned / device allocation)" if ready > 0.5 * wall else "GPU workers"
            out["limiter"] = lim
        return out
    finally:
        for p in (src, dst):
            p.unlink(missing_ok=True)
        try:
            tmp.rmdir()
        except OSError:
            pass


def cpu_pipeline(n_reads: int, threads: int, site_batch: int = 512, ctx_mask: int = 7):
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
    sites = O.batch_sites(batch, ctx_mask)

    def feats(r):
        out = []
        for c in range(3):
            if not ctx_mask & (1 << c):
                out.append(np.zeros((0, 401, 8), np.float32))
                continue
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
            "cpu_baseline": {"value": v, "unit": "sites/s", "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                             "sample": f"{n_reads} reads x {READ_LEN} b per step; {kind}; site batch 512"},
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
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch.jit GPU library baseline")
    ap.add_argument("--no-queue", action="store_true", help="skip the one-process host-work-queue run (hifimeth-b200 call)")
    ap.add_argument("--queue-reads", type=int, default=int(os.environ.get("HM_QUEUE_READS", "24000")), help="20 kb reads per GPU in the queue run")
    ap.add_argument("--queue-level", type=int, default=1, help="BGZF level of the queue run's output (1 = the CLI's default; htslib's default is 6)")
    ap.add_argument("--cnn-mode", type=int, default=int(os.environ.get("HM_CNN_MODE", "0")))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.impl == "torchscript-gpu":
        if rank == 0:
            import torch

            torch.cuda.set_device(local)
            rng = np.random.default_rng(SEED)
            feats = {n: rng.random((4096, 401, 8), dtype=np.float32) for n in ("CpG", "CHG", "CHH")}
            line = gpu_library_baseline(local, feats)
            line.update({"impl": "torchscript-gpu", "metric": "CpG+CHG+CHH sites/sec (CNN forward only)", "n_gpus": 1, "data": "synthetic (random feature windows)"})
            print(json.dumps(line), flush=True)
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
    eng = hme.Engine(device=local, n_slots=2, max_reads=args.reads, max_bases=batch.n_bases + 1024, cnn_mode=args.cnn_mode)

    # ---- e2e: HOST buffers every step -- the batch is packed into the slot's pinned SoA staging (eng.stage), copied H2D, run, and
    # its calls copied D2H into pinned memory.  Two staging slots, as the `call` driver uses them: step i + 1 is packed and submitted
    # while step i is on the device.  Wall clock between device synchronisations.
    def e2e_loop(steps):
        res, h2d, d2h = None, 0, 0
        for i in range(steps):
            slot = i & 1
            n_ = eng.stage(slot, batch)
            eng.submit(slot, n_)
            if i:
                res = eng.collect(slot ^ 1, copy=False)
        res = eng.collect((steps - 1) & 1, copy=False)
        t = eng.timing((steps - 1) & 1)
        return res, t.h2d_bytes, t.d2h_bytes

    res, _, _ = e2e_loop(max(args.warmup, 2))
    sites_step = res.n_calls
    n_sites = res.n_sites
    barrier()
    t0 = time.perf_counter()
    _, h2d, d2h = e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    n = eng.stage(0, batch)
    eng.submit(0, n)
    eng.collect(0, copy=False)

    # ---- kernel-only: inputs resident in HBM -------------------------------------------------------------------------------------
    flags = hme.HM_SUBMIT_SKIP_H2D | hme.HM_SUBMIT_SKIP_D2H
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        eng.submit(0, n, flags)
        eng.collect(0, copy=False)
    barrier()
    sampler.mark()
    dev_ms = top_ms = exec_flops = 0.0
    launches = top_launches = 0
    stage_ms = np.zeros(4)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.submit(0, n, flags)
        eng.collect(0, copy=False)
        t = eng.timing(0)
        dev_ms += t.total_ms
        top_ms += t.top_kernel_ms
        exec_flops += t.executed_flops
        launches += t.kernel_launches
        top_launches += t.top_kernel_launches
        stage_ms += np.array([t.decode_ms, t.scan_ms, t.cnn_ms, t.d2h_ms])
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # real feature windows for the GPU library baseline (validation kernel; needs a keep_debug engine: a small second one)
    lib_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline and args.cnn_mode == 0:
        try:
            small, _ = synth.make_reads(8, READ_LEN, SEED + 99)
            e2 = hme.Engine(device=local, n_slots=1, max_reads=8, max_bases=small.n_bases + 1024, keep_debug=True)
            r2 = e2.call(small)
            ctx = e2.dump_ctx(0, r2.n_calls)
            feats = {}
            for c, name in enumerate(("CpG", "CHG", "CHH")):
                idx = np.nonzero(ctx == c)[0][:4096]
                f = np.concatenate([e2.dump_features(0, int(k), 1) for k in idx[:256]])
                feats[name] = np.tile(f, (16, 1, 1))[:4096]  # 256 distinct real windows, tiled to the batch of 4096
            e2.close()
            lib_base = gpu_library_baseline(local, feats)
        except Exception as ex:  # the baseline must never take the bench line down
            lib_base = {"error": repr(ex)[:300]}

    times = torch.tensor([dev_ms / 1e3, e2e_s, top_ms / 1e3, wall_s, exec_flops], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(sites_step), float(launches), float(args.reads)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_s, e2e_s, top_s, wall_s, _ = (float(x) for x in times.tolist())
    sites_all, launches_all, reads_all = (float(x) for x in tot.tolist())
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()  # every rank has released its GPU: rank 0 now drives all N devices from one process

    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        flop_step = FLOP_K11 * (n_sites[0] + n_sites[1]) + FLOP_K13 * n_sites[2]
        achieved = flop_step * args.steps / max(top_s, 1e-9) / 1e12 if top_s > 0 else 0.0
        # what the silicon did (rank 0's own counters; every rank runs the same plan on its own batch)
        executed = exec_flops / max(top_ms / 1e3, 1e-9) / 1e12 if top_ms > 0 else 0.0
        tr = traffic_record()
        traffic = tr[0] * args.reads / tr[1] if tr else None
        hbm_frac = traffic / (top_ms / 1e3 / args.steps) / (peak_gbs * 1e9) if tr and top_ms > 0 else None
        executed_frac = executed / peak_tf
        roof = {"bound": "hbm" if (hbm_frac or 0) > executed_frac else "tensor",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "frac_is": "ALGORITHMIC per-site FLOPs (SURVEY.md s8d) / plan time / sustained bf16 peak; the plan shares conv work between "
                           "overlapping windows and pays 3 split-precision passes, so read it with executed_frac and hbm_frac",
                "algorithmic_flops_per_step": flop_step,
                "executed_tflops": executed, "executed_frac": executed_frac, "executed_flops_per_step": exec_flops / max(args.steps, 1),
                "traffic": traffic, "traffic_unit": "DRAM bytes per step (read + write) of the kernel family", "hbm_frac": hbm_frac,
                "hbm_peak_gbs": peak_gbs, "traffic_source": (tr[2] + f" ({tr[1]}-read step, scaled by reads; tools/traffic_from_ncu.py)") if tr else None,
                "algorithmic_bytes": int(h2d + d2h),
                "kernel": "dense_gemm2_kernel + dense_fused12_kernel + site_chain_kernel + dense_gemm_kernel (every op of the dense plan)",
                "plan_ms_per_step": top_ms / max(args.steps, 1), "plan_ms_is": "CUDA events around all launches of the family on their stream (includes "
                "the ~1 % of small kernels between them)", "launches_per_step": top_launches // max(args.steps, 1), "peak_source": peak_src}
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
            "e2e": {"value": sites_all * args.steps / e2e_s, "unit": "sites/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "hm_batch_acquire / fill the pinned SoA staging / hm_batch_submit / hm_batch_collect every step, two slots pipelined, wall clock"},
            "gpu_launches": int(launches_all),
            "roofline": roof,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is a rank-0, N = 1 measurement
            threads = physical_cores()
            nr = int(os.environ.get("HM_CPU_SAMPLE_READS", "32"))
            s, r, dt, kind = cpu_pipeline(nr, threads)
            line["cpu_baseline"] = {"value": s / dt, "unit": "sites/s", "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                                    "sample": f"{nr} reads x {READ_LEN} b ({s} sites, {dt:.1f} s); {kind}; site batch 512"}
            # the reference's default site batch (-s 32, src/app/hifimeth/mod_options.cpp:11) and configs[0] (CpG only)
            nv = max(2, nr // 4)
            s32, _, dt32, _ = cpu_pipeline(nv, threads, site_batch=32)
            sc, _, dtc, _ = cpu_pipeline(nr, threads, site_batch=32, ctx_mask=1)
            line["cpu_baseline"]["variants"] = [
                {"what": "all contexts, site batch 32 (the reference's default -s)", "value": s32 / dt32, "sample": f"{nv} reads ({s32} sites, {dt32:.1f} s)"},
                {"what": "configs[0]: CpG only, site batch 32", "value": sc / dtc, "sample": f"{nr} reads ({sc} sites, {dtc:.1f} s)"}]
            line["cpu_baseline"]["readme_derived"] = "README.md:31 implies 150-220 k sites/s on 48 threads with OpenVINO (3-4.6 k sites/s/thread); OpenVINO is not installable here"
        if lib_base is not None:
            line["gpu_library_baseline"] = lib_base
        if not args.no_queue and args.cnn_mode == 0:
            q = queue_bench(world, args.queue_reads, level=args.queue_level)
            line["queue"] = q
            if world >= 4 and "value" in q and args.queue_level != 6 and not os.environ.get("HM_QUEUE_NO_LEVEL6"):
                # below ~12 host cores per GPU zlib sets the pace: the same run with the output at htslib's default level 6 (what the
                # reference writes) shows how much of the host budget that deflate takes
                q6 = queue_bench(world, args.queue_reads, level=6)
                line["queue_level6"] = {k: q6[k] for k in ("value", "wall_s", "sites", "out_bam_bytes", "steady_sites_per_s", "limiter", "phase_s_summed_per_role",
                                                           "timeline_s", "cmd", "error") if k in q6}
            line["kernel_scaling"] = {"value": line["value"], "unit": "sites/s", "what": "N ranks, N resident batches, device time, max over ranks"}
            if world > 1 and "value" in q:
                # for N > 1 the end-to-end number IS the product's one-process queue path (BAM in -> mod BAM out)
                line["e2e_per_rank_abi"] = line["e2e"]
                line["e2e"] = {"value": q["value"], "unit": "sites/s", "h2d_bytes_per_step": int(q["input"]["reads"] * (4.5 * q["input"]["read_len"] + 16)),
                               "d2h_bytes_per_step": int(q["sites"] * 5 + q["input"]["reads"] * 8), "step": "the whole BAM",
                               "what": "hifimeth-b200 call --devices 0..N-1: BAM inflate, packing, H2D, kernels, D2H, MM/ML records, deflate; process wall clock incl. engine creation"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
