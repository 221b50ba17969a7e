"""BASELINE.json configs[3] shape, scaled: `hifimeth-b200 call in.bam mod.bam` end to end -- BGZF inflate, record framing, the
device path, MM/ML record assembly, BGZF deflate -- on a synthetic BAM written here.  Wall clock of the whole process.
Run on a B200:  python tools/cli_bench.py [--reads 1000] [--len 15000] [--threads 0] [--level 6] [--devices 0]"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hifimeth_b200 import build as hmbuild, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1000)
    ap.add_argument("--len", type=int, default=15000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--devices", default="")
    ap.add_argument("--batch-reads", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--extra", default="", help="extra CLI arguments, space separated")
    ap.add_argument("--warm", action="store_true", help="hold a CUDA context on every device while the CLI runs (what nvidia-persistenced "
                    "does on a production box: without it a fresh process pays the GPU's re-initialisation, ~2 s per device here)")
    ap.add_argument("--sweep", default="", help="comma separated CORES[:ENV=VAL[:ENV=VAL]] -- one run per entry on the same BAM, pinned to "
                    "that many host cores with taskset (-t CORES): the host-bound regime of a many-GPU box on one GPU.  One JSON line each")
    a = ap.parse_args()
    hmbuild.build()
    tmp = Path(tempfile.mkdtemp(prefix="hm_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    src, dst = tmp / "in.bam", tmp / "mod.bam"
    t0 = time.time()
    bodies = []
    chunk = 250
    for first in range(0, a.reads, chunk):  # chunked so the per-read dicts never all live at once
        _, reads = synth.make_reads(min(chunk, a.reads - first), a.len, seed=20263 + first)
        for i, r in enumerate(reads):
            r["name"], r["zm"] = f"synth/{first + i}/ccs", first + i
        bodies += [synth.record_body(r) for r in reads]
    in_bytes = synth.write_bam(src, bodies, level=1)
    raw_bytes = sum(len(b) + 4 for b in bodies)
    del bodies
    gen_s = time.time() - t0
    exe = hmbuild.EXE
    cmd = [str(exe), "call", "--level", str(a.level)]
    if a.threads:
        cmd += ["-t", str(a.threads)]
    if a.devices:
        cmd += ["--devices", a.devices]
    if a.batch_reads:
        cmd += ["-b", str(a.batch_reads)]
    cmd += a.extra.split() + [str(src), str(dst)]
    if a.warm:
        import torch

        for i in range(torch.cuda.device_count()):
            torch.zeros(1, device=f"cuda:{i}")
        torch.cuda.synchronize()
    if a.sweep:
        for spec in a.sweep.split(","):
            parts = spec.split(":")
            cores = int(parts[0])
            env = dict(os.environ)
            env.update(dict(kv.split("=", 1) for kv in parts[1:]))
            c = ["taskset", "-c", f"0-{cores - 1}", str(exe), "call", "--level", str(a.level), "-t", str(cores)]
            if a.devices:
                c += ["--devices", a.devices]
            c += a.extra.split() + [str(src), str(dst)]
            best = None
            for _ in range(a.repeat):
                t0 = time.time()
                r = subprocess.run(c, capture_output=True, text=True, env=env)
                wall = time.time() - t0
                if r.returncode != 0:
                    sys.stderr.write(r.stderr)
                    raise SystemExit(r.returncode)
                if best is None or wall < best[0]:
                    best = (wall, r.stderr)
            wall, log = best
            sites = sum(int(x) for x in re.search(r"CpG (\d+), CHG (\d+), CHH (\d+)", log).groups())
            tl = re.search(r"engines ready ([\d.]+), input inflated ([\d.]+), last batch collected ([\d.]+), engines destroyed ([\d.]+), output closed ([\d.]+)", log)
            ph = re.search(r"read\+inflate ([\d.]+), engine create ([\d.]+), pack ([\d.]+), submit ([\d.]+), collect\(wait\) ([\d.]+), assemble ([\d.]+), write\+deflate ([\d.]+)", log)
            out = {"cores": cores, "env": parts[1:], "reads": a.reads, "read_len": a.len, "level": a.level, "wall_s": round(wall, 3), "sites": sites,
                   "sites_per_s": sites / wall, "out_bam_bytes": dst.stat().st_size, "in_bam_bytes": in_bytes}
            if tl:
                ready, inflated, collected, destroyed, closed = (float(x) for x in tl.groups())
                out["timeline_s"] = {"engines_ready": ready, "input_inflated": inflated, "last_batch_collected": collected, "output_closed": closed}
                out["steady_sites_per_s"] = sites / max(closed - ready, 1e-9)
            if ph:
                out["phase_s"] = dict(zip(("read_inflate", "engine_create", "pack", "submit", "collect_wait", "assemble", "write_deflate"), (float(x) for x in ph.groups())))
            print(json.dumps(out), flush=True)
        for p in (src, dst):
            p.unlink(missing_ok=True)
        tmp.rmdir()
        return
    best = None
    for _ in range(a.repeat):
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.time() - t0
        if r.returncode != 0 or os.environ.get("HM_VERBOSE"):
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise SystemExit(r.returncode)
        m = re.search(r"CpG (\d+), CHG (\d+), CHH (\d+)", r.stderr)
        sites = sum(int(x) for x in m.groups())
        if best is None or wall < best["wall_s"]:
            best = {"wall_s": wall, "sites": sites, "stderr": r.stderr.strip().splitlines()[-7:]}
    print(json.dumps({"workload": f"{a.reads} reads x {a.len} b BAM -> mod BAM through the CLI", "cmd": " ".join(cmd[1:-2]),
                      "in_bam_bytes": in_bytes, "in_raw_bytes": raw_bytes, "out_bam_bytes": dst.stat().st_size,
                      "wall_s": best["wall_s"], "sites": best["sites"], "sites_per_s": best["sites"] / best["wall_s"],
                      "reads_per_s": a.reads / best["wall_s"], "gen_s": gen_s, "host_cores": os.cpu_count(), "log": best["stderr"]}))
    for p in (src, dst):
        p.unlink(missing_ok=True)
    tmp.rmdir()


if __name__ == "__main__":
    main()
