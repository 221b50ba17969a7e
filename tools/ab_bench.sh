#!/bin/bash
# A/B of library builds on ONE box: tools/ab_bench.sh <out-prefix> <lib>[+ENV=VAL...] [...]   ("cur" = the in-tree library)
# Runs every build twice, interleaved, and prints ms_per_step per build.  "cur+HM_NO_CHAIN=1" = in-tree library with that env.
out=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for spec in "$@"; do
    lib=${spec%%+*}
    envs=""
    if [ "$spec" != "$lib" ]; then envs=$(echo "${spec#*+}" | tr '+' ' '); fi
    if [ "$lib" = "cur" ]; then unset HM_ENGINE_LIB; else export HM_ENGINE_LIB=$PWD/ab/$lib.so; fi
    lib=$(echo "$spec" | tr '+=' '__')
    env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-queue > gpurun_out/${out}_${lib}_${rep}.json 2> gpurun_out/${out}_${lib}_${rep}.err
    python - "$lib" "$rep" gpurun_out/${out}_${lib}_${rep}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:>10} rep {sys.argv[2]}: {d['ms_per_step']:.2f} ms/step  {d['value']/1e6:.2f} M sites/s  e2e {d['e2e']['value']/1e6:.2f} M  clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
unset HM_ENGINE_LIB
