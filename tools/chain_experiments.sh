#!/bin/bash
# Timing experiments on the chain kernel (results of the x* builds are wrong by construction): per-context chain time from HM_OP_TIMES.
run() { # label, lib ("cur" or ab/<name>), extra env
  lib=$2; if [ "$lib" = "cur" ]; then unset HM_ENGINE_LIB; else export HM_ENGINE_LIB=$PWD/ab/$lib.so; fi
  env $3 HM_CHAIN=1 HM_OP_TIMES=1 timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-queue 2>&1 >/dev/null | grep "op ms" | tail -3 | awk -v v="$1" '{printf "%s %s %s chain %s total %s | ", v, $1, $2, $(NF-3), $NF} END {print ""}'
  unset HM_ENGINE_LIB
}
run "cur copies=16" cur "HM_CHAIN_WCOPIES=16"
run "cur copies=1 " cur "HM_CHAIN_WCOPIES=1"
run "cur copies=64" cur "HM_CHAIN_WCOPIES=64"
run "x3 (no mma, no epi) copies=16" x3 "HM_CHAIN_WCOPIES=16"
run "x3 copies=1" x3 "HM_CHAIN_WCOPIES=1"
run "x7 (+ no slabs) copies=16" x7 "HM_CHAIN_WCOPIES=16"
run "x7 copies=1" x7 "HM_CHAIN_WCOPIES=1"
