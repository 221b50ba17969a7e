#!/bin/bash
# chain kernel check on one box: smoke + GPU tests, then per-op times and the chain's clock64 stamps.  $1 = output prefix.
# The stamps are compiled in only with -DHM_CHAIN_STAMPS_BUILD=1: build that variant first (here, no GPU needed):
#   python tools/build_variant.py stamps -DHM_CHAIN_STAMPS_BUILD=1
p=${1:-chk}
timeout 120 python __graft_entry__.py smoke > gpurun_out/${p}_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/${p}_smoke.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${p}_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/${p}_pytest.log
if [ -f ab/stamps.so ]; then export HM_ENGINE_LIB=$PWD/ab/stamps.so; fi
HM_OP_TIMES=1 HM_CHAIN_STAMPS=1 timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-baseline --no-queue 2> gpurun_out/${p}_chain.log > gpurun_out/${p}_bench.json; echo bench rc=$?
unset HM_ENGINE_LIB
grep "op ms" gpurun_out/${p}_chain.log | tail -3
