#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for cfg in "32 100000" "32 1024" "256 100000" "100000 100000" "32 32"; do set -- $cfg
  echo "== HEAVY_FWD=$1 HEAVY_BWD=$2"; FAST=1 GAB200_HEAVY_FWD=$1 GAB200_HEAVY_BWD=$2 python scripts/quick_timing.py 2>&1 | grep -E "stages|bw="
done
