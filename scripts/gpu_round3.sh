#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for hf in 64 128 256 512 1984; do for hb in 256 512 1024; do
  echo "== HEAVY_FWD=$hf HEAVY_BWD=$hb"; FAST=1 GAB200_HEAVY_FWD=$hf GAB200_HEAVY_BWD=$hb python scripts/quick_timing.py 2>&1 | grep -E "stages|bw=True"
done; done
echo "== KH=2"; FAST=1 GAB200_FWD_KH=2 python scripts/quick_timing.py 2>&1 | grep -E "stages|bw=True"
python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
