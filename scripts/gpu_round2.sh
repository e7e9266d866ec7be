#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for k in 1 2 4; do
  echo "== FWD_K=$k BWD_K=$k"; FAST=1 GAB200_FWD_K=$k GAB200_BWD_K=$k python scripts/quick_timing.py 2>&1 | tail -4
done
python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
