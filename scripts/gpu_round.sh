#!/bin/bash
# One gpurun call: parity tests, smoke, bench, K sweep, ncu launch list + full capture of the blend kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for k in 1 2 4 8; do
  echo "== FWD_K=$k BWD_K=$k"; FAST=1 GAB200_FWD_K=$k GAB200_BWD_K=$k python scripts/quick_timing.py 2>&1 | tail -4
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:blend_ -s 8 -c 4 -o gpurun_out/prof_blend -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
