#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-2900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python scripts/train_step_bench.py > gpurun_out/train_step.jsonl 2> gpurun_out/train_step.err; echo "train rc=$?"; cat gpurun_out/train_step.jsonl; tail -3 gpurun_out/train_step.err
python scripts/fps_sweep.py > gpurun_out/fps_sweep.jsonl 2> gpurun_out/fps_sweep.err; echo "sweep rc=$?"; cat gpurun_out/fps_sweep.jsonl; tail -3 gpurun_out/fps_sweep.err
