#!/bin/bash
# Round-end validation on one B200: smoke, GPU parity tests, the bench line (both arms), the training-step bench,
# the ncu launch list of the bench command and --set full captures of the blend and loss kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 60 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 300 python scripts/train_step_bench.py > gpurun_out/train_step.jsonl 2> gpurun_out/train_step.err; echo "train rc=$?"; cat gpurun_out/train_step.jsonl
timeout 100 python scripts/loss_profile.py > gpurun_out/loss_profile.jsonl 2>&1; cat gpurun_out/loss_profile.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
FRAMES=3 ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 4 -c 2 -o gpurun_out/prof_blend3 -f \
    python scripts/one_frame.py > gpurun_out/ncu_blend3.log 2>&1; echo "ncu blend rc=$?"
ITERS=2 WARM=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"ssim_|adam_" -s 2 -c 3 -o gpurun_out/prof_loss -f \
    python scripts/loss_profile.py > gpurun_out/ncu_loss.log 2>&1; echo "ncu loss rc=$?"
