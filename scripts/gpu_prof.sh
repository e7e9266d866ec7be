#!/bin/bash
set -u
mkdir -p gpurun_out
python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
FRAMES=3 ncu --set full --clock-control none --import-source on -k regex:"preprocess|emit_keys|tile_order|face_frame" -s 12 -c 6 -o gpurun_out/prof_small -f \
    python scripts/one_frame.py > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
FRAMES=3 ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 4 -c 2 -o gpurun_out/prof_blend2 -f \
    python scripts/one_frame.py > gpurun_out/ncu_blend2.log 2>&1; echo "ncu blend rc=$?"
ls -la gpurun_out | head -20
