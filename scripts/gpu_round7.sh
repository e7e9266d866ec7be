#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), '| warm', round(d['warm_l2']['value'],1), round(d['warm_l2']['ms_per_step'],4), '| e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],4))
print('stage_ms', d['stage_ms']); print(d.get('stage_ms_note')); print('host', d['host_us_in_forward'], 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/bench.err
