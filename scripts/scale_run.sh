#!/bin/bash
# usage: scale_run.sh N   -> bench at N GPUs with both gradient collectives
N=$1
for c in nccl nvls; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 5 --collective $c 2> gpurun_out/scale_${N}_$c.err > gpurun_out/scale_${N}_$c.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/scale_${N}_$c.json').read().strip().splitlines()[-1]); print($N, d['config']['grad_collective'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'warm', round(d['warm_l2']['value'],1), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/scale_${N}_$c.err
done
