#!/bin/bash
# ncu sweep over blend kernel variants: duration, instructions, SM active vs elapsed cycles, issue rate.
set -u
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.per_cycle_active,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread
out=gpurun_out/variants.csv; : > $out
for k in 1 2 4; do for b in 0 1; do
  echo "## K=$k BATCH=$b" >> $out
  GAB200_FWD_K=$k GAB200_BWD_K=$k GAB200_FWD_BATCH=$b GAB200_BWD_BATCH=$b \
    ncu --metrics $M --clock-control none -k regex:blend_ -s 4 -c 2 --csv python scripts/one_frame.py 2>/dev/null | grep -E "blend_" >> $out
done; done
python - <<'PY'
import csv
cur=None; rows={}
for line in open('gpurun_out/variants.csv'):
    if line.startswith('##'): cur=line[3:].strip(); continue
    r=next(csv.reader([line]))
    name='fwd' if 'forward' in r[4] else 'bwd'
    rows.setdefault((cur,name),{})[r[-3]]=r[-1]
for (cur,name),m in rows.items():
    print(cur, name, ' '.join(f"{k.split('__')[1][:22]}={v}" for k,v in m.items()))
PY
