#!/bin/bash
mkdir -p gpurun_out
for k in 20 200 1000; do
  timeout 300 python bench.py --steps $k --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('steps', d['steps'], 'value', round(d['value']), 'ms_per_step', round(d['ms_per_step'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 4), d['clocks'])
    else: print(l, end='')
" >> gpurun_out/exp2.log
done
nvidia-smi -q -d POWER | grep -iE "power limit|cap|draw" | head -12 >> gpurun_out/exp2.log
cat gpurun_out/exp2.log
