#!/bin/bash
# tools/gpu_exp.sh -- A/B experiments on kernel variants (YG_FAST_VARIANT), burst (20 steps) and sustained (1000).
mkdir -p gpurun_out
for v in ${VARIANTS:-1}; do
  echo "=== variant $v" >> gpurun_out/exp.log
  YG_FAST_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or config3 or fft_stage" >> gpurun_out/exp.log 2>&1
  for k in 20 20 1000; do
    YG_FAST_VARIANT=$v timeout 300 python bench.py --steps $k --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('steps', d['steps'], 'value', round(d['value']), 'ms_per_step', round(d['ms_per_step'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 4), d['clocks'])
    else: print(l, end='')
" >> gpurun_out/exp.log
  done
done
cat gpurun_out/exp.log
