#!/bin/bash
# tools/gpu_quick.sh "<pytest -k expr>" "<bench_kernels args>" [bench steps] -- quick correctness + timing loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$1" > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_pytest.log
timeout 600 python tools/bench_kernels.py $2 > gpurun_out/quick_bench.log 2>&1
if [ -n "$3" ]; then
  for k in $3; do timeout 300 python bench.py --steps $k --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('analysis steps', d['steps'], 'value', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 4), d['clocks'])
" >> gpurun_out/quick_bench.log; done
fi
tail -n 15 gpurun_out/quick_pytest.log; cat gpurun_out/quick_bench.log
