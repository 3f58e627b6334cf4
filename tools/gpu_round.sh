#!/bin/bash
# tools/gpu_round.sh -- one gpurun call: GPU parity tests, smoke, bench, kernel benches.  Logs to gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 900 python tools/bench_kernels.py synth small ana1024 largeM pfbch firfilt > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/bench_kernels.log
tail -n 12 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log
tail -c 1800 gpurun_out/bench.log; cat gpurun_out/bench_kernels.log
