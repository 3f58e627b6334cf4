#!/bin/bash
# tools/gpu_round.sh [nokernels] -- one gpurun call: GPU parity tests, smoke, bench (both arms), kernel benches.  Logs to gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_ref.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
if [ "$1" != "nokernels" ]; then
timeout 900 python tools/bench_kernels.py synth small ana1024 largeM pfbch firfilt > gpurun_out/bench_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/bench_kernels.log
fi
tail -n 16 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/smoke.log; tail -c 600 gpurun_out/bench_ref.log
tail -c 5000 gpurun_out/bench.log; if [ "$1" != "nokernels" ]; then cat gpurun_out/bench_kernels.log; fi
