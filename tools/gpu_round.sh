#!/bin/bash
# tools/gpu_round.sh -- one gpurun call: GPU parity tests, smoke, bench, optional extras.  Logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.log 2>&1
nproc >> gpurun_out/nvsmi.log; free -g >> gpurun_out/nvsmi.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
tail -c 1500 gpurun_out/bench.log
