#!/bin/bash
# tools/gpu_r2_check.sh -- round-2 tests, randomized stress of the many-stream objects (incl. the tensor-core firfilt), kernel
# table, and the acquire-fence A/B of the large-M kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r02_pytest_round2.log 2>&1; tail -n 4 gpurun_out/r02_pytest_round2.log
timeout 200 python tools/stress_parity.py 90 7 streams > gpurun_out/r02_stress_streams.log 2>&1; echo "rc=$?" >> gpurun_out/r02_stress_streams.log; tail -n 3 gpurun_out/r02_stress_streams.log
timeout 200 python tools/tc_probe.py timing > gpurun_out/tc_timing.log 2>&1; grep "63 taps" gpurun_out/tc_timing.log
timeout 600 python tools/bench_kernels.py ana1024 largeM > gpurun_out/r02_large_fence1.log 2>&1; grep -E "1024|512" gpurun_out/r02_large_fence1.log | cut -c1-200
touch yagi_b200/csrc/firpfbch2_large.cu
YG_NVCC_EXTRA=-DYG_LARGE_ACQUIRE_FENCE=0 python -m yagi_b200.build > gpurun_out/rebuild.log 2>&1
timeout 600 python tools/bench_kernels.py ana1024 largeM > gpurun_out/r02_large_fence0.log 2>&1; grep -E "1024|512" gpurun_out/r02_large_fence0.log | cut -c1-200
