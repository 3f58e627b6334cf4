#!/bin/bash
# tools/gpu_sweep1024.sh -- single-SM M=1024 kernels over call sizes 2^21 .. 2^27 (time = a + b N fit)
mkdir -p gpurun_out
for n in 21 22 23 24 25 26 27; do
  YG_LOG2N=$n timeout 300 python tools/bench_kernels.py ana1024 2>&1 | grep '^{'
done | tee gpurun_out/sweep1024.log
