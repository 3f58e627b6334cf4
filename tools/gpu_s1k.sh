#!/bin/bash
# tools/gpu_s1k.sh -- parity + A/B timing of the single-SM M=1024 analysis kernel against the group kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "1024 or config4 or hopping or large" > gpurun_out/s1k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1k_pytest.log
tail -n 6 gpurun_out/s1k_pytest.log
for n in 24 26; do for s in 0 1; do
  echo "== YG_LOG2N=$n YG_LARGE_SINGLE_SM=$s"
  YG_LOG2N=$n YG_LARGE_SINGLE_SM=$s timeout 300 python tools/bench_kernels.py ana1024 2>&1 | grep analysis
done; done | tee gpurun_out/s1k_ab.log
