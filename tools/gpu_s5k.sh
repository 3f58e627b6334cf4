#!/bin/bash
# tools/gpu_s5k.sh -- parity + A/B timing of the single-SM M=512 kernels against the group kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "single_sm or large or 512 or config4 or write_outside or hopping or dft_stage" > gpurun_out/s5k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5k_pytest.log
tail -n 8 gpurun_out/s5k_pytest.log
for s in 0 1; do
  echo "== YG_LARGE_SINGLE_SM=$s"
  YG_LARGE_SINGLE_SM=$s timeout 300 python tools/bench_kernels.py largeM 2>&1 | grep "M=512"
done | tee gpurun_out/s5k_ab.log
