#!/bin/bash
# tools/gpu_s1ks.sh -- parity + A/B timing of the single-SM M=1024 kernels (analysis and synthesis) against the group
# kernels, then one ncu capture of the synthesis kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "1024 or config4 or single_sm or large or hopping or launch" > gpurun_out/s1ks_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1ks_pytest.log
tail -n 8 gpurun_out/s1ks_pytest.log
for n in 24 26; do for s in 0 1; do
  echo "== YG_LOG2N=$n YG_LARGE_SINGLE_SM=$s"
  YG_LOG2N=$n YG_LARGE_SINGLE_SM=$s timeout 300 python tools/bench_kernels.py ana1024 2>&1 | grep firpfbch2
done; done | tee gpurun_out/s1ks_ab.log
if [ "$1" != "noncu" ]; then
  timeout 120 python tools/pfb_one.py 1024 4 26 3 1 > gpurun_out/pfb_one_s.log 2>&1 && \
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_m1024_synth -c 1 -s 1 -f -o gpurun_out/r02_prof_s1ks \
      python tools/pfb_one.py 1024 4 26 3 1 > gpurun_out/ncu_s1ks.log 2>&1
  tail -n 3 gpurun_out/ncu_s1ks.log
fi
