#!/bin/bash
# tools/gpu_ncu_more.sh -- full ncu captures of the tiny-M (M=16) analysis / synthesis kernels and the small-M (M=64) synthesis kernel
mkdir -p gpurun_out
CMD="python tools/bench_generic_small.py"
$CMD > gpurun_out/plain_more1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_analysis_tiny -s 12 -c 1 -f -o gpurun_out/prof_tiny_analysis $CMD > gpurun_out/ncu_m1.log 2>&1
$CMD > gpurun_out/plain_more2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_synthesis_tiny -s 12 -c 1 -f -o gpurun_out/prof_tiny_synthesis $CMD > gpurun_out/ncu_m2.log 2>&1
CMD="python tools/bench_kernels.py small"
$CMD > gpurun_out/plain_more3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_synthesis_small -s 3 -c 1 -f -o gpurun_out/prof_small_synthesis $CMD > gpurun_out/ncu_m3.log 2>&1
tail -n 1 gpurun_out/ncu_m1.log gpurun_out/ncu_m2.log gpurun_out/ncu_m3.log
