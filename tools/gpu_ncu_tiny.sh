#!/bin/bash
# tools/gpu_ncu_tiny.sh -- one full ncu capture of the tiny-M analysis kernel (second launch group = M=16)
mkdir -p gpurun_out
CMD="python tools/bench_generic_small.py"
$CMD > gpurun_out/plain_tiny.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_analysis_tiny -s 12 -c 1 -f -o gpurun_out/prof_tiny $CMD > gpurun_out/ncu_tiny.log 2>&1
tail -n 2 gpurun_out/ncu_tiny.log
