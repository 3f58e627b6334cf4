#!/bin/bash
# tools/gpu_ncu_large.sh -- one full ncu capture each of the fused large-M analysis and synthesis kernels (M=1024, N=2^26)
mkdir -p gpurun_out
export YG_LOG2N=26
CMD="python tools/bench_kernels.py ana1024"
$CMD > gpurun_out/plain_large.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_large_fused -s 3 -c 1 -f -o gpurun_out/prof_large_analysis $CMD > gpurun_out/ncu_la.log 2>&1
$CMD > gpurun_out/plain_large2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_large_synth_fused -s 3 -c 1 -f -o gpurun_out/prof_large_synthesis $CMD > gpurun_out/ncu_ls.log 2>&1
cat gpurun_out/plain_large.log; tail -n 2 gpurun_out/ncu_la.log gpurun_out/ncu_ls.log
