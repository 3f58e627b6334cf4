#!/bin/bash
# tools/gpu_ncu1.sh "<cmd>" <kernel-regex> <outname> -- one full ncu capture of one kernel
mkdir -p gpurun_out
$1 > gpurun_out/plain_$3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$3 $1 > gpurun_out/ncu_$3.log 2>&1
tail -n 2 gpurun_out/ncu_$3.log
