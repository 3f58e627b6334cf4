#!/bin/bash
# tools/gpu_ncu_s1.sh -- ncu --set full of the single-SM M=1024 kernels at config #4 size (N = 2^24), each after a plain run
mkdir -p gpurun_out
for cfg in "24 0 k_m1024_fused s1k_ana24" "24 1 k_m1024_synth s1k_syn24"; do      # at most 64 MiB come back per call
  set -- $cfg
  timeout 120 python tools/pfb_one.py 1024 4 $1 3 $2 > gpurun_out/pfb_one.log 2>&1 || { echo "plain run failed"; cat gpurun_out/pfb_one.log; exit 1; }
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$3 -c 1 -s 1 -f -o gpurun_out/r02_prof_$4 \
      python tools/pfb_one.py 1024 4 $1 3 $2 > gpurun_out/ncu_$4.log 2>&1
  tail -n 1 gpurun_out/ncu_$4.log
done
