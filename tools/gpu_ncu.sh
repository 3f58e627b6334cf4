#!/bin/bash
# tools/gpu_ncu.sh -- launch list + full captures of the fused kernels (1 GPU).  Logs to gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_analysis_fused -s 3 -c 1 -f -o gpurun_out/prof_analysis $CMD > gpurun_out/ncu_full.log 2>&1
CMD2="python tools/bench_kernels.py synth"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_synthesis_fused -s 3 -c 1 -f -o gpurun_out/prof_synthesis $CMD2 > gpurun_out/ncu_full2.log 2>&1
tail -n 3 gpurun_out/ncu_full.log gpurun_out/ncu_full2.log
ls -la gpurun_out | tail -8
