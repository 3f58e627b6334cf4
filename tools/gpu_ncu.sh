#!/bin/bash
# tools/gpu_ncu.sh -- launch list + one full capture of the fused kernel (1 GPU).  Logs to gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_analysis_fused -s 3 -c 2 -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/plain.log gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
ls -la gpurun_out
