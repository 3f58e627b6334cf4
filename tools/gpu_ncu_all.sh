#!/bin/bash
# tools/gpu_ncu_all.sh -- launch list of the bench + one full capture per fused kernel (1 GPU).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_analysis_fused -s 3 -c 1 -f -o gpurun_out/prof_analysis $CMD > gpurun_out/ncu_a.log 2>&1
for pair in "synth:k_firpfbch2_synthesis_fused:prof_synthesis" "pfbch:k_firpfbch_analysis_fused:prof_firpfbch" "firfilt:k_firfilt_fast:prof_firfilt"; do
  IFS=: read arg kern out <<< "$pair"
  python tools/bench_kernels.py $arg > gpurun_out/plain_$out.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -f -o gpurun_out/$out python tools/bench_kernels.py $arg > gpurun_out/ncu_$out.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
