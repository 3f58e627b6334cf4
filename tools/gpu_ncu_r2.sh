#!/bin/bash
# tools/gpu_ncu_r2.sh -- round 2: full GPU test-suite, launch list of the bench, one full capture of the metric kernel and of
# the tensor-core firfilt kernel (1 GPU).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -n 14 gpurun_out/r02_pytest_gpu.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-sustained"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firpfbch2_analysis_fused -s 3 -c 1 -f -o gpurun_out/r02_prof_analysis $CMD > gpurun_out/ncu_a.log 2>&1
python tools/tc_one.py 1 63 3 > gpurun_out/plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_firfilt_tc -s 1 -c 1 -f -o gpurun_out/r02_prof_firfilt_tc python tools/tc_one.py 1 63 3 > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep; tail -n 2 gpurun_out/ncu_a.log gpurun_out/ncu_tc.log
