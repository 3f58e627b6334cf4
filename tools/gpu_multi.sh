#!/bin/bash
# tools/gpu_multi.sh N [extra bench flags] -- bench.py on N GPUs of one box (torchrun): reference arm, then ours.
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${N}_gpus.log 2>&1; nproc >> gpurun_out/multi_${N}_gpus.log; free -g >> gpurun_out/multi_${N}_gpus.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/multi_${N}_ref.log 2>&1; echo "rc=$?" >> gpurun_out/multi_${N}_ref.log
timeout 900 $TR bench.py --gpus $N --steps 50 --warmup 3 "$@" > gpurun_out/multi_${N}.log 2>&1; echo "rc=$?" >> gpurun_out/multi_${N}.log
grep -E "^\{|rc=" gpurun_out/multi_${N}_ref.log | cut -c1-400; grep -E "^\{|rc=" gpurun_out/multi_${N}.log | cut -c1-3000
