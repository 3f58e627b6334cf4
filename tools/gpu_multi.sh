#!/bin/bash
# tools/gpu_multi.sh N -- bench.py on N GPUs of one box (torchrun), our arm then the reference arm.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${N}_gpus.log 2>&1; nproc >> gpurun_out/multi_${N}_gpus.log; free -g >> gpurun_out/multi_${N}_gpus.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/multi_${N}.log 2>&1
echo "rc=$?" >> gpurun_out/multi_${N}.log
tail -c 2500 gpurun_out/multi_${N}.log
