#!/bin/bash
# tools/gpu_multi_r2.sh N -- round-2 multi-GPU evidence on one box with N GPUs:
#   host topology, bare pinned-copy ceiling at 1, 2, 4 .. N concurrent ranks, bench.py (weak + strong + shard parity) at N,
#   and the two-rank / N-rank GPU tests (NCCL gather, device mismatch).
N=${1:-2}
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; free -g; nvidia-smi topo -m; lscpu | grep -i -E "model name|socket|numa|^cpu\(s\)"; } > gpurun_out/r02_multi_${N}_host.log 2>&1
: > gpurun_out/r02_copy_bench_${N}.log
for R in 1 2 4 8; do
  if [ $R -le $N ]; then
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port $((29520+R))"
    timeout 300 $TR tools/copy_bench.py >> gpurun_out/r02_copy_bench_${N}.log 2>&1
  fi
done
grep -E "^\{" gpurun_out/r02_copy_bench_${N}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 50 --warmup 3 --gather > gpurun_out/r02_bench_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_bench_${N}gpu.log
grep -E "^\{|rc=" gpurun_out/r02_bench_${N}gpu.log | cut -c1-6000
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "nccl_gather or device_mismatch" > gpurun_out/r02_pytest_multi_${N}.log 2>&1; tail -3 gpurun_out/r02_pytest_multi_${N}.log
