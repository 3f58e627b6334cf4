#!/bin/bash
# 8-GPU: headline bench (reference arm + ours with --gather) and the stream-sharded config #5 bench
bash tools/gpu_multi.sh 8 --gather
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR tools/bench_streams.py --steps 20 > gpurun_out/streams_8.log 2>&1; echo "rc=$?" >> gpurun_out/streams_8.log
grep -E "^\{|rc=" gpurun_out/streams_8.log | cut -c1-600
