#!/usr/bin/env python
"""tools/bench_streams.py -- BASELINE config #5: 4096 independent firpfbch_crcf M=64 m=7 analysers spread over
N GPUs (stream sharding, no halo, no collective).  Launch with torchrun like bench.py; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_streams.py --steps 20
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import yagi_b200 as yb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--log2-samples", type=int, default=18)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    M, m = 64, 7
    # config #5 spreads 4096 streams over 8 GPUs (512 per GPU); with fewer GPUs each rank still takes 512
    # (weak scaling, the per-GPU share of the published configuration)
    mine = yb.stream_shards(args.streams, 8)[rank % 8]
    S, n = len(mine), 1 << args.log2_samples
    x = torch.empty(S * n, 2, dtype=torch.float32, device=dev)
    x.normal_(0, 1)
    x = torch.view_as_complex(x)
    y = torch.empty(S * n, dtype=torch.complex64, device=dev)
    q = yb.FirPfbCh.new_kaiser(yb.ANALYZER, M, m, 60.0, n_streams=S)
    for _ in range(args.warmup):
        q.execute_block(x, n // M, out=y)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        q.execute_block(x, n // M, out=y)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        peak = 6537.6
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        total = S * n * world
        gbs = 16.0 * S * n / (ms * 1e-3) / 1e9
        print(json.dumps({"workload": "BASELINE config #5: firpfbch_crcf analysis M=64 m=7, %d streams x 2^%d samples per GPU, stream-sharded" % (S, args.log2_samples),
                          "n_gpus": world, "ms_per_step": ms, "value": total / (ms * 1e-3) / 1e6, "unit": "Msps", "scaling": "weak",
                          "per_gpu_algorithmic_GBps": gbs, "frac_of_measured_hbm": gbs / peak, "steps": args.steps}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
