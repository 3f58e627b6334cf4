#!/usr/bin/env python
"""tools/copy_bench.py -- bare host<->device copy ceiling with N ranks copying at once (torchrun, one rank per GPU).

The e2e leg of bench.py moves 2 GiB host->device and 4 GiB device->host per step and rank through pinned buffers.
This measures what the box can do for exactly that traffic with NOTHING else running: H2D alone, D2H alone, both at
once (two streams, 32 MiB chunks like the library's pipeline, and as single large copies), max over ranks.  The
aggregate at N ranks is the ceiling the e2e number should be read against."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import yagi_b200 as yb

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

NIN, NOUT = 1 << 28, 1 << 29                      # cf32 elements: 2 GiB in, 4 GiB out
hx, hy = yb.PinnedArray(NIN), yb.PinnedArray(NOUT)
hx.array[:] = 1.0
hxt, hyt = torch.from_numpy(hx.array), torch.from_numpy(hy.array)
dx = torch.empty(NIN, dtype=torch.complex64, device=dev)
dy = torch.zeros(NOUT, dtype=torch.complex64, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allmax(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run(name, h2d, d2h, chunk):
    def once():
        if h2d:
            with torch.cuda.stream(s_in):
                for a in range(0, NIN, chunk):
                    dx[a:a + chunk].copy_(hxt[a:a + chunk], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                for a in range(0, NOUT, 2 * chunk):
                    hyt[a:a + 2 * chunk].copy_(dy[a:a + 2 * chunk], non_blocking=True)
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    dt = allmax((time.perf_counter() - t0) / 3)
    nbytes = (8 * NIN if h2d else 0) + (8 * NOUT if d2h else 0)
    barrier()
    if rank == 0:
        print(json.dumps({"copy": name, "ranks": world, "ms": round(dt * 1e3, 2), "GBps_per_rank": round(nbytes / dt / 1e9, 1),
                          "GBps_aggregate": round(world * nbytes / dt / 1e9, 1),
                          "equivalent_e2e_Msps": round(world * NIN / dt / 1e6) if (h2d and d2h) else None}), flush=True)


run("H2D 2 GiB, one copy", True, False, NIN)
run("D2H 4 GiB, one copy", False, True, NOUT)
run("H2D 2 GiB + D2H 4 GiB at once, one copy each", True, True, NIN)
run("H2D 2 GiB + D2H 4 GiB at once, 32 MiB / 64 MiB chunks", True, True, 1 << 22)
hx.close(); hy.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
