#!/usr/bin/env python
"""tools/bench_generic.py -- device-resident throughput of the generic firpfbch2 kernels (geometries without a fused
kernel: M not a power of two, or m beyond the fused range); their transform is a mixed-radix Stockham."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import yagi_b200 as yb
from tools.bench_kernels import PEAK, randc, timed

GEOMS = ((10, 5), (24, 5), (48, 5), (60, 5), (100, 5), (240, 4), (250, 4), (384, 4), (1000, 4), (3000, 2), (256, 10), (1024, 8))
if os.environ.get("YG_GEN_GEOM"):                         # e.g. YG_GEN_GEOM=48:5,100:5
    GEOMS = tuple(tuple(int(v) for v in g.split(":")) for g in os.environ["YG_GEN_GEOM"].split(","))
LOG2N = int(os.environ.get("YG_LOG2N", "24"))
for M, m in GEOMS:
    K = (1 << LOG2N) // (M // 2)                 # frames
    N = K * (M // 2)
    x = randc(N)
    Y = torch.empty(2 * N, dtype=torch.complex64, device="cuda")
    qa = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
    ms = timed(lambda: qa.execute_block(x, K, out=Y), steps=5, warmup=2)
    print("analysis  M=%4d m=%2d path %d: %8.3f ms  %6.1f Gsps  %.3f of measured HBM peak" % (M, m, qa.last_path(), ms, N / ms / 1e6, 24.0 * N / ms / 1e6 / PEAK), flush=True)
    y = torch.empty(N, dtype=torch.complex64, device="cuda")
    qs = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
    ms = timed(lambda: qs.execute_block(Y, K, out=y), steps=5, warmup=2)
    print("synthesis M=%4d m=%2d path %d: %8.3f ms  %6.1f Gsps  %.3f of measured HBM peak" % (M, m, qs.last_path(), ms, N / ms / 1e6, 24.0 * N / ms / 1e6 / PEAK), flush=True)
    del x, Y, y, qa, qs

# firpfbch (critically sampled) at sizes without a fused kernel: 64 streams
for M, m in ((48, 5), (100, 5), (128, 7), (256, 7), (1000, 4), (1024, 4)):
    if os.environ.get("YG_GEN_GEOM"):
        break
    S, n = 64, (1 << LOG2N) // 64 // M * M
    x = randc(S * n)
    y = torch.empty(S * n, dtype=torch.complex64, device="cuda")
    for name, t in (("analysis ", yb.ANALYZER), ("synthesis", yb.SYNTHESIZER)):
        q = yb.FirPfbCh.new_kaiser(t, M, m, 60.0, n_streams=S)
        ms = timed(lambda: q.execute_block(x, n // M, out=y), steps=5, warmup=2)
        print("firpfbch %s M=%4d m=%2d path %d: %8.3f ms  %6.1f Gsps  %.3f of measured HBM peak" % (name, M, m, q.last_path(), ms, S * n / ms / 1e6, 16.0 * S * n / ms / 1e6 / PEAK), flush=True)
        del q
    del x, y
