#!/usr/bin/env python
"""tools/bench_generic_small.py -- device-resident throughput of the generic firpfbch2 kernels at small M
(geometries without a fused kernel): several frames per block, direct M-point DFT."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import yagi_b200 as yb
from tools.bench_kernels import PEAK, randc, timed

for M in (8, 16, 32, 48):
    m = 5
    K = (1 << 26) // (M // 2)                    # frames
    N = K * (M // 2)
    x = randc(N)
    Y = torch.empty(2 * N, dtype=torch.complex64, device="cuda")
    qa = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
    ms = timed(lambda: qa.execute_block(x, K, out=Y), steps=5)
    print("analysis  M=%2d m=5 path %d: %7.3f ms  %6.1f Gsps  %.3f of measured HBM peak" % (M, qa.last_path(), ms, N / ms / 1e6, 24.0 * N / ms / 1e6 / PEAK))
    y = torch.empty(N, dtype=torch.complex64, device="cuda")
    qs = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
    ms = timed(lambda: qs.execute_block(Y, K, out=y), steps=5)
    print("synthesis M=%2d m=5 path %d: %7.3f ms  %6.1f Gsps  %.3f of measured HBM peak" % (M, qs.last_path(), ms, N / ms / 1e6, 24.0 * N / ms / 1e6 / PEAK))
    del x, Y, y, qa, qs
