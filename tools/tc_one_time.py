#!/usr/bin/env python
"""tools/tc_one_time.py -- time the tensor-core firfilt on BASELINE config #2 (used for the YG_TC_SEG sweep)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["YG_FIRFILT_TC"] = "1"

import torch

import yagi_b200 as yb

S, N = 1024, 1 << 20
x = torch.view_as_complex(torch.randn(S * N, 2, device="cuda"))
y = torch.empty_like(x)
q = yb.FirFilt.new(yb.fir_design_kaiser(63, 0.25, 60.0, 0.0), n_streams=S)
for _ in range(3):
    q.execute_block(x, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    q.execute_block(x, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("path %d  %.4f ms  %.1f GB/s  frac %.4f" % (q.last_path(), ms, 16.0 * S * N / ms / 1e6, 16.0 * S * N / ms / 1e6 / 6537.6), flush=True)
