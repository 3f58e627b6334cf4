#!/usr/bin/env python
"""tools/tc_one.py [tc=1] [taps=63] [steps=3] -- a few launches of one firfilt kernel on BASELINE config #2 (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tc = sys.argv[1] if len(sys.argv) > 1 else "1"
taps = int(sys.argv[2]) if len(sys.argv) > 2 else 63
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
os.environ["YG_FIRFILT_TC"] = tc

import torch

import yagi_b200 as yb

S, N = 1024, 1 << 20
x = torch.view_as_complex(torch.randn(S * N, 2, device="cuda"))
y = torch.empty_like(x)
q = yb.FirFilt.new(yb.fir_design_kaiser(taps, 0.25, 60.0, 0.0), n_streams=S)
for _ in range(steps):
    q.execute_block(x, out=y)
torch.cuda.synchronize()
print("path", q.last_path())
