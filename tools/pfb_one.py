#!/usr/bin/env python
"""tools/pfb_one.py [M=1024] [m=4] [log2N=26] [steps=3] [synth=0] -- a few launches of one firpfbch2 kernel (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 26)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
synth = (sys.argv[5] == "1") if len(sys.argv) > 5 else False

import torch

import yagi_b200 as yb

K = N // (M // 2)
if synth:
    x = torch.view_as_complex(torch.randn(K * M, 2, device="cuda"))
    y = torch.empty(N, dtype=torch.complex64, device="cuda")
    q = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
else:
    x = torch.view_as_complex(torch.randn(N, 2, device="cuda"))
    y = torch.empty(K * M, dtype=torch.complex64, device="cuda")
    q = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
for _ in range(steps):
    q.execute_block(x, K, out=y)
torch.cuda.synchronize()
print("path", q.last_path())
