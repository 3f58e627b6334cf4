#!/usr/bin/env python
"""tools/footprint_probe.py -- does the achievable copy bandwidth depend on the footprint?  torch copy_ (the method behind
MEASURED_PEAKS.json hbm_gbs) at 2 x 1 GiB ... 2 x 8 GiB, then both firfilt kernels at several stream counts / lengths."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import yagi_b200 as yb


def t_ms(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for gib in (1, 2, 4, 8):
    n = gib << 27                                   # complex64 elements
    a = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
    b = torch.empty_like(a)
    ms = t_ms(lambda: b.copy_(a))
    print(json.dumps({"copy": "%d GiB -> %d GiB" % (gib, gib), "ms": round(ms, 4), "GBps": round(2 * n * 8 / ms / 1e6, 1)}), flush=True)
    del a, b

h = yb.fir_design_kaiser(63, 0.25, 60.0, 0.0)
for S, lg in ((1024, 20), (256, 20), (1024, 18), (256, 22), (4096, 18)):
    n = S << lg
    x = torch.view_as_complex(torch.randn(n, 2, device="cuda"))
    y = torch.empty_like(x)
    for tc in ("0", "1"):
        os.environ["YG_FIRFILT_TC"] = tc
        q = yb.FirFilt.new(h, n_streams=S)
        ms = t_ms(lambda: q.execute_block(x, out=y))
        print(json.dumps({"firfilt": "63 taps, %d streams x 2^%d, path %d" % (S, lg, q.last_path()), "ms": round(ms, 4),
                          "GBps": round(16.0 * n / ms / 1e6, 1)}), flush=True)
    del x, y
