#!/usr/bin/env python
"""tools/strong_probe.py -- what a strong-scaled shard costs per step on ONE GPU.

BASELINE config #3 split over G GPUs gives every rank 2^28 / G samples per step; the per-rank step is the same
on 1 or 8 GPUs (no collective), so its efficiency can be read on a single GPU: time K back-to-back steps of
2^28, 2^27, 2^26, 2^25 samples and compare ms_per_step(2^28) / G with ms_per_step(2^28 / G)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import yagi_b200 as yb

M, m = 256, 7
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
x = torch.view_as_complex(torch.randn(1 << 28, 2, device="cuda"))
y = torch.empty(1 << 29, dtype=torch.complex64, device="cuda")
res = {}
for timing in (True, False):
    for lg in (28, 27, 26, 25):
        n = 1 << lg
        q = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
        q.set_kernel_timing(timing)
        nf = n // (M // 2)
        for _ in range(5):
            q.execute_block(x[:n], nf, out=y[: 2 * n])
        torch.cuda.synchronize()
        steps = K * (1 << (28 - lg))
        l0 = yb.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            q.execute_block(x[:n], nf, out=y[: 2 * n])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kms = float(q.kernel_times_ms(64).mean()) if timing else None
        res[(timing, lg)] = ms
        print(json.dumps({"log2_samples": lg, "timing_events": timing, "steps": steps, "ms_per_step": round(ms, 5),
                          "kernel_ms": None if kms is None else round(kms, 5), "launches_per_step": (yb.launch_count() - l0) / steps,
                          "strong_efficiency_vs_2^28": round(res[(timing, 28)] / (1 << (28 - lg)) / ms, 4)}), flush=True)
