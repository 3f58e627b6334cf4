#!/usr/bin/env python
"""tools/tc_probe.py -- A/B of the two firfilt kernels (<= 65 taps): register-blocked FFMA2 (path 2) vs tcgen05 3xTF32
Toeplitz GEMM (path 4).  Parity of both against an f64 convolution on a small case, then BASELINE config #2 timing."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import yagi_b200 as yb

PEAK = 6537.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def make(h, S, tc):
    os.environ["YG_FIRFILT_TC"] = "1" if tc else "0"
    return yb.FirFilt.new(h, n_streams=S)


def parity(S, N, taps, cuts=None, scale=1.0):
    rng = np.random.default_rng(S * 131 + N + taps)
    h = rng.standard_normal(taps).astype(np.float32)
    x = (rng.standard_normal((S, N)) + 1j * rng.standard_normal((S, N))).astype(np.complex64)
    ref = np.stack([np.convolve(x[s].astype(np.complex128), h.astype(np.float64))[:N] for s in range(S)]) * scale
    out = {}
    for tc in (0, 1):
        q = make(h, S, tc)
        q.set_scale(scale)
        xd = torch.from_numpy(x).cuda()
        cs = cuts or [0, N]
        ys = []
        paths = []
        for a, b in zip(cs, cs[1:]):
            ys.append(q.execute_block(xd[:, a:b].contiguous()).view(S, b - a))
            paths.append(q.last_path())
        torch.cuda.synchronize()
        y = torch.cat(ys, dim=1).cpu().numpy()
        d = y - ref
        rel = float(np.sqrt((np.abs(d) ** 2).sum() / (np.abs(ref) ** 2).sum()))
        out[tc] = (rel, float(np.abs(d).max()), paths)
    print(json.dumps({"case": "S=%d N=%d taps=%d cuts=%s" % (S, N, taps, cuts), "ffma2": out[0], "tcgen05": out[1]}), flush=True)
    return out


def timing(S, N, taps, steps=20):
    h = yb.fir_design_kaiser(taps, 0.25, 60.0, 0.0)
    x = torch.view_as_complex(torch.randn(S * N, 2, device="cuda"))
    y = torch.empty_like(x)
    for tc in (0, 1):
        q = make(h, S, tc)
        for _ in range(3):
            q.execute_block(x, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            q.execute_block(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = 16.0 * S * N / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "firfilt %d taps, %d streams x 2^%d, path %d" % (taps, S, int(np.log2(N)), q.last_path()),
                          "ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 4),
                          "Gsps": round(S * N / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "timing"]
    if "parity" in what:
        parity(64, 1024, 63)
        parity(64, 4096, 63, cuts=[0, 128, 1000, 1002, 4096], scale=0.5)
        parity(100, 2000, 40)
        parity(7, 5000, 65)
        parity(130, 1 << 14, 1)
        parity(256, 1 << 15, 63)
    if "timing" in what:
        timing(1024, 1 << 20, 63)
        timing(1024, 1 << 20, 33)
        timing(1024, 1 << 20, 17)
