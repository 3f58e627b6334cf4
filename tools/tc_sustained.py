#!/usr/bin/env python
"""tools/tc_sustained.py -- BASELINE config #2 (1024 streams x 2^20, 63 taps) burst (20 steps) and sustained (>= 1.5 s)
timing of both firfilt kernels, with the clocks seen (nvidia-smi), like bench.py does for the metric kernel."""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import yagi_b200 as yb

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6537.6
rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()

S, N = 1024, 1 << 20
x = torch.view_as_complex(torch.randn(S * N, 2, device="cuda"))
y = torch.empty_like(x)
h = yb.fir_design_kaiser(63, 0.25, 60.0, 0.0)
for tc in ("1", "0"):
    os.environ["YG_FIRFILT_TC"] = tc
    q = yb.FirFilt.new(h, n_streams=S)
    for steps in (20, 500):
        for _ in range(3):
            q.execute_block(x, out=y)
        torch.cuda.synchronize()
        time.sleep(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            q.execute_block(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = e0.elapsed_time(e1) / steps
        sel = [r.split(",") for ts, r in rows if t0 <= ts <= t1 + 0.1]
        mhz = sorted(float(r[0]) for r in sel if len(r) >= 3)
        pw = [float(r[1]) for r in sel if len(r) >= 3]
        cap = any("Active" in r[2] for r in sel if len(r) >= 3)
        print(json.dumps({"kernel": "firfilt 63 taps, 1024 x 2^20, path %d" % q.last_path(), "steps": steps, "seconds": round(ms * steps / 1e3, 2),
                          "ms": round(ms, 4), "frac_of_measured_hbm": round(16.0 * S * N / ms / 1e6 / PEAK, 4),
                          "sm_mhz_median": mhz[len(mhz) // 2] if mhz else None, "power_w_max": max(pw) if pw else None, "sw_power_cap": cap}), flush=True)
proc.terminate()
