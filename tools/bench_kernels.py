#!/usr/bin/env python
"""tools/bench_kernels.py -- device-resident throughput of the other kernels on the path (not the
headline bench): firpfbch2 synthesis, firpfbch many-stream analysis, batched firfilt, and the
M=1024 round trip of BASELINE config #4.  Prints one JSON line per kernel with its roofline fraction.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import yagi_b200 as yb

PEAK = 6537.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(name, ms, algo_bytes, units, unit_name, extra=None):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "ms": round(ms, 4), "algorithmic_GBps": round(gbs, 1), "peak_GBps": PEAK,
            "frac_of_measured_hbm": round(gbs / PEAK, 4), "M%s_per_s" % unit_name: round(units / (ms * 1e-3) / 1e6, 1)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def randc(n):
    x = torch.empty(n, 2, dtype=torch.float32, device="cuda")
    x.normal_(0, 1)
    return torch.view_as_complex(x)


def main():
    torch.cuda.set_device(0)
    which = sys.argv[1:] or ["synth", "ana1024", "synth1024", "pfbch", "firfilt"]
    if "synth" in which:
        M, m = 256, 7
        K = (1 << 28) // M                      # 2^28 channel samples in, 2^27 samples out
        X = randc(K * M)
        y = torch.empty(K * M // 2, dtype=torch.complex64, device="cuda")
        q = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
        ms = timed(lambda: q.execute_block(X, K, out=y))
        report("firpfbch2 synthesis M=256 m=7 (path %d)" % q.last_path(), ms, 24.0 * K * M // 2, K * M // 2, "samples_out",
               {"kernel_ms": round(float(np.mean(q.kernel_times_ms(16))), 4)})
        del X, y, q
    if "ana1024" in which or "synth1024" in which:
        M, m = 1024, 4
        N = 1 << (int(os.environ.get("YG_LOG2N", "24")))          # BASELINE config #4: N = 2^24
        x = randc(N)
        Y = torch.empty(2 * N, dtype=torch.complex64, device="cuda")
        qa = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
        ms = timed(lambda: qa.execute_block(x, N // (M // 2), out=Y), steps=20)
        kms = round(float(np.mean(qa.kernel_times_ms(4))), 4)
        qa.set_kernel_timing(False)                              # no event records between the launches: what a streaming caller gets
        ms_s = timed(lambda: qa.execute_block(x, N // (M // 2), out=Y), steps=100)
        report("firpfbch2 analysis M=1024 m=4 N=2^%d (path %d)" % (N.bit_length() - 1, qa.last_path()), ms, 24.0 * N, N, "samples_in",
               {"kernel_ms": kms, "ms_back_to_back": round(ms_s, 4), "frac_back_to_back": round(24.0 * N / (ms_s * 1e-3) / 1e9 / PEAK, 4)})
        y = torch.empty(N, dtype=torch.complex64, device="cuda")
        qs = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
        ms = timed(lambda: qs.execute_block(Y, N // (M // 2), out=y), steps=20)
        kms = round(float(np.mean(qs.kernel_times_ms(4))), 4)
        qs.set_kernel_timing(False)
        ms_s = timed(lambda: qs.execute_block(Y, N // (M // 2), out=y), steps=100)
        report("firpfbch2 synthesis M=1024 m=4 N=2^%d (path %d)" % (N.bit_length() - 1, qs.last_path()), ms, 24.0 * N, N, "samples_out",
               {"kernel_ms": kms, "ms_back_to_back": round(ms_s, 4), "frac_back_to_back": round(24.0 * N / (ms_s * 1e-3) / 1e9 / PEAK, 4)})
        del x, Y, y, qa, qs
    if "small" in which:
        for M in (64, 128):
            m, N = 7, 1 << 28
            x = randc(N)
            Y = torch.empty(2 * N, dtype=torch.complex64, device="cuda")
            qa = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
            ms = timed(lambda: qa.execute_block(x, N // (M // 2), out=Y))
            report("firpfbch2 analysis M=%d m=7 N=2^28 (path %d)" % (M, qa.last_path()), ms, 24.0 * N, N, "samples_in")
            del x, qa
            y = torch.empty(N, dtype=torch.complex64, device="cuda")         # synthesis of the 2^28 channel samples: 2^27 out
            qs = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
            No = N // 2
            ms = timed(lambda: qs.execute_block(Y[:N], N // M, out=y[:No]))
            report("firpfbch2 synthesis M=%d m=7 N=2^27 out (path %d)" % (M, qs.last_path()), ms, 24.0 * No, No, "samples_out")
            del Y, y, qs
    if "largeM" in which:
        geoms = ((512, 7), (2048, 4), (4096, 4))
        if os.environ.get("YG_LARGE_GEOM"):                     # e.g. YG_LARGE_GEOM=512:4,512:7
            geoms = tuple(tuple(int(v) for v in g.split(":")) for g in os.environ["YG_LARGE_GEOM"].split(","))
        for M, m in geoms:
            N = 1 << 26
            x = randc(N)
            Y = torch.empty(2 * N, dtype=torch.complex64, device="cuda")
            qa = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
            ms = timed(lambda: qa.execute_block(x, N // (M // 2), out=Y), steps=20)
            report("firpfbch2 analysis M=%d m=%d N=2^26 (path %d)" % (M, m, qa.last_path()), ms, 24.0 * N, N, "samples_in")
            y = torch.empty(N, dtype=torch.complex64, device="cuda")
            qs = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
            ms = timed(lambda: qs.execute_block(Y, N // (M // 2), out=y), steps=20)
            report("firpfbch2 synthesis M=%d m=%d N=2^26 (path %d)" % (M, m, qs.last_path()), ms, 24.0 * N, N, "samples_out")
            del x, Y, y, qa, qs
    if "pfbch" in which:
        M, m, S, n = 64, 7, 512, 1 << 18       # config #5: 512 streams per GPU
        x = randc(S * n)
        y = torch.empty(S * n, dtype=torch.complex64, device="cuda")
        q = yb.FirPfbCh.new_kaiser(yb.ANALYZER, M, m, 60.0, n_streams=S)
        ms = timed(lambda: q.execute_block(x, n // M, out=y), steps=5)
        report("firpfbch analysis M=64 m=7, 512 streams x 2^18", ms, 16.0 * S * n, S * n, "samples_in")
        qs = yb.FirPfbCh.new_kaiser(yb.SYNTHESIZER, M, m, 60.0, n_streams=S)
        ms = timed(lambda: qs.execute_block(x, n // M, out=y), steps=5)
        report("firpfbch synthesis M=64 m=7, 512 streams x 2^18", ms, 16.0 * S * n, S * n, "samples_out")
        del qs, q
        for M in (8, 16, 32):                   # the same stream-sharded workload on the tiny-M kernels
            q = yb.FirPfbCh.new_kaiser(yb.ANALYZER, M, m, 60.0, n_streams=S)
            ms = timed(lambda: q.execute_block(x, n // M, out=y), steps=5)
            report("firpfbch analysis M=%d m=7, 512 streams x 2^18 (path %d)" % (M, q.last_path()), ms, 16.0 * S * n, S * n, "samples_in")
            qs = yb.FirPfbCh.new_kaiser(yb.SYNTHESIZER, M, m, 60.0, n_streams=S)
            ms = timed(lambda: qs.execute_block(x, n // M, out=y), steps=5)
            report("firpfbch synthesis M=%d m=7, 512 streams x 2^18 (path %d)" % (M, qs.last_path()), ms, 16.0 * S * n, S * n, "samples_out")
            del q, qs
        del x, y
    if "firfilt" in which:
        S, n = 1024, 1 << 20                    # BASELINE config #2: 1024 streams x 2^20 samples (8 GiB in, 8 GiB out)
        x = randc(S * n)
        y = torch.empty(S * n, dtype=torch.complex64, device="cuda")
        q = yb.FirFilt.new_kaiser(63, 0.25, 60.0, 0.0, n_streams=S)
        ms = timed(lambda: q.execute_block(x, out=y), steps=5)
        flops = 252.0 * S * n
        report("firfilt_crcf 63 taps, 1024 streams x 2^20 (path %d%s)" % (q.last_path(), ", tensor cores" if q.last_path() == 4 else ""),
               ms, 16.0 * S * n, S * n, "samples", {"useful_f32_TFLOPs": round(flops / (ms * 1e-3) / 1e12, 2)})
        for taps in (127, 255):                 # the same kernel at tap capacities 128 / 256 (FP32-bound, not a BASELINE config)
            q = yb.FirFilt.new_kaiser(taps, 0.25, 60.0, 0.0, n_streams=S)
            ms = timed(lambda: q.execute_block(x, out=y), steps=3)
            report("firfilt_crcf %d taps, 1024 streams x 2^20 (path %d)" % (taps, q.last_path()), ms, 16.0 * S * n, S * n, "samples",
                   {"fp32_TFLOPs": round(4.0 * taps * S * n / (ms * 1e-3) / 1e12, 2)})


if __name__ == "__main__":
    main()
