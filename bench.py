#!/usr/bin/env python
"""bench.py -- firpfbch2 analysis throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W           # our arm (CUDA, via the C ABI)
    python bench.py --impl reference --gpus N ...           # reference arm: the CPU path on host cores

Workload (`config.workload`): BASELINE config #3 -- firpfbch2_crcf analysis, M=256, m=7, Kaiser As=60,
2^28 cf32 input samples per GPU, synthetic complex noise plus tones generated on the device.
A "step" is one pass of the hot path over that block.  Multi-GPU = time-block sharding with
replicated filter history (no data-path collective; weak scaling: every rank owns one 2^28-sample
time block of one long stream, primed with the (4m-1)M/2-sample halo that precedes it).

value   : whole-job input Msamples/s, inputs resident in HBM (CUDA events, max over ranks).
e2e     : the same through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside the
          timed region).
roofline: dominant kernel only: 24 algorithmic bytes per input sample (8 read + 16 written) over
          the kernel's mean duration (CUDA events on its stream, recorded by the library around
          that kernel during the timed steps), against MEASURED_PEAKS.json hbm_gbs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, SEMI, AS = 256, 7, 60.0
BYTES_PER_SAMPLE = 24.0            # 8 B read + 16 B written per input sample
METRIC = "firpfbch2 analysis Msps/GPU at M=256, % HBM roofline, 1/2/4/8 GPUs"
UNIT = "Msps"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
# dram__bytes_read.sum + dram__bytes_write.sum of one fused-kernel launch on this workload, from the
# `ncu --set full` capture summarised in profiles/r01_ncu_analysis.txt (2.152 GB + 4.237 GB)
NCU_TRAFFIC_BYTES_2P28 = 2.151935e9 + 4.236829e9


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_path(n_per_thread: int, passes: int, threads: int):
    """One oracle firpfbch2 analyser per host thread, each over its own n_per_thread-sample slice."""
    import numpy as np
    import stimulus
    from oracle import pyoracle as po
    base = stimulus.noise_plus_tones(0, n_per_thread, M)
    x = np.tile(base, threads)
    secs = po.bench_firpfbch2_analysis(M, SEMI, AS, x, n_per_thread, threads, passes)
    if secs <= 0:
        raise RuntimeError("oracle bench failed")
    total = float(n_per_thread) * threads * passes
    return total / secs / 1e6, secs


def reference_arm(args):
    """Reference arm: the reference's CPU structure (oracle port -- yagi cannot be built here and has
    no channelizer) with all host threads; each step is a bounded sample of the same workload."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = cpu_threads()
    n_per_thread = 1 << 21
    passes = 1
    vals, times = [], []
    for _ in range(max(args.warmup, 1)):
        run_cpu_path(n_per_thread, passes, threads)
    t0 = time.time()
    for _ in range(args.steps):
        v, s = run_cpu_path(n_per_thread, passes, threads)
        vals.append(v); times.append(s)
        if time.time() - t0 > 150:
            break
    value = statistics.median(vals)
    sample = "%d threads x 2^21 samples of the M=256 m=7 workload per step (%d steps run)" % (threads, len(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.log2_samples),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C restatement of yagi's CPU structure (per-branch Window + scalar dotprod + one FFT per frame); yagi itself "
                "has no channelizer (src/multichannel/mod.rs is empty) and cannot be compiled in this image",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus: int, log2_samples: int):
    return {
        "workload": "BASELINE config #3: firpfbch2_crcf analysis M=256 m=7 Kaiser As=60, 2^%d cf32 samples per GPU, "
                    "time-block sharded with replicated filter history" % log2_samples,
        "M": M, "m": SEMI, "As": AS, "samples_per_gpu": 1 << log2_samples, "n_gpus": n_gpus,
        "sharding": "time blocks, halo (4m-1)*M/2 = %d samples per boundary, no collective" % ((4 * SEMI - 1) * M // 2),
        "l2_policy": "inputs (2 GiB) and outputs (4 GiB) per step exceed the 126 MB L2; no flush needed",
    }


# ----------------------------------------------------------------------------- device stimulus
def make_block(torch, device, n: int, t0: int, seed: int):
    """cf32 noise plus tones for samples [t0, t0+n) on the device, generated in chunks."""
    x = torch.empty(n, dtype=torch.complex64, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    chunk = 1 << 24
    xr = torch.view_as_real(x)
    tones = ((1.0, 3.0 / M, 0.0), (0.5, -17.25 / M, 0.7), (0.25, 0.123, 1.9), (0.1, -0.377, 2.6))
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        xr[a:b].normal_(0.0, 0.1, generator=g)
        t = torch.arange(t0 + a, t0 + b, device=device, dtype=torch.float64)
        for amp, f, ph in tones:
            phase = torch.remainder(t * f, 1.0) * (2.0 * 3.141592653589793) + ph
            xr[a:b, 0] += (amp * torch.cos(phase)).float()
            xr[a:b, 1] += (amp * torch.sin(phase)).float()
        del t
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-samples", type=int, default=28, help="input samples per GPU per step (default 2^28)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL all-gather of the outputs (off the hot path)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import yagi_b200 as yb

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # host placement for the e2e leg: pinned buffers local to this GPU's PCIe root (yagi_b200/_numa.py)
    numa_node = yb.bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    N = 1 << args.log2_samples
    K, W = args.steps, args.warmup
    n_frames = N // (M // 2)
    halo_len = (4 * SEMI - 1) * M // 2

    # this rank's time block of the long stream, plus the halo that precedes it
    x = make_block(torch, device, N, rank * N, seed=0x5EED0001 + rank)
    y = torch.empty(2 * N, dtype=torch.complex64, device=device)
    q = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, SEMI, AS)
    if rank > 0:
        halo = make_block(torch, device, N, (rank - 1) * N, seed=0x5EED0001 + rank - 1)[N - halo_len:].cpu().numpy()
        q.set_state(halo, 0)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        q.execute_block(x, n_frames, out=y)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(K):
        q.execute_block(x, n_frames, out=y)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1)
    path = q.last_path()
    ktimes = q.kernel_times_ms(min(K, 64))
    kernel_ms = float(np.mean(ktimes)) if len(ktimes) else float("nan")

    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / K
    value = (N * world) / (ms_per_step * 1e-3) / 1e6

    # ---- optional: NCCL all-gather of a slice of the per-channel outputs (off the hot path, timed separately)
    gather = None
    if args.gather and world > 1:
        from yagi_b200.gather import all_gather_frames
        gf = min(n_frames, 1 << 15)                       # 64 MiB per rank
        yl = y[: gf * M]
        all_gather_frames(yl, [gf] * world, M)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            all_gather_frames(yl, [gf] * world, M)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / 5
        gather = {"frames_per_rank": gf, "bytes_per_rank": gf * M * 8, "ms": gms,
                  "busbw_GBps": gf * M * 8 * (world - 1) / (gms * 1e-3) / 1e9, "backend": "nccl all_gather_into_tensor"}

    # ---- e2e: host pinned buffers through the host-pointer C-ABI call
    e2e = None
    if not args.no_e2e:
        try:
            hx = yb.PinnedArray(N)
            hy = yb.PinnedArray(2 * N)
            hx.array[:] = x.cpu().numpy()
            q2 = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, SEMI, AS)
            q2.execute_block(hx.array, n_frames, out=hy.array)          # warm-up (allocates staging)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                q2.execute_block(hx.array, n_frames, out=hy.array)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.e2e_steps
            td = torch.tensor([dt], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dt = float(td.item())
            # result check of the e2e path against the device-resident path (same input, fresh state)
            e2e = {"value": (N * world) / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 16 * N,
                   "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "numa_node_rank0": numa_node,
                   "api": "yg_firpfbch2_crcf_execute_block (host pointers, pinned, chunked 3-stream pipeline)"}
            hx.close(); hy.close()
        except Exception as exc:  # report, do not hide
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_PER_SAMPLE * N / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, args.log2_samples),
            "per_gpu_msps": value / world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_2P28 if (path == 2 and args.log2_samples == 28) else None,
                         "traffic_source": "ncu --set full, profiles/r01_ncu_analysis.txt", "peak_source": peak_src, "kernel": "firpfbch2 analysis (%s)" % ("fused" if path == 2 else "generic"),
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * N},
            "clocks": clocks,
            "gpu_launches": 2 * K,
            "kernel_path": path,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if gather is not None:
            line["output_gather"] = gather
        if not args.no_cpu and world == 1:
            threads = cpu_threads()
            npt = 1 << 22
            v, secs = run_cpu_path(npt, 2, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d threads x 2 passes x 2^22 samples each of the same workload (%.1f s wall)" % (threads, secs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
