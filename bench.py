#!/usr/bin/env python
"""bench.py -- firpfbch2 analysis throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W           # our arm (CUDA, via the C ABI)
    python bench.py --impl reference --gpus N ...           # reference arm: the CPU path on host cores

Workload (`config.workload`): BASELINE config #3 -- firpfbch2_crcf analysis, M=256, m=7, Kaiser As=60,
synthetic complex noise plus tones (a counter-based generator: sample t is a pure function of t, so every
rank can regenerate any part of the one long stream).  A "step" is one pass of the hot path over a block.
Multi-GPU = time-block sharding with replicated filter history, no data-path collective.  Two curves are
measured in the same run:

  weak   : every rank owns 2^28 samples of the stream (rank r: samples [r 2^28, (r+1) 2^28)), primed with
           the (4m-1)M/2-sample halo before them.  This is the top-level `value` (default `--scaling weak`).
  strong : BASELINE configs[2] as written -- 2^28 samples in TOTAL, rank r takes frames [r K/G, (r+1) K/G)
           (yagi_b200.sharding.firpfbch2_time_shards).  Reported under "strong"; `--scaling strong` makes
           it the top-level value instead.

value    : whole-job input Msamples/s, inputs resident in HBM (CUDA events, max over ranks), burst of K steps.
sustained: the same loop run back to back for >= 1 s with its own clocks record (the 1 kW power cap bites here).
e2e      : the same through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside the
           timed region), and its result is compared bit for bit with the device-resident path.
roofline : dominant kernel only: 24 algorithmic bytes per input sample (8 read + 16 written) over
           the kernel's mean duration (CUDA events on its stream, recorded by the library around
           that kernel during the timed steps), against MEASURED_PEAKS.json hbm_gbs.
shard_boundary_parity: on every rank, the first and the last 64 frames of its shard against the CPU oracle
           primed with the regenerated halo (rel-RMS / max-abs, max over ranks) -- continuity across GPU shards
           checked on the hardware that runs them.
"""
from __future__ import annotations

import argparse
import json
import math
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
M2 = M // 2
HALO = (4 * SEMI - 1) * M2         # samples of history a shard needs: (4m-1) M/2
BYTES_PER_SAMPLE = 24.0            # 8 B read + 16 B written per input sample
METRIC = "firpfbch2 analysis Msps/GPU at M=256, % HBM roofline, 1/2/4/8 GPUs"
UNIT = "Msps"
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written by tools/ncu_summary.py --traffic


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


def ncu_traffic(kernel_key: str, log2_samples: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel on this workload, from
    the committed `ncu --set full` capture (profiles/ncu_traffic.json, produced by tools/ncu_summary.py);
    None when no capture of this kernel at this size is on file."""
    try:
        with open(TRAFFIC_JSON) as f:
            e = json.load(f).get("%s@2^%d" % (kernel_key, log2_samples))
        return (float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


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

    def window(self, t0, t1):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
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

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

    def summary(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return self.window(t0, t1)


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
    n_per_thread = 1 << 22                     # per step and thread: 3 passes over 2^22 samples (~0.4 s per step), long enough
    passes = 3                                 # that thread start-up and stragglers do not set the number
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
    sample = "%d threads x %d passes x 2^22 samples of the M=256 m=7 workload per step (%d steps run)" % (threads, passes, len(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.log2_samples, args.scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "C restatement of yagi's CPU structure (per-branch Window + scalar dotprod + one FFT per frame); yagi itself "
                "has no channelizer (src/multichannel/mod.rs is empty) and cannot be compiled in this image",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus: int, log2_samples: int, scaling: str):
    per = ("2^%d cf32 samples per GPU" % log2_samples) if scaling == "weak" else \
          ("2^%d cf32 samples in total, split over the GPUs" % log2_samples)
    return {
        "workload": "BASELINE config #3: firpfbch2_crcf analysis M=256 m=7 Kaiser As=60, %s, "
                    "time-block sharded with replicated filter history" % per,
        "M": M, "m": SEMI, "As": AS, "samples_per_gpu": (1 << log2_samples) if scaling == "weak" else (1 << log2_samples) // n_gpus,
        "n_gpus": n_gpus,
        "sharding": "time blocks, halo (4m-1)*M/2 = %d samples per boundary, no collective" % HALO,
        "l2_policy": "per-GPU inputs (%d MiB) and outputs (%d MiB) per step exceed the 126 MB L2; no flush needed"
                     % (((1 << log2_samples) // (1 if scaling == "weak" else n_gpus)) * 8 >> 20,
                        ((1 << log2_samples) // (1 if scaling == "weak" else n_gpus)) * 16 >> 20),
    }


# ----------------------------------------------------------------------------- device stimulus
_TONES = ((1.0, 3.0 / M, 0.0), (0.5, -17.25 / M, 0.7), (0.25, 0.123, 1.9), (0.1, -0.377, 2.6))


def make_block(torch, device, n: int, t0: int):
    """cf32 noise plus tones for samples [t0, t0+n) of the one long stream, generated on the device.
    Counter-based: sample t depends on t only (a 64-bit mix of t gives two 24-bit uniforms), so any rank can
    regenerate any range -- the halo a shard needs is the tail of its predecessor's range -- whatever the chunking."""
    x = torch.empty(n, dtype=torch.complex64, device=device)
    if n == 0:
        return x
    chunk = 1 << 24
    xr = torch.view_as_real(x)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        t = torch.arange(t0 + a, t0 + b, device=device, dtype=torch.int64)
        h = t * (-7046029254386353131)                      # 0x9E3779B97F4A7C15 as int64 (wraps)
        h = h ^ ((h >> 29) & 0x7FFFFFFFF)
        h = h * (-4658895280553007687)                      # 0xBF58476D1CE4E5B9
        h = h ^ ((h >> 32) & 0xFFFFFFFF)
        re = (h & 0xFFFFFF).to(torch.float32)
        im = ((h >> 24) & 0xFFFFFF).to(torch.float32)
        xr[a:b, 0] = (re * (1.0 / 16777216.0) - 0.5) * 0.35          # uniform, sigma ~ 0.1
        xr[a:b, 1] = (im * (1.0 / 16777216.0) - 0.5) * 0.35
        td = t.to(torch.float64)
        for amp, f, ph in _TONES:
            phase = torch.remainder(td * f, 1.0) * (2.0 * math.pi) + ph
            xr[a:b, 0] += (amp * torch.cos(phase)).float()
            xr[a:b, 1] += (amp * torch.sin(phase)).float()
        del t, td, h, re, im
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="which curve is the top-level value; both are measured and printed when N > 1")
    ap.add_argument("--log2-samples", type=int, default=28, help="input samples per GPU (weak) / in total (strong) per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=1.2)
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
    from yagi_b200.sharding import firpfbch2_time_shards

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v: float) -> float:
        if world == 1:
            return float(v)
        t = torch.tensor([v], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()

    def timed_loop(q, xb, yb_, n_frames, steps, warm):
        """`steps` back-to-back execute_block calls on device pointers between barrier + synchronize on both sides;
        returns (ms for all steps, max over ranks; kernels launched by the library inside the region; wall window)."""
        for _ in range(warm):
            q.execute_block(xb, n_frames, out=yb_)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = yb.launch_count()
        tw0 = time.time()
        ev0.record()
        for _ in range(steps):
            q.execute_block(xb, n_frames, out=yb_)
        ev1.record()
        torch.cuda.synchronize()
        tw1 = time.time()
        l1 = yb.launch_count()
        barrier()
        return allmax(ev0.elapsed_time(ev1)), l1 - l0, (tw0, tw1)

    def boundary_parity(q, xb, yb_, n_frames, t0):
        """First and last 64 frames of this rank's shard [t0, t0 + n_frames M/2) against the CPU oracle primed with
        28 frames of regenerated history (>= the 27 the filter spans, even so the parity matches)."""
        from oracle import pyoracle as po
        from parity import errors as rel_rms_and_max
        pre, chk = 28, 64
        q.reset()
        if t0 > 0:
            q.set_state(make_block(torch, device, HALO, t0 - HALO).cpu().numpy(), 0)
        q.execute_block(xb, n_frames, out=yb_)
        torch.cuda.synchronize()
        worst_rel, worst_abs = 0.0, 0.0
        for f0 in sorted({0, max(0, n_frames - chk)}):
            nf = min(chk, n_frames - f0)
            s0 = t0 + f0 * M2                                # global index of the first new sample of frame f0
            npre = min(pre, s0 // M2)                        # frames of real history available (0 at the stream start)
            seg = make_block(torch, device, (npre + nf) * M2, s0 - npre * M2).cpu().numpy()
            ref = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, SEMI, AS).execute_block(seg)[npre * M:]
            got = yb_[f0 * M:(f0 + nf) * M].cpu().numpy()
            rel, mx = rel_rms_and_max(got, ref)
            worst_rel, worst_abs = max(worst_rel, rel), max(worst_abs, mx)
        return allmax(worst_rel), allmax(worst_abs)

    # ------------------------------------------------------------------ weak leg: 2^log2 samples on every rank
    n_frames = N // M2
    x = make_block(torch, device, N, rank * N)
    y = torch.empty(2 * N, dtype=torch.complex64, device=device)
    q = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, SEMI, AS)
    if rank > 0:
        q.set_state(make_block(torch, device, HALO, rank * N - HALO).cpu().numpy(), 0)
    torch.cuda.synchronize()

    ms_total, launches, wall = timed_loop(q, x, y, n_frames, K, W)
    clocks = sampler.summary(*wall)
    path = q.last_path()
    ktimes = q.kernel_times_ms(min(K, 64))
    kernel_ms = float(np.mean(ktimes)) if len(ktimes) else float("nan")
    ms_per_step = ms_total / K
    weak_value = (N * world) / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak()

    # ------------------------------------------------------------------ strong leg: 2^log2 samples in total
    strong = None
    if world > 1:
        sh = firpfbch2_time_shards(n_frames, M, SEMI, world)[rank]
        ns = sh.n_samples
        xs = make_block(torch, device, ns, sh.sample_begin)
        ys = y[: 2 * ns]
        qs = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, SEMI, AS)
        if rank > 0:
            qs.set_state(make_block(torch, device, HALO, sh.halo_begin).cpu().numpy(), 0)
        torch.cuda.synchronize()
        time.sleep(0.5)                                   # both burst legs start from an idle board
        ms_s, launches_s, wall_st = timed_loop(qs, xs, ys, sh.n_frames, K, W)
        kts = qs.kernel_times_ms(min(K, 64))
        k_s = float(np.mean(kts)) if len(kts) else float("nan")
        # the same loop without the two timing events per call (about a microsecond of stream time each)
        qs.set_kernel_timing(False)
        ms_s2, _, _ = timed_loop(qs, xs, ys, sh.n_frames, K, 1)
        qs.set_kernel_timing(True)
        rel, mx = boundary_parity(qs, xs, ys, sh.n_frames, sh.sample_begin)
        best = min(ms_s, ms_s2)
        strong = {"value": N / (best / K * 1e-3) / 1e6, "unit": UNIT, "samples_total": N, "samples_per_gpu": ns,
                  "ms_per_step": best / K, "ms_per_step_with_timing_events": ms_s / K, "ms_per_step_without_timing_events": ms_s2 / K,
                  "kernel_ms": k_s, "roofline_frac": BYTES_PER_SAMPLE * ns / (k_s * 1e-3) / 1e9 / peak,
                  "gpu_launches": launches_s, "kernel_path": qs.last_path(), "clocks": sampler.summary(*wall_st),
                  "shard_boundary_parity": {"rel_rms": rel, "max_abs": mx, "frames": "first and last 64 of every rank's shard",
                                            "tolerance": {"rel_rms": 1e-5, "max_abs": 1e-4}, "pass": bool(rel <= 1e-5 and mx <= 1e-4)},
                  "l2_policy": "per-GPU input (%d MiB) + output (%d MiB) per step exceed the 126 MB L2" % (ns * 8 >> 20, ns * 16 >> 20)}
        del xs, qs

    # ------------------------------------------------------------------ sustained: the same loop for >= 1 s
    sustained = None
    if not args.no_sustained:
        n_sus = max(K, int(math.ceil(args.sustained_seconds * 1e3 / ms_per_step)))
        ms_sus, _, wall_s = timed_loop(q, x, y, n_frames, n_sus, 0)
        kt = q.kernel_times_ms(64)
        k_sus = float(np.mean(kt)) if len(kt) else float("nan")
        sustained = {"value": (N * world) / (ms_sus / n_sus * 1e-3) / 1e6, "unit": UNIT, "steps": n_sus,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus, "kernel_ms": k_sus,
                     "roofline_frac": BYTES_PER_SAMPLE * N / (k_sus * 1e-3) / 1e9 / peak,
                     "clocks": sampler.summary(*wall_s),
                     "note": "last 64 kernel durations of a >= %.1f s back-to-back run; the 1 kW board power cap (sw_power_cap) "
                             "sets the SM clock here" % args.sustained_seconds}

    weak_parity = boundary_parity(q, x, y, n_frames, rank * N) if world == 1 else None

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
            l0 = yb.launch_count()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                q2.execute_block(hx.array, n_frames, out=hy.array)
            torch.cuda.synchronize()
            dt = allmax((time.perf_counter() - t0) / args.e2e_steps)
            l_e2e = (yb.launch_count() - l0) // args.e2e_steps
            # result check: the host path (chunked, its own launches) against the device-resident path on the same
            # input from the same (reset) state -- the arithmetic per output is identical, so bit for bit
            q2.reset(); q.reset()
            q2.execute_block(hx.array, n_frames, out=hy.array)
            q.execute_block(x, n_frames, out=y)
            torch.cuda.synchronize()
            same, compared = True, 0
            for f0 in (0, n_frames // 3, n_frames - (1 << 14)):
                f0 = max(0, min(f0, n_frames - 1)) & ~1
                nf = min(1 << 14, n_frames - f0)
                a = torch.from_numpy(hy.array[f0 * M:(f0 + nf) * M]).to(device)
                same = same and bool(torch.equal(torch.view_as_real(a), torch.view_as_real(y[f0 * M:(f0 + nf) * M])))
                compared += nf
            e2e = {"value": (N * world) / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 16 * N,
                   "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "numa_node_rank0": numa_node, "gpu_launches_per_step": l_e2e,
                   "result_check": {"bit_equal_to_device_path": same, "frames_compared": compared},
                   "scaling": "weak",
                   "api": "yg_firpfbch2_crcf_execute_block (host pointers, pinned, chunked 3-stream pipeline)"}
            hx.close(); hy.close()
        except Exception as exc:  # report, do not hide
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)}

    sampler.stop()
    if rank == 0:
        achieved = BYTES_PER_SAMPLE * N / (kernel_ms * 1e-3) / 1e9
        kname = "k_firpfbch2_analysis_fused" if path == 2 else "generic"
        traffic, traffic_src = ncu_traffic(kname, args.log2_samples)
        top_strong = args.scaling == "strong" and strong is not None
        line = {
            "metric": METRIC, "value": strong["value"] if top_strong else weak_value, "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": strong["ms_per_step"] if top_strong else ms_per_step,
            "higher_is_better": True, "scaling": "strong" if top_strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, args.log2_samples, "strong" if top_strong else "weak"),
            "per_gpu_msps": (strong["value"] if top_strong else weak_value) / world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "firpfbch2 analysis (%s)" % ("fused" if path == 2 else "generic"),
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * N,
                         "sustained_frac": sustained["roofline_frac"] if sustained else None},
            "clocks": clocks,
            "gpu_launches": strong["gpu_launches"] if top_strong else launches,
            "gpu_launches_source": "yg_launch_count() difference around the timed region (rank 0)",
            "kernel_path": path,
            "weak": {"value": weak_value, "ms_per_step": ms_per_step, "samples_per_gpu": N},
        }
        if sustained is not None:
            line["sustained"] = sustained
        if strong is not None:
            line["strong"] = strong
        if weak_parity is not None:
            line["shard_boundary_parity"] = {"rel_rms": weak_parity[0], "max_abs": weak_parity[1],
                                             "frames": "first and last 64 frames of the block",
                                             "pass": bool(weak_parity[0] <= 1e-5 and weak_parity[1] <= 1e-4)}
        elif strong is not None:
            line["shard_boundary_parity"] = strong["shard_boundary_parity"]
        if e2e is not None:
            line["e2e"] = e2e
        if gather is not None:
            line["output_gather"] = gather
        if not args.no_cpu and world == 1:
            threads = cpu_threads()
            npt, passes = 1 << 22, 6                      # ~20 core-seconds of CPU work on the 16-core box
            v, secs = run_cpu_path(npt, passes, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d threads x %d passes x 2^22 samples each of the same workload (%.1f s wall, %.0f core-seconds)"
                                              % (threads, passes, secs, secs * threads)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
