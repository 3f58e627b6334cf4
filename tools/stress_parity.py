#!/usr/bin/env python
"""tools/stress_parity.py [seconds] [seed] [streams] -- randomized parity stress of every firpfbch2 path on a GPU
(with `streams`: of the many-stream objects firpfbch and firfilt instead).

Random geometry (every fused size plus a few generic ones), random semi-length, random sequence of call sizes
(tiny, odd, ragged, large), now and then a device buffer that starts on an odd sample; the whole stream is
compared with the CPU oracle (rel-RMS <= 1e-5, per-frame max-abs <= 1e-4 of the output scale).  Complements the
fixed cases of tests/test_gpu_parity.py; prints one line per case and exits non-zero on the first mismatch.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import yagi_b200 as yb
from oracle import pyoracle as po


def rand_c(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def check(name, y, ref, rows_axis=1):
    scale = max(1.0, float(np.abs(ref).max()))
    worst = np.abs(y - ref).max(axis=rows_axis) / scale
    rel = float(np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-30))
    ok = worst.max() <= 1e-4 and rel <= 1e-5
    print("%s rel-RMS=%.2e max=%.2e %s" % (name, rel, float(worst.max()), "ok" if ok else "MISMATCH at row %d" % int(worst.argmax())), flush=True)
    if not ok:
        sys.exit(1)


def stream_objects(rng, budget):
    """firpfbch (critically sampled, many streams) and batched firfilt: random geometry, streams, call cuts."""
    t0 = time.time()
    case = 0
    while time.time() - t0 < budget:
        case += 1
        which = int(rng.integers(0, 3))
        if which == 2:
            # firfilt shapes the tensor-core kernel takes (<= 65 taps, whole 64-sample blocks per segment, >= 2^16 samples)
            h_len = int(rng.choice([1, 2, 17, 33, 62, 63, 64, 65, 66, 96, 97, 98, 127, 128, 160, 161]))
            N = int(rng.choice([4096, 8192, 512 * 37, 65536, 3 * 8192 * 8, 131072]))
            S_ = int(rng.integers(1, max(2, min(300, (1 << 23) // N))))
            h = (rng.standard_normal(h_len) / np.sqrt(h_len)).astype(np.float32)
            x = rand_c(rng, S_ * N).reshape(S_, N)
            q = yb.FirFilt.new(h, n_streams=S_)
            sc = float(rng.choice([1.0, 0.37]))
            q.set_scale(sc)
            cuts = sorted(set([0, N] + [int(c) * 512 for c in rng.integers(0, N // 512 + 1, size=int(rng.integers(0, 3)))]))
            xd = torch.from_numpy(x).cuda()
            ys, paths = [], []
            for a, b in zip(cuts, cuts[1:]):
                if b > a:
                    ys.append(q.execute_block(xd[:, a:b].contiguous()).view(S_, b - a).cpu().numpy())
                    paths.append(q.last_path())
            y = np.concatenate(ys, axis=1)
            ref = np.stack([po.firfilt_crcf(h, x[s], scale=sc) for s in range(S_)])
            check("stream case %3d firfilt(tc) h_len=%2d S=%3d N=%6d paths=%s" % (case, h_len, S_, N, paths), y, ref)
        elif which == 1:
            M = int(rng.choice([64, 64, 64, 16, 5, 32]))
            p = int(rng.integers(1, 17))
            S_ = int(rng.integers(1, 40)) if M != 64 else int(rng.choice([1, 3, 4, 9, 37, 130, 700]))
            Q = int(rng.integers(3, 120))                      # frames per stream
            synth = bool(rng.integers(0, 2))
            h = rng.standard_normal(M * p).astype(np.float32)
            x = rand_c(rng, S_ * Q * M).reshape(S_, Q * M)
            q = yb.FirPfbCh.new(yb.SYNTHESIZER if synth else yb.ANALYZER, M, p, h, n_streams=S_)
            cuts = sorted(set([0, Q] + [int(c) for c in rng.integers(0, Q + 1, size=int(rng.integers(0, 4)))]))
            y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, a * M: b * M])).reshape(S_, -1)
                                for a, b in zip(cuts, cuts[1:]) if b > a], axis=1)
            ref = np.stack([po.FirPfbCh.new(po.SYNTHESIZER if synth else po.ANALYZER, M, p, h).execute_block(x[s]) for s in range(S_)])
            check("stream case %3d firpfbch %s M=%2d p=%2d S=%3d Q=%3d calls=%d" % (case, "syn" if synth else "ana", M, p, S_, Q, len(cuts) - 1), y, ref)
        else:
            h_len = int(rng.choice([1, 7, 23, 63, 64, 65, 100, 128, 200, 256, 257, 400]))
            S_ = int(rng.integers(1, 6))
            N = int(rng.choice([300, 4096, 5000, 20000, 70000]))
            h = rng.standard_normal(h_len).astype(np.float32)
            x = rand_c(rng, S_ * N).reshape(S_, N)
            q = yb.FirFilt.new(h, n_streams=S_)
            sc = float(rng.choice([1.0, 0.37]))
            q.set_scale(sc)
            cuts = sorted(set([0, N] + [int(c) for c in rng.integers(0, N + 1, size=int(rng.integers(0, 3)))]))
            y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, a:b])).reshape(S_, -1) for a, b in zip(cuts, cuts[1:]) if b > a], axis=1)
            ref = np.stack([po.firfilt_crcf(h, x[s], scale=sc) for s in range(S_)])
            check("stream case %3d firfilt h_len=%3d S=%d N=%5d calls=%d" % (case, h_len, S_, N, len(cuts) - 1), y, ref)
    print("stream objects ok: %d cases" % case)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    if len(sys.argv) > 3 and sys.argv[3] == "streams":
        return stream_objects(rng, budget)
    sizes = [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 6, 24, 48, 100]
    if os.environ.get("YG_STRESS_SIZES"):                     # e.g. YG_STRESS_SIZES=1024 YG_STRESS_MMAX=4: the single-SM kernels only
        sizes = [int(v) for v in os.environ["YG_STRESS_SIZES"].split(",")]
    m_max = int(os.environ.get("YG_STRESS_MMAX", "8"))
    t0 = time.time()
    case = 0
    while time.time() - t0 < budget:
        case += 1
        M = int(rng.choice(sizes))
        synth = bool(rng.integers(0, 2))
        m = int(rng.integers(1, min(m_max, 7 if synth else 8) + 1))
        # total frames: enough for several batches per slab sometimes, small otherwise; bounded by oracle time
        total_samples = int(rng.choice([1 << 16, 1 << 18, 1 << 20, 3 << 20]))
        K = max(40, total_samples // (M // 2))
        if M >= 1024:
            K = min(K, 6000)
        h = rng.standard_normal(2 * M * m).astype(np.float32)
        nin = M if synth else M // 2
        nout = M // 2 if synth else M
        x = rand_c(rng, K * nin)
        otype, gtype = (po.SYNTHESIZER, yb.SYNTHESIZER) if synth else (po.ANALYZER, yb.ANALYZER)
        ref = po.FirPfbCh2.new(otype, M, m, h).execute_block(x).reshape(K, nout)
        q = yb.FirPfbCh2.new(gtype, M, m, h)
        # random cuts
        cuts = [0]
        while cuts[-1] < K:
            kind = rng.integers(0, 4)
            step = int([rng.integers(1, 8), rng.integers(30, 300), rng.integers(300, 5000), rng.integers(5000, 200000)][kind])
            cuts.append(min(K, cuts[-1] + step))
        outs, paths = [], set()
        for a, b in zip(cuts, cuts[1:]):
            seg = x[a * nin: b * nin]
            if rng.integers(0, 6) == 0:                       # device buffer starting on an odd sample
                buf = torch.zeros(seg.size + 1, dtype=torch.complex64, device="cuda")
                buf[1:] = torch.from_numpy(seg).cuda()
                outs.append(q.execute_block(buf[1:]).cpu().numpy())
            else:
                outs.append(q.execute_block(seg))
            paths.add(q.last_path())
        y = np.concatenate(outs).reshape(K, nout)
        scale = max(1.0, float(np.abs(ref).max()))
        per_frame = np.abs(y - ref).max(axis=1) / scale
        rel = float(np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-30))
        ok = per_frame.max() <= 1e-4 and rel <= 1e-5
        print("case %3d %s M=%4d m=%d K=%6d calls=%3d paths=%s rel-RMS=%.2e max=%.2e %s" % (
            case, "syn" if synth else "ana", M, m, K, len(cuts) - 1, sorted(paths), rel, float(per_frame.max()),
            "ok" if ok else "MISMATCH at frame %d" % int(per_frame.argmax())), flush=True)
        if not ok:
            sys.exit(1)
    print("stress ok: %d cases in %.0f s" % (case, time.time() - t0))


if __name__ == "__main__":
    main()
