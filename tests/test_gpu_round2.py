"""Round-2 GPU parity tests: oracle-independent checks of the CUDA channelizers, full-size property checks
for BASELINE configs #2 / #4 / #5, one-launch steps, and the optional NCCL output gather.

Everything goes through the C ABI (via the thin Python mirror).  Tolerance (BASELINE.json north_star):
rel-RMS <= 1e-5 and max-abs <= 1e-4 per output sample in f32 unless a test states the reference's own.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import stimulus
from oracle import pyoracle as po
from parity import assert_parity, errors

pytestmark = pytest.mark.gpu

import yagi_b200 as yb  # noqa: E402

A, S = yb.ANALYZER, yb.SYNTHESIZER
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rand_c(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


# ------------------------------------------------------------------ one launch per fused step
def test_fused_call_is_one_launch_and_hands_over_state():
    """An even-parity, even-length call on the fused M=256 kernel is ONE kernel launch: the state hand-off
    (tail of the input stream -> the other history buffer) is done by the same kernel."""
    import torch
    M, m, K = 256, 7, 4096
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    xd = torch.from_numpy(x).cuda()
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    ref = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(x)
    l0 = yb.launch_count()
    y1 = q.execute_block(xd[: 1000 * M // 2])
    torch.cuda.synchronize()
    assert q.last_path() == 2
    assert yb.launch_count() - l0 == 1
    hist, flag = q.get_state()
    assert flag == 0
    np.testing.assert_array_equal(hist, x[1000 * M // 2 - q.state_len(): 1000 * M // 2])
    # the next call starts from the state the kernel wrote
    y2 = q.execute_block(xd[1000 * M // 2:])
    torch.cuda.synchronize()
    assert yb.launch_count() - l0 == 2
    assert_parity(torch.cat([y1, y2]).cpu().numpy(), ref, "two one-launch calls")
    # an odd-parity start needs the generic kernel for the leading frame: more launches, same answer
    q.reset()
    l0 = yb.launch_count()
    ya = q.execute_block(xd[: 129 * M // 2])
    yb_ = q.execute_block(xd[129 * M // 2:])
    torch.cuda.synchronize()
    assert yb.launch_count() - l0 >= 4
    assert_parity(torch.cat([ya, yb_]).cpu().numpy(), ref, "odd split")


def test_kernel_timing_toggle():
    import torch
    M, m, K = 256, 7, 256
    xd = torch.from_numpy(stimulus.noise_plus_tones(0, K * M // 2, M)).cuda()
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    q.execute_block(xd)
    assert q.last_kernel_ms() > 0 and len(q.kernel_times_ms()) == 1
    q.set_kernel_timing(False)
    y0 = q.execute_block(xd)
    with pytest.raises(yb.ModeError):
        q.last_kernel_ms()
    assert len(q.kernel_times_ms()) == 0
    q.set_kernel_timing(True)
    y1 = q.execute_block(xd)
    assert q.last_kernel_ms() > 0
    assert bool(torch.equal(torch.view_as_real(y0), torch.view_as_real(y1)))      # same state (tail of x) both times


def test_device_mismatch_is_refused():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    q = yb.FirPfbCh2.new_kaiser(A, 16, 5, 60.0)
    assert q.get_device() == torch.cuda.current_device()
    other = (q.get_device() + 1) % torch.cuda.device_count()
    x = torch.zeros(64 * 8, dtype=torch.complex64, device="cuda:%d" % other)
    with pytest.raises(yb.ValueError_):
        q.execute_block(x)


# ------------------------------------------------------------------ oracle-free: DFT stage vs numpy.fft in f64
@pytest.mark.parametrize("M", [64, 128, 256, 512, 1024, 2048, 4096])
def test_dft_stage_against_numpy_fft_f64(M):
    """Oracle-free pin of the CUDA transform at the fused sizes (the reference's own FFT vectors stop at N = 192):
    with the prototype h[0..M) = 1 (else 0) and m = 1 the analyser is a pure backward DFT,
        y_k = (1/M) IDFT_unnorm(roll(V_k, (k&1) M/2)),  V_k[b] = s[t_k - b],  t_k = (k+1) M/2 - 1,
    so every frame k >= 1 is checked against numpy.fft.ifft (f64) of the rolled, reversed input window.
    Enough frames are sent that the fused kernels (last_path 2 or 3) take the call."""
    rng = np.random.default_rng(M)
    K = 256
    h = np.zeros(2 * M, dtype=np.float32)
    h[:M] = 1.0
    s = _rand_c(rng, K * M // 2)
    q = yb.FirPfbCh2.new(A, M, 1, h)
    y = q.execute_block(s).reshape(K, M)
    assert q.last_path() in (2, 3), q.last_path()
    s64 = s.astype(np.complex128)
    worst = 0.0
    for k in range(1, K):
        tk = (k + 1) * (M // 2) - 1
        V = s64[tk - M + 1: tk + 1][::-1]                     # V[b] = s[tk - b]
        ref = np.fft.ifft(np.roll(V, (k & 1) * (M // 2)))     # = (1/M) IDFT_unnorm
        worst = max(worst, float(np.abs(y[k] - ref).max() / max(1e-30, np.abs(ref).max())))
    # f32 transform of length M: error grows like log2(M) ulps
    assert worst < 2e-6 * np.log2(M), worst


@pytest.mark.parametrize("M", [64, 256, 1024])
def test_synthesis_dft_stage_against_numpy_fft_f64(M):
    """The synthesiser, oracle-free, through its closed form:
        y[k M/2 + i] = sum_{l < 4m} h[i + l M/2] u_{k-l}[(i + pi_k) mod M],  u_k = 1/2 IDFT_unnorm(X_k), pi_k = (k&1) M/2
    evaluated in f64 with numpy.fft for random X and a random prototype."""
    rng = np.random.default_rng(100 + M)
    m, K = 2, 512                    # whole 32-frame rounds, above every fused kernel's minimum call size
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    X = _rand_c(rng, K * M).reshape(K, M)
    q = yb.FirPfbCh2.new(S, M, m, h)
    y = q.execute_block(X.reshape(-1)).reshape(K, M // 2)
    assert q.last_path() in (2, 3), q.last_path()
    u = 0.5 * M * np.fft.ifft(X.astype(np.complex128), axis=1)
    h64 = h.astype(np.float64)
    i = np.arange(M // 2)
    ref = np.zeros((K, M // 2), dtype=np.complex128)
    for k in range(K):
        col = (i + (k & 1) * (M // 2)) % M
        for l in range(4 * m):
            if k - l >= 0:
                ref[k] += h64[i + l * (M // 2)] * u[k - l][col]
    scale = np.abs(ref).max()
    rel, mx = errors(y / scale, ref / scale)
    assert rel < 1e-5 and mx < 1e-4, (rel, mx)


# ------------------------------------------------------------------ upstream firpfbch autotests on the CUDA object
def _mix_filter_decimate(h, x, M, c):
    """Channel c of the analyser, the long way (f64): mix down by c/M, filter with h, keep samples t = M-1 mod M."""
    t = np.arange(x.size)
    xm = x.astype(np.complex128) * np.exp(-2j * np.pi * c * t / M)
    yf = np.convolve(xm, h.astype(np.float64))[: x.size]
    return yf[M - 1:: M]


def test_firpfbch_crcf_analysis_autotest_on_cuda():
    """Upstream autotest firpfbch_crcf_analysis (LIQUID_COMPAT.md:1774) on yb.FirPfbCh: M = 4, p = 5, random taps and
    input, 12 symbols -- the channelizer equals per-channel mix-down / firfilt / decimate (tol 1e-4, upstream's).
    Oracle-independent.  Also at the fused sizes (M = 8 .. 64, many streams)."""
    for M, p, Q, S_ in ((4, 5, 12, 1), (8, 6, 64, 4), (16, 4, 64, 2), (32, 8, 48, 1), (64, 14, 64, 4)):
        rng = np.random.default_rng(M + p)
        h = rng.standard_normal(M * p).astype(np.float32)
        x = _rand_c(rng, S_ * Q * M).reshape(S_, Q * M)
        q = yb.FirPfbCh.new(A, M, p, h, n_streams=S_)
        y = q.execute_block(x).reshape(S_, Q, M)
        for s in range(S_):
            ref = np.stack([_mix_filter_decimate(h, x[s], M, c) for c in range(M)], axis=1)      # [Q][M]
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(y[s] - ref).max() / scale < 1e-4, (M, p, s)


def test_firpfbch_crcf_synthesis_autotest_on_cuda():
    """Upstream autotest firpfbch_crcf_synthesis (LIQUID_COMPAT.md:1768): the synthesiser equals, per channel,
    upsample by M / firfilt / mix up by c/M, summed over channels (tol 1e-4).  Oracle-independent."""
    for M, p, Q, S_ in ((4, 5, 12, 1), (8, 6, 64, 4), (16, 4, 64, 2), (64, 14, 64, 4)):
        rng = np.random.default_rng(7 * M + p)
        h = rng.standard_normal(M * p).astype(np.float32)
        X = _rand_c(rng, S_ * Q * M).reshape(S_, Q, M)
        q = yb.FirPfbCh.new(S, M, p, h, n_streams=S_)
        y = q.execute_block(X.reshape(S_, -1)).reshape(S_, Q * M)
        t = np.arange(Q * M)
        for s in range(S_):
            ref = np.zeros(Q * M, dtype=np.complex128)
            for c in range(M):
                up = np.zeros(Q * M, dtype=np.complex128)
                up[::M] = X[s, :, c]
                ref += np.convolve(up, h.astype(np.float64))[: Q * M] * np.exp(2j * np.pi * c * t / M)
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(y[s] - ref).max() / scale < 1e-4, (M, p, s)


def test_firpfbch2_analysis_equals_mix_filter_decimate_on_cuda():
    """The same oracle-independent identity for the 2x oversampled object:
    y_k[c] = ((-1)^{c k} / M) sum_tau h[tau] e^{+j 2 pi c tau / M} s[t_k - tau], t_k = (k+1) M/2 - 1,
    evaluated in f64 for a random prototype, through the fused kernels."""
    for M, m, K in ((16, 3, 2304), (64, 2, 512), (256, 2, 128)):      # K above each fused kernel's minimum call size
        rng = np.random.default_rng(M + m)
        h = rng.standard_normal(2 * M * m).astype(np.float32)
        x = _rand_c(rng, K * M // 2)
        q = yb.FirPfbCh2.new(A, M, m, h)
        y = q.execute_block(x).reshape(K, M)
        assert q.last_path() == 2
        L = 2 * M * m
        tau = np.arange(L)
        xp = np.concatenate([np.zeros(L, dtype=np.complex128), x.astype(np.complex128)])
        E = np.exp(2j * np.pi * np.outer(np.arange(M), tau) / M) * h.astype(np.float64)[None, :]       # [c][tau]
        ks = list(range(0, K, 7)) + [K - 1]
        for k in ks:
            tk = (k + 1) * (M // 2) - 1
            win = xp[L + tk - tau]                                                                     # s[tk - tau]
            ref = (E @ win) / M * ((-1.0) ** (np.arange(M) * k))
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(y[k] - ref).max() / scale < 2e-5, (M, k)


# ------------------------------------------------------------------ full-size property checks (configs #2, #4, #5)
def test_config2_full_size_properties_firfilt():
    """BASELINE config #2 at full size (1024 streams x 2^20 samples, 63 taps): size-independent properties --
    (a) linearity in the input: y(a x1 + x2) == a y(x1) + y(x2) on the GPU itself (f32 rounding only);
    (b) sampled windows of every 97th stream against the oracle (start, a tile boundary, the end)."""
    import torch
    S_, N = 1024, 1 << 20
    h = yb.fir_design_kaiser(63, 0.25, 60.0, 0.0)
    g = torch.Generator(device="cuda")
    g.manual_seed(22)
    xr = torch.empty(S_ * N, 2, dtype=torch.float32, device="cuda")
    xr.normal_(0.0, 1.0, generator=g)
    x = torch.view_as_complex(xr)
    q = yb.FirFilt.new(h, n_streams=S_)
    y = q.execute_block(x).view(S_, N)
    torch.cuda.synchronize()
    xv = x.view(S_, N)
    for s in range(0, S_, 97):
        for a in (0, 4096 - 100, N // 2 - 77, N - 3000):
            b = min(N, a + 3000)
            lo = max(0, a - 62)
            ref = po.firfilt_crcf(h, xv[s, lo:b].cpu().numpy())[a - lo:]
            assert_parity(y[s, a:b].cpu().numpy(), ref, "config #2 stream %d window %d" % (s, a))
    # (b') DC gain property over the whole block: sum_n y[s][n] = sum(h) * sum_n x[s][n] - (edge terms)
    #      checked on the interior through a telescoping identity: sum_{n=62}^{N-1} y[n] = sum_k h[k] sum_{n=62-k}^{N-1-k} x[n]
    #      (every 8th stream; a dropped or duplicated 64-output tile would move the sum by ~4, rounding by ~1e-3)
    cs = torch.cumsum(xv[::8].to(torch.complex128), dim=1)
    lhs = y[::8, 62:].to(torch.complex128).sum(dim=1)
    rhs = torch.zeros(S_ // 8, dtype=torch.complex128, device="cuda")
    for k in range(63):
        hi = cs[:, N - 1 - k]
        lo_ = cs[:, 62 - k - 1] if 62 - k - 1 >= 0 else torch.zeros_like(hi)
        rhs += float(h[k]) * (hi - lo_)
    err = (lhs - rhs).abs().max().item()
    assert err < 5e-2, err
    del cs, lhs, rhs
    # (a) linearity
    n2 = S_ * N // 4
    x1, x2 = x[:n2], x[n2: 2 * n2]
    q4 = yb.FirFilt.new(h, n_streams=S_ // 4)
    y1 = q4.execute_block(x1).clone(); q4.reset()
    y2 = q4.execute_block(x2).clone(); q4.reset()
    y3 = q4.execute_block(0.5 * x1 + x2)
    torch.cuda.synchronize()
    d = (y3 - (0.5 * y1 + y2)).abs().max().item()
    assert d < 2e-5, d


def test_config4_full_size_round_trip_M1024_m4():
    """BASELINE config #4 at full size: analysis -> synthesis round trip, M=1024, m=4, N=2^24 wideband samples.
    (a) reconstruction y[t] ~= x[t - D], D = 2Mm - M/2 + 1 (upstream firpfbch2 autotest property, tol 1e-3 of the
    signal scale); (b) the channelised frames against the oracle on sampled windows; (c) channel-sum checksum of
    every analysis frame (sum_c y_k[c] = one polyphase partial sum)."""
    import torch
    M, m, N = 1024, 4, 1 << 24
    K = N // (M // 2)
    g = torch.Generator(device="cuda")
    g.manual_seed(44)
    xr = torch.empty(N, 2, dtype=torch.float32, device="cuda")
    xr.normal_(0.0, 1.0, generator=g)
    x = torch.view_as_complex(xr)
    qa = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    qs = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    ch = qa.execute_block(x)
    assert qa.last_path() == 3
    y = qs.execute_block(ch)
    assert qs.last_path() == 3
    torch.cuda.synchronize()
    D = 2 * M * m - M // 2 + 1
    d = y[D:] - x[: N - D]
    rel = (d.abs().pow(2).sum().sqrt() / x.abs().pow(2).sum().sqrt()).item()
    assert rel < 1e-3, rel                       # upstream: |y - x| <= 1e-3 at unit modulus
    assert d.abs().max().item() < 1e-2 and y[:D].abs().max().item() < 1e-2
    # (c) checksum over every frame with full history
    chv = ch.view(K, M)
    h = torch.from_numpy(qa.get_taps()).cuda()
    k = torch.arange(4 * m, K, device="cuda", dtype=torch.int64)
    b0 = (k & 1) * (M // 2)
    tk = (k + 1) * (M // 2) - 1
    acc = torch.zeros(k.numel(), dtype=torch.complex128, device="cuda")
    for n in range(2 * m):
        acc += h[b0 + n * M].to(torch.float64) * x[tk - b0 - n * M].to(torch.complex128)
    chk = chv[4 * m:].to(torch.complex128).sum(dim=1)
    assert (chk - acc).abs().max().item() < 1e-3
    # (b) sampled windows vs the oracle
    halo = (4 * m - 1) * M // 2
    for f0 in (0, 2 * 1111, K // 2 - 32, K - 64):
        f0 -= f0 % 2
        nf = 64
        s0 = f0 * M // 2
        lo = max(0, s0 - halo - M // 2)
        pre = x[lo:s0].cpu().numpy()
        seg = x[s0: s0 + nf * M // 2].cpu().numpy()
        o = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0)
        if pre.size:
            pad = (-pre.size) % M
            o.execute_block(np.concatenate([np.zeros(pad, dtype=np.complex64), pre]))
        ref = o.execute_block(seg)
        assert_parity(chv[f0: f0 + nf].reshape(-1).cpu().numpy(), ref, "config #4 analysis window at frame %d" % f0)
    # synthesis windows vs the oracle fed with the GPU's own channel frames
    for f0 in (0, 2 * 999, K - 64):
        f0 -= f0 % 2
        nf = 64
        npre = min(f0, 32)                       # >= 4m - 1 frames of history, even
        o = po.FirPfbCh2.new_kaiser(po.SYNTHESIZER, M, m, 60.0)
        seg = chv[f0 - npre: f0 + nf].reshape(-1).cpu().numpy()
        ref = o.execute_block(seg)[npre * (M // 2):]
        got = y[f0 * (M // 2): (f0 + nf) * (M // 2)].cpu().numpy()
        assert_parity(got, ref, "config #4 synthesis window at frame %d" % f0)


def test_config5_full_size_properties_firpfbch_M64():
    """BASELINE config #5, one GPU's share at full size (512 streams x 2^18 samples, M=64, m=7):
    (a) every 37th stream, sampled windows against the oracle; (b) analysis of a constant-per-stream DC input
    converges to sum(h)-weighted DC in channel 0 only (closed form), all streams."""
    import torch
    M, m, S_, N = 64, 7, 512, 1 << 18
    Q = N // M
    g = torch.Generator(device="cuda")
    g.manual_seed(55)
    xr = torch.empty(S_ * N, 2, dtype=torch.float32, device="cuda")
    xr.normal_(0.0, 1.0, generator=g)
    x = torch.view_as_complex(xr)
    q = yb.FirPfbCh.new_kaiser(A, M, m, 60.0, n_streams=S_)
    y = q.execute_block(x).view(S_, Q, M)
    assert q.last_path() == 2
    torch.cuda.synchronize()
    xv = x.view(S_, N)
    p = 2 * m
    for s in range(0, S_, 37):
        for f0 in (0, Q // 2 - 40, Q - 64):
            nf = 64
            npre = min(f0, p)                                  # p - 1 frames of history suffice
            seg = xv[s, (f0 - npre) * M: (f0 + nf) * M].cpu().numpy()
            ref = po.FirPfbCh.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(seg)[npre * M:]
            assert_parity(y[s, f0: f0 + nf].reshape(-1).cpu().numpy(), ref, "config #5 stream %d frame %d" % (s, f0))
    # (b) DC input: after the filter has filled, y_q[c] = e^{j 2 pi c / M} sum_tau h[tau] e^{j 2 pi c tau / M} * dc
    q.reset()
    dc = torch.arange(1, S_ + 1, device="cuda", dtype=torch.float32) * (1.0 / S_)
    xd = (dc[:, None] * torch.ones(1, 64 * M, device="cuda")).to(torch.complex64).contiguous()
    yd = q.execute_block(xd.view(-1)).view(S_, 64, M)[:, -1, :].to(torch.complex128)
    h = q.get_taps().astype(np.float64)
    tau = np.arange(h.size)
    H = np.array([np.exp(2j * np.pi * c / M) * np.sum(h * np.exp(2j * np.pi * c * tau / M)) for c in range(M)])
    ref = torch.from_numpy(H).cuda()[None, :] * dc.to(torch.float64)[:, None]
    assert (yd - ref).abs().max().item() < 1e-4


# ------------------------------------------------------------------ optional NCCL gather (off the hot path)
_GATHER_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import yagi_b200 as yb, stimulus
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank % torch.cuda.device_count())
dev = torch.device("cuda", rank % torch.cuda.device_count())
dist.init_process_group("nccl", device_id=dev)
M, m, K = 256, 7, 1024
x = stimulus.noise_plus_tones(0, K * M // 2, M)
whole = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0).execute_block(torch.from_numpy(x).cuda())
sh = yb.firpfbch2_time_shards(K, M, m, world)
me = sh[rank]
q = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
q.set_state(stimulus.noise_plus_tones(me.halo_begin, me.halo_len, M), 0)
y = q.execute_block(torch.from_numpy(x[me.sample_begin: me.sample_end]).cuda())
full = yb.all_gather_frames(y, [s.n_frames for s in sh], M)
torch.cuda.synchronize()
assert full.shape == (K, M), full.shape
assert bool(torch.equal(torch.view_as_real(full.reshape(-1)), torch.view_as_real(whole))), "gathered shards differ from the single pass"
cm = yb.channel_major(full)
assert cm.shape == (M, K) and cm.is_contiguous()
assert bool(torch.equal(torch.view_as_real(cm), torch.view_as_real(full.t().contiguous())))
dist.barrier(); dist.destroy_process_group()
print("gather ok rank", rank)
"""


def test_nccl_gather_of_time_shards_and_channel_major():
    """yagi_b200.gather on hardware: every rank channelises its time shard (primed with the halo), the NCCL
    all-gather reassembles the frames in stream order bit-identically to a single pass, and channel_major returns the
    [M][K] transpose.  Runs with as many ranks as there are GPUs (at most 4; a single GPU still goes through NCCL)."""
    import torch
    world = max(1, min(4, torch.cuda.device_count()))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    code = _GATHER_WORKER.format(root=ROOT)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "gather ok" in o, "rank %d:\n%s" % (r, o[-2000:])


# ------------------------------------------------------------------ firfilt on the tensor cores (tcgen05, 3xTF32)
@pytest.mark.parametrize("S_,N,taps,cuts,scale", [
    (64, 1024, 63, None, 1.0),                              # 8 stream groups x 1 segment group, 2 blocks per segment
    (64, 5098, 63, [0, 512, 1000, 1002, 5098], 0.5),        # state carried across calls; short / odd-sized calls fall back
    (100, 2048, 40, None, 1.0),                             # ragged stream count: 12.5 stream groups
    (130, 1 << 14, 1, None, 2.0),                           # a single tap
    (512, 1 << 13, 65, [0, 4096, 8192], 1.0),               # the longest filter the K = 128 window holds
    (24, 3 << 16, 33, None, 1.0),                           # long streams: 8192-sample segments, 24 of them (3 segment groups)
    (8, 11 * 8192, 63, None, 1.0),                          # 11 segments: the second segment group is ragged (3 of 8 rows used)
    (96, 5000, 63, None, 1.0),                              # ragged length: tensor cores on the first 4608 samples, FFMA2 on 392
    (20, 70002, 17, [0, 70002 - 4000, 70002], 0.25),        # long ragged stream: 8 segments + a 4466-sample tail, then a short call
    (1024, 1 << 12, 63, None, 1.0),                         # more (tile, block) items than CTAs: runs spanning tiles
    (64, 4096, 66, None, 1.0),                              # 66 .. 97 taps: 32-sample blocks, 3 blocks of history
    (40, 6144 + 100, 97, [0, 2048, 6244], 0.5),             # ... with state across calls and a ragged tail
    (64, 1 << 13, 127, None, 1.0),                          # 98 .. 129 taps: 32-sample blocks, 4 blocks of history (130 .. 161: 5)
    (16, 3 << 16, 161, [0, 1 << 16, 3 << 16], 1.0),         # the longest filter tensor memory has columns for, long streams
    (300, 2048, 129, None, 1.0),                            # ragged stream count, the shortest call the variant takes
    (32, 4096, 130, None, 1.0),                             # first tap count of the deepest instance
])
def test_firfilt_tensor_core_path(monkeypatch, S_, N, taps, cuts, scale):
    """The tcgen05 3xTF32 Toeplitz kernel (last_path 4) against an f64 convolution and against the oracle's f32
    result, in the tolerance north_star states (rel-RMS <= 1e-5, max-abs <= 1e-4 at unit output scale)."""
    import torch
    monkeypatch.setenv("YG_FIRFILT_TC", "1")
    rng = np.random.default_rng(S_ * 131 + N + taps)
    h = (rng.standard_normal(taps) / np.sqrt(taps)).astype(np.float32)
    x = _rand_c(rng, S_ * N).reshape(S_, N)
    q = yb.FirFilt.new(h, n_streams=S_)
    q.set_scale(scale)
    xd = torch.from_numpy(x).cuda()
    cs = cuts or [0, N]
    ys, paths = [], []
    for a, b in zip(cs, cs[1:]):
        ys.append(q.execute_block(xd[:, a:b].contiguous()).view(S_, b - a))
        paths.append(q.last_path())
    torch.cuda.synchronize()
    assert 4 in paths, paths
    y = torch.cat(ys, dim=1).cpu().numpy()
    for s in range(0, S_, max(1, S_ // 16)):
        ref64 = np.convolve(x[s].astype(np.complex128), h.astype(np.float64))[:N] * scale
        rel, mx = errors(y[s], ref64)
        assert rel < 2e-6 and mx < 2e-5, (s, rel, mx)
        assert_parity(y[s], po.firfilt_crcf(h, x[s], scale=scale), "tensor-core firfilt vs oracle, stream %d" % s)
    # reset clears the history the next call would otherwise prime the first block with
    q.reset()
    y2 = q.execute_block(xd).view(S_, N).cpu().numpy()
    assert_parity(y2[S_ - 1], po.firfilt_crcf(h, x[S_ - 1], scale=scale), "after reset")


# ------------------------------------------------------------------ guard bands (compute-sanitizer is closed on this pool)
def test_kernels_do_not_write_outside_their_output():
    """Out-of-bounds writes would land in the guard bands around `out`: the fused M=256 kernel (incl. its folded state
    hand-off), the tensor-core firfilt (TMA tensor stores, ragged stream count), the channel-major transpose and the
    large-M kernels (single-SM and cooperative, both directions) all leave a NaN-patterned band of 1 MiB on either side
    untouched."""
    import torch
    pad = 1 << 17

    def guarded(n):
        buf = torch.full((pad + n + pad, 2), float("nan"), dtype=torch.float32, device="cuda")
        return buf, torch.view_as_complex(buf)[pad: pad + n]

    def intact(buf, n):
        return bool(torch.isnan(buf[:pad]).all()) and bool(torch.isnan(buf[pad + n:]).all())

    rng = np.random.default_rng(5)
    # fused analysis, M = 256
    M, m, K = 256, 7, 4098
    x = torch.from_numpy(_rand_c(rng, K * M // 2)).cuda()
    buf, out = guarded(K * M)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    q.execute_block(x, K, out=out)
    torch.cuda.synchronize()
    assert q.last_path() == 2 and intact(buf, K * M) and bool(torch.isfinite(torch.view_as_real(out)).all())
    # large-M cooperative kernel
    M, m, K = 1024, 4, 2048 + 3
    x = torch.from_numpy(_rand_c(rng, K * M // 2)).cuda()
    buf, out = guarded(K * M)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    q.execute_block(x, K, out=out)
    torch.cuda.synchronize()
    assert q.last_path() == 3 and intact(buf, K * M) and bool(torch.isfinite(torch.view_as_real(out)).all())
    # the same geometry the other way (single-SM synthesis kernel, state written by the kernel), then m = 6, which
    # stays on the cooperative group kernels
    K = 2048 + 64
    X = torch.from_numpy(_rand_c(rng, K * M)).cuda()
    buf, out = guarded(K * M // 2)
    q = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    q.execute_block(X, K, out=out)
    torch.cuda.synchronize()
    assert q.last_path() == 3 and intact(buf, K * M // 2) and bool(torch.isfinite(torch.view_as_real(out)).all())
    for type_, n_in, n_out in ((A, K * M // 2, K * M), (S, K * M, K * M // 2)):
        buf, out = guarded(n_out)
        q = yb.FirPfbCh2.new_kaiser(type_, M, 6, 60.0)
        q.execute_block(X[:n_in], K, out=out)
        torch.cuda.synchronize()
        assert q.last_path() == 3 and intact(buf, n_out) and bool(torch.isfinite(torch.view_as_real(out)).all())
    # tensor-core firfilt, 100 streams (12.5 tiles of 8) x 4096
    S_, N = 100, 4096
    x = torch.from_numpy(_rand_c(rng, S_ * N)).cuda()
    buf, out = guarded(S_ * N)
    f = yb.FirFilt.new(yb.fir_design_kaiser(63, 0.25, 60.0, 0.0), n_streams=S_)
    f.execute_block(x, out=out)
    torch.cuda.synchronize()
    assert f.last_path() == 4 and intact(buf, S_ * N) and bool(torch.isfinite(torch.view_as_real(out)).all())
    # channel-major transpose with ragged tile edges
    Kf, Mc = 1000 + 7, 48
    y = torch.from_numpy(_rand_c(rng, Kf * Mc)).cuda().view(Kf, Mc)
    buf, out = guarded(Kf * Mc)
    cm = yb.channel_major(y, out=out.view(Mc, Kf))
    torch.cuda.synchronize()
    assert intact(buf, Kf * Mc)
    assert bool(torch.equal(torch.view_as_real(cm), torch.view_as_real(y.t().contiguous())))


# ------------------------------------------------------------------ the reference's own firfilt autotests, on the GPU object
def _psd_regions_ok(H, regions):
    """utility/test_helpers.rs validate_psd_spectrum: every bin of a region inside [pmin, pmax] dB on the tested sides."""
    n = H.size
    f = np.arange(n) / n - 0.5
    psd = 20.0 * np.log10(np.maximum(np.abs(np.fft.fftshift(H)), 1e-12))
    for fmin, fmax, pmin, pmax, test_lo, test_hi in regions:
        sel = (f >= fmin) & (f <= fmax)
        if test_lo and not (psd[sel] >= pmin).all():
            return False
        if test_hi and not (psd[sel] <= pmax).all():
            return False
    return True


def test_firfilt_crcf_kaiser_autotest_on_cuda(monkeypatch):
    """Reference test_firfilt_crcf_kaiser (src/filter/fir/firfilt.rs:354-370): new_kaiser(51, 0.2, 60, 0), scale 0.4;
    stop bands <= -60 dB, pass band within +-0.1 dB.  The reference reads the response off the taps; here it is MEASURED
    by pushing an impulse through the GPU filter -- through the tensor-core kernel and through the FFMA2 kernel."""
    import torch
    regions = [(-0.5, -0.25, 0.0, -60.0, False, True), (-0.15, 0.15, -0.1, 0.1, True, True), (0.25, 0.5, 0.0, -60.0, False, True)]
    S_, N = 32, 4096
    for tc, want in (("1", 4), ("0", 2)):
        monkeypatch.setenv("YG_FIRFILT_TC", tc)
        q = yb.FirFilt.new_kaiser(51, 0.2, 60.0, 0.0, n_streams=S_)
        q.set_scale(0.4)
        x = np.zeros((S_, N), dtype=np.complex64)
        for s in range(S_):
            x[s, 7 * s] = 1.0                                # a different delay per stream
        y = q.execute_block(torch.from_numpy(x).cuda()).view(S_, N).cpu().numpy()
        assert q.last_path() == want
        for s in (0, 5, 31):
            h_meas = y[s, 7 * s: 7 * s + 1200]               # impulse response (51 taps, zero after)
            assert np.abs(y[s, : 7 * s]).max(initial=0.0) == 0.0
            assert _psd_regions_ok(np.fft.fft(h_meas, 1200), regions), (tc, s)


def test_firfilt_crcf_copy_autotest_on_cuda(monkeypatch):
    """Reference test_firfilt_crcf_copy (firfilt.rs:541-583): new_kaiser(21, 0.345, 60, 0), scale 2; run, clone, keep running
    both: same coefficients, scale, length and outputs.  Sample-sized calls (generic kernel) and block-sized calls
    (tensor-core kernel)."""
    import torch
    rng = np.random.default_rng(21)
    for S_, n in ((1, 32), (32, 4096)):
        q = yb.FirFilt.new_kaiser(21, 0.345, 60.0, 0.0, n_streams=S_)
        q.set_scale(2.0)
        x1 = torch.from_numpy(_rand_c(rng, S_ * n)).cuda()
        x2 = torch.from_numpy(_rand_c(rng, S_ * n)).cuda()
        q.execute_block(x1)
        c = q.clone()
        assert c.get_scale() == q.get_scale() == 2.0 and c.len() == q.len() == 21
        ya, yb_ = q.execute_block(x2), c.execute_block(x2)
        torch.cuda.synchronize()
        assert bool(torch.equal(torch.view_as_real(ya), torch.view_as_real(yb_)))
        ref = po.firfilt_crcf(yb.fir_design_kaiser(21, 0.345, 60.0, 0.0), np.concatenate([x1.view(S_, n)[0].cpu().numpy(), x2.view(S_, n)[0].cpu().numpy()]), scale=2.0)[n:]
        assert_parity(ya.view(S_, n)[0].cpu().numpy(), ref, "continuation after clone, %d streams" % S_)


def test_back_to_back_calls_under_programmatic_dependent_launch():
    """The fused M=256 kernel is launched with programmatic stream serialization: consecutive calls may overlap one
    launch's prologue with the previous one's tail.  Nothing may be read or written early: (a) 40 unsynchronised calls of
    random even sizes writing consecutive slices, (b) the same calls all writing the SAME buffer with a torch copy kernel
    (reader of that buffer) between them -- both must reproduce the oracle."""
    import torch
    M, m = 256, 7
    rng = np.random.default_rng(77)
    sizes = [int(2 * rng.integers(32, 700)) for _ in range(40)]
    K = sum(sizes)
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    ref = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(x).reshape(K, M)
    xd = torch.from_numpy(x).cuda()
    # (a) consecutive slices
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    y = torch.empty(K * M, dtype=torch.complex64, device="cuda")
    l0 = yb.launch_count()
    f = 0
    for n in sizes:
        q.execute_block(xd[f * M // 2:(f + n) * M // 2], n, out=y[f * M:(f + n) * M])
        f += n
    torch.cuda.synchronize()
    assert yb.launch_count() - l0 == len(sizes)                  # one launch per call
    assert_parity(y.cpu().numpy().reshape(K, M), ref, "consecutive slices")
    # (b) one output buffer, copied out by a torch kernel between the calls
    q.reset()
    buf = torch.empty(max(sizes) * M, dtype=torch.complex64, device="cuda")
    keep = torch.empty(K * M, dtype=torch.complex64, device="cuda")
    f = 0
    for n in sizes:
        q.execute_block(xd[f * M // 2:(f + n) * M // 2], n, out=buf[: n * M])
        keep[f * M:(f + n) * M].copy_(buf[: n * M])
        f += n
    torch.cuda.synchronize()
    assert_parity(keep.cpu().numpy().reshape(K, M), ref, "shared output buffer")


def test_calls_hopping_between_streams_keep_the_state_in_order():
    """A handle's state lives on the device and is handed from call to call; the caller may issue consecutive calls on
    different (non-blocking) streams, through host pointers, or touch the state (get_state / set_state / reset / clone)
    in between without synchronising.  The library orders them (StreamOrder, common.cuh): the concatenated output must be
    the oracle's single-run output for every kernel family."""
    import torch
    rng = np.random.default_rng(5)
    for M, m, K in ((256, 7, 2048), (1024, 4, 1024), (32, 3, 1024), (96, 3, 512)):
        x = stimulus.noise_plus_tones(M, K * M // 2, M)
        ref = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(x).reshape(K, M)
        xd = torch.from_numpy(x).cuda()
        y = torch.empty(K * M, dtype=torch.complex64, device="cuda")
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.default_stream()]
        q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
        cuts = sorted(set(int(c) for c in rng.integers(1, K, 11))) + [K]   # odd cuts too
        f = 0
        host_parts = {}
        for i, e in enumerate(cuts):
            n = e - f
            kind = i % 4
            if kind == 3:                                   # host-pointer call between device calls
                host_parts[f] = q.execute_block(x[f * M // 2:e * M // 2], n)
            else:
                with torch.cuda.stream(streams[kind]):
                    q.execute_block(xd[f * M // 2:e * M // 2], n, out=y[f * M:e * M])
            if i == 4:                                      # state read + written back, then a clone takes over
                st = q.get_state()
                q.set_state(*st)
                q = q.clone()
            f = e
        torch.cuda.synchronize()
        out = y.cpu().numpy().reshape(K, M)
        for f0, part in host_parts.items():
            out[f0:f0 + part.size // M] = part.reshape(-1, M)
        assert_parity(out, ref, "stream hopping M=%d" % M)
        # reset between streams: second half restarts from zero state
        q.reset()
        with torch.cuda.stream(streams[0]):
            q.execute_block(xd[: 64 * M // 2], 64, out=y[: 64 * M])
        q.reset()
        with torch.cuda.stream(streams[1]):
            q.execute_block(xd[: K * M // 2], K, out=y)
        torch.cuda.synchronize()
        assert_parity(y.cpu().numpy().reshape(K, M), ref, "reset between streams M=%d" % M)


def test_firfilt_calls_hopping_between_streams(monkeypatch):
    import torch
    S_, N, taps = 32, 1 << 14, 65
    rng = np.random.default_rng(9)
    h = rng.standard_normal(taps).astype(np.float32)
    x = (rng.standard_normal((S_, N)) + 1j * rng.standard_normal((S_, N))).astype(np.complex64)
    ref = np.stack([po.firfilt_crcf(h, x[s]) for s in range(S_)])
    xd = torch.from_numpy(x).cuda()
    y = torch.empty((S_, N), dtype=torch.complex64, device="cuda")
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.default_stream(), torch.cuda.Stream()]
    q = yb.FirFilt.new(h, n_streams=S_)
    cuts = [4096, 4096 + 2048, 4096 + 2048 + 66, 12288, N]
    f = 0
    paths = []
    for i, e in enumerate(cuts):
        with torch.cuda.stream(streams[i % 3]):
            xs = xd[:, f:e].contiguous()
            ys = q.execute_block(xs)
            y[:, f:e] = ys.view(S_, e - f)
        paths.append(q.last_path())
        f = e
    torch.cuda.synchronize()
    assert 4 in paths                                        # the tensor-core kernel took the large slices
    assert_parity(y.cpu().numpy(), ref, "firfilt stream hopping")


# ------------------------------------------------------------------ reference golden vectors straight on the GPU (no oracle)
def test_reference_dotprod_vectors_on_the_gpu(monkeypatch):
    """dotprod_crcf rand01 / rand02 / rand01-reversed (src/dotprod/mod.rs:455-524, tol 1e-3) as one output of the GPU FIR:
    with taps g = reverse(h), out[len-1] = sum_i h[i] x[i].  Once through the small-call kernels and once embedded in a
    32-stream x 4096-sample call that the tensor-core kernel takes."""
    import golden_vectors as gv
    import torch
    G = gv.load()
    cases = [(G["DOTPROD_CRCF_RAND01_H"], G["DOTPROD_CRCF_RAND01_X"], G["DOTPROD_CRCF_RAND01_Y"][0]),
             (G["DOTPROD_CRCF_RAND02_H"], G["DOTPROD_CRCF_RAND02_X"], G["DOTPROD_CRCF_RAND02_Y"][0]),
             (G["DOTPROD_CRCF_RAND01_H"][::-1], G["DOTPROD_CRCF_RAND01_X"], G["DOTPROD_CRCF_RAND01_YREV"][0])]
    monkeypatch.setenv("YG_FIRFILT_TC", "1")
    for h, x, want in cases:
        n = h.size
        g = np.ascontiguousarray(h[::-1], dtype=np.float32)
        y = yb.FirFilt.new(g).execute_block(x.astype(np.complex64))
        assert abs(y[n - 1] - want) < 1e-3
        S_, N = 32, 4096
        X = np.zeros((S_, N), dtype=np.complex64)
        for s in range(S_):
            X[s, 100 * s + 5: 100 * s + 5 + n] = x
        q = yb.FirFilt.new(g, n_streams=S_)
        Y = q.execute_block(torch.from_numpy(X).cuda()).view(S_, N).cpu().numpy()
        assert q.last_path() == 4
        for s in range(S_):
            assert abs(Y[s, 100 * s + 5 + n - 1] - want) < 1e-3, s


def test_reference_firdecim_vectors_on_the_gpu():
    """firdecim_crcf golden vectors (src/filter/fir/firdecim_test_data.rs, tol 1e-3): y[k] = (x * h)[k M], i.e. every M-th
    output of the GPU FIR (the decimator keeps the output after the FIRST sample of each block, firdecim.rs:179-191)."""
    import golden_vectors as gv
    G = gv.load()
    for M, case in gv.FIRDECIM_CASES:
        h, x, yr = (G["FIRDECIM_CRCF_DATA_%s_%s" % (case, k)] for k in "HXY")
        y = yb.FirFilt.new(h).execute_block(x.astype(np.complex64))
        np.testing.assert_allclose(y[::M][: yr.size], yr, atol=1e-3, rtol=1e-3)


# ------------------------------------------------------------------ single-SM M = 1024 kernels (config #4)
@pytest.mark.parametrize("M,m", [(1024, 1), (1024, 2), (1024, 3), (1024, 4), (512, 1), (512, 2), (512, 3), (512, 4), (512, 5),
                                 (512, 6), (512, 7)])
def test_single_sm_synthesis(M, m):
    """The one-CTA-per-SM synthesis kernels (M = 1024, m <= 4 and M = 512, m <= 7; running output sums in registers,
    frames transformed in place in a 6-stage shared ring): random prototype, three calls -- a short one (one batch per
    CTA, warm-up out of the zero prefix), a long one starting on an ODD frame (lead frame on the generic kernel, warm-up
    batches that straddle prefix | x, every CTA wrapping its ring twice) and the rest -- against the CPU path, per frame."""
    K = (8320 if M == 1024 else 16640) + 75
    rng = np.random.default_rng(5100 + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    X = _rand_c(rng, K * M)
    ref = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(X).reshape(K, M // 2)
    q = yb.FirPfbCh2.new(S, M, m, h)
    cuts = [0, 97, 97 + (6400 if M == 1024 else 12800) + 1, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(X[a * M: b * M]))
        assert q.last_path() == 3, (a, b)
    y = np.concatenate(outs).reshape(K, M // 2)
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "single-SM synthesis M=%d m=%d" % (M, m))
    # whole batches only: ONE launch per call, the kernel writes the next state (the last 32 input frames) itself
    q2 = yb.FirPfbCh2.new(S, M, m, h)
    outs = []
    for a, b in ((0, 640), (640, 704), (704, 1344)):
        n0 = yb.launch_count()
        outs.append(q2.execute_block(X[a * M: b * M]))
        assert q2.last_path() == 3 and yb.launch_count() - n0 == 1, (a, b, yb.launch_count() - n0)
    y2 = np.concatenate(outs).reshape(1344, M // 2)
    per_frame = np.abs(y2 - ref[:1344]).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())


@pytest.mark.parametrize("M,m", [(1024, 1), (1024, 3), (512, 1), (512, 2), (512, 3), (512, 4), (512, 5), (512, 6), (512, 7)])
def test_single_sm_analysis_state_in_kernel(M, m):
    """The one-CTA-per-SM analysis kernels (M = 1024, m <= 4; M = 512, m <= 7) also write the object's next state (one
    launch per call): uneven calls, odd-parity starts and a call shorter than the history, against the CPU path."""
    K = 2600 if M == 1024 else 5200
    rng = np.random.default_rng(5200 + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = po.FirPfbCh2.new(po.ANALYZER, M, m, h).execute_block(x).reshape(K, M)
    q = yb.FirPfbCh2.new(A, M, m, h)
    cuts = [0, 64, 64 + 2, 64 + 2 + 1301, 64 + 2 + 1301 + 66, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        n0 = yb.launch_count()
        outs.append(q.execute_block(x[a * M // 2: b * M // 2]))
        if b - a >= 64 and (b - a) % (4 if M == 1024 else 8) == 0 and a % 2 == 0:
            assert q.last_path() == 3 and yb.launch_count() - n0 == 1, (a, b, yb.launch_count() - n0)
    y = np.concatenate(outs).reshape(K, M)
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "single-SM analysis M=%d m=%d" % (M, m))


# ------------------------------------------------------------------ generic kernels: mixed-radix transform for any even M
@pytest.mark.parametrize("M,m", [(10, 3), (18, 2), (24, 4), (48, 5), (50, 2), (60, 3), (66, 2), (74, 1), (96, 4), (100, 5),
                                 (120, 3), (126, 2), (202, 2), (240, 4), (250, 3), (384, 9), (768, 2), (1000, 4),
                                 (1536, 3), (3000, 1), (256, 10), (64, 12)])
def test_generic_kernels_any_even_M(M, m):
    """Geometries without a fused kernel (M not a power of two, or m beyond the fused range) run on the generic
    kernels, whose transform is a mixed-radix Stockham (one pass per prime factor, 4 before 2; a prime factor p costs
    p MACs per bin -- M = 202 = 2 x 101 exercises a large one): analysis, then synthesis of the channel frames, against
    the CPU path over several uneven calls, and the transform alone (delta prototype) against numpy's f64 FFT."""
    rng = np.random.default_rng(6000 + M + m)
    K = max(24, min(600, (1 << 17) // M))
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = po.FirPfbCh2.new(po.ANALYZER, M, m, h).execute_block(x).reshape(K, M)
    q = yb.FirPfbCh2.new(A, M, m, h)
    cuts = [0, 1, 8, K // 2 + 1, K]
    y = np.concatenate([q.execute_block(x[a * M // 2: b * M // 2]) for a, b in zip(cuts, cuts[1:])]).reshape(K, M)
    assert q.last_path() == 1
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "generic analysis M=%d m=%d" % (M, m))
    refs = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(ref.reshape(-1)).reshape(K, M // 2)
    qs = yb.FirPfbCh2.new(S, M, m, h)
    ys = np.concatenate([qs.execute_block(ref.reshape(-1)[a * M: b * M]) for a, b in zip(cuts, cuts[1:])]).reshape(K, M // 2)
    assert qs.last_path() == 1
    sc = max(1.0, np.abs(refs).max())
    per_frame = np.abs(ys - refs).max(axis=1) / sc
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(ys / sc, refs / sc, "generic synthesis M=%d m=%d" % (M, m))
    # oracle-free: delta prototype h[n] = [n < M] makes frame k of the analyser M^-1 * IDFT_unnorm of the (rotated) last M
    # input samples, i.e. numpy.fft.ifft of them
    hd = np.zeros(2 * M * m, dtype=np.float32)
    hd[:M] = 1.0
    qd = yb.FirPfbCh2.new(A, M, m, hd)
    Kd = 16
    xd = _rand_c(rng, Kd * M // 2)
    yd = qd.execute_block(xd).reshape(Kd, M)
    xs = np.concatenate([np.zeros(M, dtype=np.complex128), xd.astype(np.complex128)])
    for k in range(2, Kd):
        tk = (k + 1) * (M // 2) - 1 + M                    # index of the newest sample in xs
        w = xs[tk - M + 1: tk + 1][::-1]                   # w[b] = x[tk - b]
        X = np.roll(w, (M // 2) if (k & 1) else 0)         # branch b lands at (b + parity * M/2) mod M
        expect = np.fft.ifft(X)
        err = np.abs(yd[k] - expect).max() / max(1e-30, np.abs(expect).max())
        assert err <= 2e-6 * max(2.0, np.log2(M)), (k, err)


@pytest.mark.parametrize("M,p,S_", [(1, 3, 2), (2, 4, 3), (6, 5, 2), (10, 14, 3), (12, 3, 1), (48, 8, 3), (100, 6, 2), (128, 14, 3),
                                    (202, 4, 1), (240, 10, 2), (256, 14, 2), (1000, 8, 1), (1024, 6, 2), (64, 20, 5)])
def test_firpfbch_generic_kernels_any_M(M, p, S_):
    """firpfbch geometries without a fused kernel (M other than 8 / 16 / 32 / 64, or p > 16): analysis and synthesis on the
    tiled generic kernels against the CPU path, several streams, three uneven calls; then the analyser's transform alone
    (prototype = one tap per branch) against numpy's f64 FFT."""
    rng = np.random.default_rng(7000 + M + p)
    h = rng.standard_normal(M * p).astype(np.float32)
    Q = max(3 * p + 5, min(400, (1 << 15) // M))
    x = _rand_c(rng, S_ * Q * M).reshape(S_, Q * M)
    cuts = [0, 1, Q // 2, Q]
    for type_, otype in ((A, po.ANALYZER), (S, po.SYNTHESIZER)):
        q = yb.FirPfbCh.new(type_, M, p, h, n_streams=S_)
        y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, a * M: b * M])).reshape(S_, -1) for a, b in zip(cuts, cuts[1:])],
                           axis=1)
        assert q.last_path() == 1
        ref = np.stack([po.FirPfbCh.new(otype, M, p, h).execute_block(x[s]) for s in range(S_)])
        scale = max(1.0, np.abs(ref).max())
        assert_parity(y / scale, ref / scale, "generic firpfbch type=%d M=%d p=%d" % (int(type_), M, p))
        per_frame = np.abs(y - ref).reshape(S_, Q, M).max(axis=2) / scale
        assert per_frame.max() <= 1e-4, np.unravel_index(per_frame.argmax(), per_frame.shape)
    # oracle-free: h[b + n M] = [n == 0] makes frame q the forward DFT of X[M-1-b] = x[q M + M-1 - b], i.e. of the frame
    # itself (X[c] = x[q M + c])
    hd = np.zeros(M * p, dtype=np.float32)
    hd[:M] = 1.0
    qd = yb.FirPfbCh.new(A, M, p, hd, n_streams=1)
    xd = _rand_c(rng, 8 * M)
    yd = qd.execute_block(xd).reshape(8, M)
    expect = np.fft.fft(xd.astype(np.complex128).reshape(8, M), axis=1)
    err = np.abs(yd - expect).max() / max(1e-30, np.abs(expect).max())
    assert err <= 2e-6 * max(2.0, np.log2(M)), err
