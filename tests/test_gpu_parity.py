"""GPU parity tests proper: the CUDA path, called through the C ABI (via the thin Python mirror),
against the CPU oracle on identical seeded inputs.  Tolerance (BASELINE.json north_star):
rel-RMS <= 1e-5 and max-abs <= 1e-4 per output sample, f32 -- see tests/parity.py.
"""
import numpy as np
import pytest

import closed_forms as cf
import golden_vectors as gv
import stimulus
from oracle import pyoracle as po
from parity import assert_parity, errors

pytestmark = pytest.mark.gpu

import yagi_b200 as yb  # noqa: E402

A, S = yb.ANALYZER, yb.SYNTHESIZER


def _rand_c(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def _oracle_analysis(M, m, x, h=None, as_=60.0):
    q = po.FirPfbCh2.new(po.ANALYZER, M, m, h) if h is not None else po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, as_)
    return q.execute_block(x)


# ------------------------------------------------------------------ firpfbch2 analysis
def test_config1_M16_m5_65536_samples():
    """BASELINE config #1: firpfbch2 analysis M=16 m=5 on 65536 cf32 samples vs the CPU path."""
    M, m, N = 16, 5, 65536
    x = stimulus.noise_plus_tones(0, N, M)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    np.testing.assert_array_equal(q.get_taps(), po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0).taps())
    y = q.execute_block(x)
    assert y.size == 2 * N
    assert_parity(y, _oracle_analysis(M, m, x), "config #1")


def test_config3_geometry_M256_m7_prefix():
    """BASELINE config #3 geometry (M=256, m=7) on a 2^20-sample prefix vs the oracle."""
    M, m, N = 256, 7, 1 << 20
    x = stimulus.noise_plus_tones(0, N, M)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    y = q.execute_block(x)
    rel, mx = assert_parity(y, _oracle_analysis(M, m, x), "config #3 prefix")
    # attribute: both vs the f64 closed form on the first 64 frames
    ref64 = cf.firpfbch2_analysis(q.get_taps(), M, m, x[: 64 * M // 2]).reshape(-1)
    r_gpu, _ = errors(y[: 64 * M], ref64)
    assert r_gpu < 5e-6


@pytest.mark.parametrize("M,m", [(2, 1), (4, 3), (6, 2), (8, 5), (12, 4), (32, 5), (64, 3), (100, 2), (128, 4),
                                 (512, 3), (1024, 4), (2048, 2)])
def test_analysis_geometries_random_prototype(M, m):
    """Edge geometries incl. non-power-of-two M, arbitrary (non-Kaiser) prototypes, odd frame counts."""
    rng = np.random.default_rng(M * 31 + m)
    h = rng.standard_normal(2 * M * m + 1).astype(np.float32)        # one spare tap: must be ignored
    K = 8 * m + 3
    x = _rand_c(rng, K * M // 2)
    q = yb.FirPfbCh2.new(A, M, m, h)
    y = q.execute_block(x)
    ref = _oracle_analysis(M, m, x, h=h)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "M=%d m=%d" % (M, m))


@pytest.mark.parametrize("M,m", [(16, 5), (256, 7), (6, 2)])
def test_state_continuity_uneven_blocks_reset_clone(M, m):
    """Archetypes 4/5: block-vs-sample, uneven block sizes with odd counts
    (parity flag), reset, and clone-mid-stream producing bit-identical continuations."""
    K = 300 if M <= 16 else 96
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    ref = _oracle_analysis(M, m, x)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    cuts = [0, 1, 2, 5, 6, 37, 38, 71, K]
    ys = [q.execute_block(x[a * M // 2: b * M // 2]) for a, b in zip(cuts, cuts[1:])]
    y = np.concatenate(ys)
    assert_parity(y, ref, "uneven blocks")
    one = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0).execute_block(x)
    assert_parity(one, ref, "one block")
    # frame-at-a-time `execute`
    q.reset()
    y1 = np.concatenate([q.execute(x[k * M // 2:(k + 1) * M // 2]) for k in range(9)])
    assert_parity(y1, ref[: 9 * M], "execute()")
    # reset after an odd number of frames clears the flag
    q.reset()
    assert_parity(q.execute_block(x), ref, "after reset")
    # clone mid-stream at an odd frame count
    q.reset()
    q.execute_block(x[: 7 * M // 2])
    c = q.clone()
    a = q.execute_block(x[7 * M // 2: 40 * M // 2])
    b = c.execute_block(x[7 * M // 2: 40 * M // 2])
    np.testing.assert_array_equal(a, b)
    assert_parity(a, ref[7 * M: 40 * M], "clone continuation")


def test_state_get_set_and_time_shards():
    """Shard hand-off: shards primed with set_state(halo) reproduce the single pass,
    including a cut +-2 frames around each boundary."""
    M, m, K = 256, 7, 512
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    whole = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0).execute_block(x).reshape(K, M)
    ref = _oracle_analysis(M, m, x).reshape(K, M)
    assert_parity(whole, ref, "whole")
    for ws in (2, 3, 8):
        shards = yb.firpfbch2_time_shards(K, M, m, ws)
        parts = []
        for sh in shards:
            q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
            assert q.state_len() == sh.halo_len
            q.set_state(stimulus.noise_plus_tones(sh.halo_begin, sh.halo_len, M), 0)
            parts.append(q.execute_block(x[sh.sample_begin: sh.sample_end]))
        y = np.concatenate(parts).reshape(K, M)
        assert_parity(y, ref, "ws=%d" % ws)
        for sh in shards[1:]:
            b = sh.frame_begin
            np.testing.assert_allclose(y[b - 2: b + 2], whole[b - 2: b + 2], atol=2e-6)
    # get_state returns exactly the last (4m-1) M/2 inputs and the parity
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    q.execute_block(x[: 33 * M // 2])
    hist, flag = q.get_state()
    assert flag == 1
    np.testing.assert_array_equal(hist, x[33 * M // 2 - q.state_len(): 33 * M // 2])


@pytest.mark.parametrize("n", [2, 4, 6, 8, 10, 16, 20, 22, 24, 26, 30, 32, 36, 48, 64, 92, 96, 120, 130, 192])
def test_fft_stage_against_reference_golden_vectors(n):
    """Pins the CUDA IFFT stage on the reference's own Fft golden pairs (src/fft/test_data.rs),
    oracle-free: with the prototype h[0..M) = 1 (else 0) the analyser is a pure backward DFT:
    frame 1 gives y = (1/M) IDFT_unnorm(roll(V, M/2)), V[b] = s[M-1-b].  Feed s so that the rolled
    V equals the golden Y; the output must be the golden X (tolerance 2e-4, src/fft/mod.rs:125-151)."""
    G = gv.load()
    X, Y = G["FFT_TEST_X%d" % n], G["FFT_TEST_Y%d" % n]
    M = n
    h = np.zeros(2 * M, dtype=np.float32)
    h[:M] = 1.0
    V = np.roll(Y, -M // 2)                 # roll(V, +M/2) == Y
    s = V[::-1].astype(np.complex64)        # V[b] = s[M-1-b]
    q = yb.FirPfbCh2.new(A, M, 1, h)
    y = q.execute_block(s).reshape(2, M)[1]
    assert np.abs(y - X).max() < 2e-4


def test_device_pointer_api_matches_host_api():
    import torch
    M, m, K = 256, 7, 256
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    yh = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0).execute_block(x)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    xd = torch.from_numpy(x).cuda()
    y1 = q.execute_block(xd[: 100 * M // 2])
    y2 = q.execute_block(xd[100 * M // 2:])
    torch.cuda.synchronize()
    yd = torch.cat([y1, y2]).cpu().numpy()
    np.testing.assert_array_equal(yd, yh)
    assert q.last_path() in (1, 2)
    assert q.last_kernel_ms() > 0.0


# ------------------------------------------------------------------ firpfbch2 synthesis
@pytest.mark.parametrize("M,m", [(2, 1), (6, 2), (16, 5), (64, 3), (256, 7), (1024, 4)])
def test_synthesis_vs_oracle(M, m):
    rng = np.random.default_rng(M + m)
    K = 12 * m + 5
    X = _rand_c(rng, K * M)
    q = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    cuts = [0, 1, 4, 5, K]
    y = np.concatenate([q.execute_block(X[a * M: b * M]) for a, b in zip(cuts, cuts[1:])])
    ref = po.FirPfbCh2.new_kaiser(po.SYNTHESIZER, M, m, 60.0).execute_block(X)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "synthesis M=%d" % M)
    c = q.clone()
    v = _rand_c(rng, 6 * M)
    np.testing.assert_array_equal(q.execute_block(v), c.execute_block(v))


@pytest.mark.parametrize("M", [8, 16, 32, 64])
def test_firpfbch2_crcf_reconstruction(M):
    """autotest firpfbch2_crcf_n8..n64 on the CUDA path: analysis -> synthesis, delay 2Mm - M/2 + 1, tol 1e-3."""
    m = 5
    nb = 8 * m * 2
    n = nb * M // 2
    s = 1
    x = np.empty(n, dtype=np.complex64)
    for i in range(n):
        s = (s * 524287) % 1031
        x[i] = np.exp(2j * np.pi * s / 1031.0)
    qa = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    qs = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    y = qs.execute_block(qa.execute_block(x))
    D = 2 * M * m - M // 2 + 1
    assert np.abs(y[:D]).max() < 1e-3
    assert np.abs(y[D:] - x[: n - D]).max() < 1e-3


def test_config4_roundtrip_M1024_m4():
    """BASELINE config #4 (2^22 of its 2^24 samples): analysis -> synthesis round trip, M=1024 m=4,
    channel matrix and reconstruction vs the CPU path (two-stage large-M kernels)."""
    M, m, N = 1024, 4, 1 << 22
    x = stimulus.noise_plus_tones(0, N, M)
    qa = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    qs = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    Y = qa.execute_block(x)
    y = qs.execute_block(Y)
    oa = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0)
    os_ = po.FirPfbCh2.new_kaiser(po.SYNTHESIZER, M, m, 60.0)
    Yr = oa.execute_block(x)
    yr = os_.execute_block(Yr)
    assert qa.last_path() == 3 and qs.last_path() == 3
    assert_parity(Y, Yr, "channel matrix")
    assert_parity(y, yr, "reconstruction")
    D = 2 * M * m - M // 2 + 1
    assert np.abs(y[D:] - x[: N - D]).max() < 2e-3


# ------------------------------------------------------------------ firpfbch
@pytest.mark.parametrize("M,p,S_", [(1, 3, 1), (4, 5, 3), (5, 4, 2), (64, 14, 8), (16, 6, 5)])
def test_firpfbch_streams_vs_oracle(M, p, S_):
    rng = np.random.default_rng(M * 7 + p)
    h = rng.standard_normal(M * p).astype(np.float32)
    Q = 3 * p + 2
    x = _rand_c(rng, S_ * Q * M).reshape(S_, Q * M)
    for type_, otype in ((A, po.ANALYZER), (S, po.SYNTHESIZER)):
        q = yb.FirPfbCh.new(type_, M, p, h, n_streams=S_)
        cut = (Q // 2) * M
        y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, :cut])).reshape(S_, -1),
                            q.execute_block(np.ascontiguousarray(x[:, cut:])).reshape(S_, -1)], axis=1)
        ref = np.stack([po.FirPfbCh.new(otype, M, p, h).execute_block(x[s]) for s in range(S_)])
        scale = max(1.0, np.abs(ref).max())
        assert_parity(y / scale, ref / scale, "firpfbch type=%d" % int(type_))


def test_config5_geometry_firpfbch_M64_m7():
    """BASELINE config #5 geometry (M=64, m=7 Kaiser), 16 streams x 2^12 samples."""
    M, m, S_, N = 64, 7, 16, 1 << 12
    x = np.stack([stimulus.noise_plus_tones(0, N, M, stream=s) for s in range(S_)])
    q = yb.FirPfbCh.new_kaiser(A, M, m, 60.0, n_streams=S_)
    y = q.execute_block(x).reshape(S_, -1)
    ref = np.stack([po.FirPfbCh.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(x[s]) for s in range(S_)])
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "config #5 geometry")


# ------------------------------------------------------------------ firfilt
@pytest.mark.parametrize("case", gv.FIRFILT_CASES)
def test_firfilt_reference_golden_vectors(case):
    G = gv.load()
    h, x, yr = (G["FIRFILT_CRCF_DATA_%s_%s" % (case, k)] for k in "HXY")
    y = yb.FirFilt.new(h).execute_block(x.astype(np.complex64))
    np.testing.assert_allclose(y, yr, atol=1e-3, rtol=1e-3)          # firfilt.rs:851-919 tolerance


def test_config2_geometry_firfilt_63_taps():
    """BASELINE config #2 geometry: 63-tap Kaiser firfilt over many streams, with state continuity."""
    S_, N = 64, 1 << 13
    h = yb.fir_design_kaiser(63, 0.25, 60.0, 0.0)
    np.testing.assert_array_equal(h, po.fir_design_kaiser(63, 0.25, 60.0, 0.0))
    x = np.stack([stimulus.noise_plus_tones(0, N, 256, stream=s) for s in range(S_)])
    q = yb.FirFilt.new(h, n_streams=S_)
    q.set_scale(0.5)
    cuts = [0, 1, 63, 64, 1000, N]
    y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, a:b])).reshape(S_, -1) for a, b in zip(cuts, cuts[1:])], axis=1)
    ref = np.stack([po.firfilt_crcf(h, x[s], scale=0.5) for s in range(S_)])
    assert_parity(y, ref, "config #2 geometry")
    assert q.get_scale() == 0.5
    q.reset()
    y2 = q.execute_block(x).reshape(S_, -1)
    assert_parity(y2, ref, "after reset")


# ------------------------------------------------------------------ fused fast path (M=256, m=7)
def test_fused_path_state_continuity_and_edges():
    """The fused kernel (last_path == 2) across calls of uneven sizes: odd-parity starts (a leading
    frame goes to the generic kernel), odd lengths (trailing frame), partial 16-pair batches,
    single-batch and multi-batch slabs, history taken from the previous call's tail."""
    M, m = 256, 7
    K = 6000
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    ref = _oracle_analysis(M, m, x).reshape(K, M)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    cuts = [0, 64, 129, 130, 331, 1000, 1067, 1131, 3000, 3001 + 64, 5999, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(x[a * M // 2: b * M // 2]))
        if b - a >= 64:
            assert q.last_path() == 2, (a, b)
    y = np.concatenate(outs).reshape(K, M)
    assert_parity(y, ref, "fused, uneven calls")
    # per-frame worst case, to catch a single bad frame hidden by the global RMS
    per_frame = np.abs(y - ref).max(axis=1)
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())


def test_fused_path_matches_generic_kernel_closely():
    """Same input through the fused kernel (one big call) and the generic kernel (calls < 64 frames)."""
    M, m, K = 256, 7, 640
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    qf = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    yf = qf.execute_block(x)
    assert qf.last_path() == 2
    qg = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    yg_ = np.concatenate([qg.execute_block(x[a * M // 2:(a + 32) * M // 2]) for a in range(0, K, 32)])
    assert qg.last_path() == 1
    assert_parity(yf, yg_, "fused vs generic", rel_tol=2e-6, abs_tol=2e-6)


def test_fused_path_full_size_properties():
    """BASELINE config #3 at full size (2^28 samples) through size-independent properties:
    (a) channel-sum checksum: sum_c y_k[c] = X_k[0] = the single polyphase partial sum
        V_k[b0], b0 = (k odd ? M/2 : 0)  => one 14-tap dot product per frame, checked for every frame;
    (b) sampled frames (start, slab boundaries, end) against the oracle on the same input window."""
    import torch
    M, m = 256, 7
    N = 1 << 28
    K = N // (M // 2)
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    xr = torch.empty(N, 2, dtype=torch.float32, device="cuda")
    xr.normal_(0.0, 1.0, generator=g)
    x = torch.view_as_complex(xr)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    y = q.execute_block(x).view(K, M)
    torch.cuda.synchronize()
    assert q.last_path() == 2
    h = torch.from_numpy(q.get_taps()).cuda()
    # (a) checksum over every frame with k >= 4m (full history inside x)
    k = torch.arange(4 * m, K, device="cuda", dtype=torch.int64)
    b0 = (k & 1) * (M // 2)
    tk = (k + 1) * (M // 2) - 1
    acc = torch.zeros(k.numel(), dtype=torch.complex128, device="cuda")
    for n in range(2 * m):
        acc += h[b0 + n * M].to(torch.float64) * x[tk - b0 - n * M].to(torch.complex128)
    chk = y[4 * m:].to(torch.complex128).sum(dim=1)
    err = (chk - acc).abs().max().item()
    assert err < 2e-4, err            # sum of 256 outputs each within ~1e-6
    del acc, chk
    # (b) sampled windows against the oracle
    halo = (4 * m - 1) * M // 2
    for f0 in (0, 2 * 7084, 2 * 7085 * 16, K // 2 - 64, K - 128):
        f0 -= f0 % 2
        nf = 128
        s0 = f0 * M // 2
        lo = max(0, s0 - halo - M // 2)
        pre = x[lo:s0].cpu().numpy()
        seg = x[s0: s0 + nf * M // 2].cpu().numpy()
        o = po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0)
        if pre.size:
            pad = (-pre.size) % M                      # keep the priming an even number of frames
            o.execute_block(np.concatenate([np.zeros(pad, dtype=np.complex64), pre]))
        ref = o.execute_block(seg)
        got = y[f0: f0 + nf].reshape(-1).cpu().numpy()
        assert_parity(got, ref, "full-size window at frame %d" % f0)


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6, 8])
def test_fused_path_other_semi_lengths(m):
    """The fused kernel is instantiated for M=256, m=1..8 (2m+1 register taps)."""
    M, K = 256, 700
    rng = np.random.default_rng(m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    q = yb.FirPfbCh2.new(A, M, m, h)
    y = np.concatenate([q.execute_block(x[: 301 * M // 2]), q.execute_block(x[301 * M // 2:])])
    assert q.last_path() == 2
    ref = _oracle_analysis(M, m, x, h=h)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "fused m=%d" % m)


# ------------------------------------------------------------------ fused synthesis (M=256, m=1..7)
@pytest.mark.parametrize("m", [7, 1, 4])
def test_fused_synthesis_state_continuity_and_edges(m):
    """The fused synthesis kernel (last_path == 2) across calls of uneven sizes: odd-parity starts,
    lengths that are not multiples of 32, history from the previous call, clone mid-stream."""
    M, K = 256, 3000
    rng = np.random.default_rng(100 + m)
    X = _rand_c(rng, K * M)
    ref = po.FirPfbCh2.new_kaiser(po.SYNTHESIZER, M, m, 60.0).execute_block(X).reshape(K, M // 2)
    q = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    cuts = [0, 64, 129, 130, 331, 1000, 1067, 1131, 2001, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(X[a * M: b * M]))
        if b - a >= 96:
            assert q.last_path() == 2, (a, b)
    y = np.concatenate(outs).reshape(K, M // 2)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "fused synthesis m=%d" % m)
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    # state is plain data: (4m-1) input frames + parity
    hist, flag = q.get_state()
    assert flag == K % 2 and hist.size == (4 * m - 1) * M
    np.testing.assert_array_equal(hist, X[(K - (4 * m - 1)) * M:])
    c = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    c.set_state(hist, flag)
    v = _rand_c(rng, 200 * M)
    np.testing.assert_array_equal(q.execute_block(v), c.execute_block(v))


def test_fused_round_trip_M256():
    """analysis -> synthesis on the fused kernels reconstructs the input (delay 2Mm - M/2 + 1) and
    matches the CPU path."""
    M, m, N = 256, 7, 1 << 20
    x = stimulus.noise_plus_tones(0, N, M)
    qa = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    qs = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    Y = qa.execute_block(x)
    y = qs.execute_block(Y)
    assert qa.last_path() == 2 and qs.last_path() == 2
    yr = po.FirPfbCh2.new_kaiser(po.SYNTHESIZER, M, m, 60.0).execute_block(
        po.FirPfbCh2.new_kaiser(po.ANALYZER, M, m, 60.0).execute_block(x))
    assert_parity(y, yr, "round trip vs CPU path")
    D = 2 * M * m - M // 2 + 1
    assert np.abs(y[D:] - x[: N - D]).max() < 2e-3


@pytest.mark.parametrize("h_len,S_,N", [(63, 3, 10000), (64, 2, 4096 + 17), (1, 4, 5000), (23, 1, 9000), (65, 2, 5000),
                                        (100, 2, 9000), (128, 1, 4096 * 2 + 5), (200, 3, 10000), (256, 2, 8000), (257, 1, 5000)])
def test_firfilt_fast_kernel_edges(h_len, S_, N):
    """Register-blocked firfilt kernel (tap capacities 64 / 128 / 256): ragged tile ends, history across calls,
    1..256 taps; h_len = 257 exercises the generic kernel at the same sizes."""
    rng = np.random.default_rng(h_len * 7 + S_)
    h = rng.standard_normal(h_len).astype(np.float32)
    x = _rand_c(rng, S_ * N).reshape(S_, N)
    q = yb.FirFilt.new(h, n_streams=S_)
    q.set_scale(1.25)
    cuts = [0, 4999, N]
    y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, a:b])).reshape(S_, -1) for a, b in zip(cuts, cuts[1:])], axis=1)
    ref = np.stack([po.firfilt_crcf(h, x[s], scale=1.25) for s in range(S_)])
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "firfilt h_len=%d" % h_len)


@pytest.mark.parametrize("type_,otype", [(A, po.ANALYZER), (S, po.SYNTHESIZER)])
@pytest.mark.parametrize("p,S_,Q", [(14, 9, 50), (14, 4, 16), (2, 5, 33), (16, 8, 100), (6, 13, 47), (14, 600, 40)])
def test_firpfbch_fused_M64(p, S_, Q, type_, otype):
    """Fused critically sampled channelizers (M=64, analysis and synthesis): groups of four streams on the
    fused kernel, the remainder on the generic one; ragged 16-frame batches; history carried across calls;
    more stream groups than SMs (slabs that start mid-stream)."""
    M = 64
    rng = np.random.default_rng(p * 100 + S_)
    h = rng.standard_normal(M * p).astype(np.float32)
    x = _rand_c(rng, S_ * Q * M).reshape(S_, Q * M)
    q = yb.FirPfbCh.new(type_, M, p, h, n_streams=S_)
    cut = (Q // 2 + 3) * M
    y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, :cut])).reshape(S_, -1),
                        q.execute_block(np.ascontiguousarray(x[:, cut:])).reshape(S_, -1)], axis=1)
    ref = np.stack([po.FirPfbCh.new(otype, M, p, h).execute_block(x[s]) for s in range(S_)])
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "fused firpfbch type=%d p=%d S=%d" % (int(type_), p, S_))
    per_stream = np.abs(y - ref).max(axis=1) / scale
    assert per_stream.max() <= 1e-4, int(per_stream.argmax())


@pytest.mark.parametrize("M,m", [(1024, 4), (1024, 2), (1024, 7), (512, 5), (2048, 3), (4096, 2)])
def test_large_M_two_stage_path(M, m):
    """Large-M analysis (last_path == 3): whole 32-frame batches on the fused cooperative kernel, the rest on
    the two-stage (FIR kernel + in-place FFT kernel) path; uneven call sizes, odd-parity starts, history from
    the previous call."""
    K = 700 if M <= 1024 else 300
    rng = np.random.default_rng(900 + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = _oracle_analysis(M, m, x, h=h).reshape(K, M)
    q = yb.FirPfbCh2.new(A, M, m, h)
    cuts = [0, 64, 129, 130, 331, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(x[a * M // 2: b * M // 2]))
        if b - a >= 64:
            assert q.last_path() == 3, (a, b)
    y = np.concatenate(outs).reshape(K, M)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "large-M M=%d m=%d" % (M, m))
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())


@pytest.mark.parametrize("M,m", [(1024, 4), (1024, 2), (1024, 7), (512, 5), (2048, 3), (4096, 2)])
def test_large_M_two_stage_synthesis(M, m):
    """Large-M synthesis (last_path == 3): fused cooperative kernel (DFT teams -> L2 ring -> overlap-add role)."""
    K = 500 if M <= 1024 else 400
    rng = np.random.default_rng(950 + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    X = _rand_c(rng, K * M)
    ref = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(X).reshape(K, M // 2)
    q = yb.FirPfbCh2.new(S, M, m, h)
    cuts = [0, 96, 225, 226, 323, K]              # the last call starts on an odd frame: its first pair straddles prefix | x
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(X[a * M: b * M]))
        if b - a >= 97:
            assert q.last_path() == 3, (a, b)
    y = np.concatenate(outs).reshape(K, M // 2)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "large-M synthesis M=%d m=%d" % (M, m))
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())


@pytest.mark.parametrize("M,m", [(512, 7), (1024, 4), (2048, 4), (4096, 3)])
def test_large_M_fused_many_batches_per_group(M, m):
    """Fused large-M kernels with every group of M/256 CTAs walking >= 6 batches, so that the L2 ring
    (4 slots) wraps and both counters of every slot are reused: analysis, then synthesis of the channel
    matrix, each against the CPU path over the whole call."""
    groups = 148 // (M // 256)
    K = 32 * (6 * groups + 3) + 6                  # a few frames beyond whole batches: tail on the two-stage path
    rng = np.random.default_rng(4000 + M)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = _oracle_analysis(M, m, x, h=h).reshape(K, M)
    q = yb.FirPfbCh2.new(A, M, m, h)
    y = q.execute_block(x).reshape(K, M)
    assert q.last_path() == 3
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "fused large-M analysis M=%d" % M)
    refs = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(ref.reshape(-1)).reshape(K, M // 2)
    qs = yb.FirPfbCh2.new(S, M, m, h)
    ys = qs.execute_block(ref.reshape(-1)).reshape(K, M // 2)
    assert qs.last_path() == 3
    sc = max(1.0, np.abs(refs).max())
    per_frame = np.abs(ys - refs).max(axis=1) / sc
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(ys / sc, refs / sc, "fused large-M synthesis M=%d" % M)


@pytest.mark.parametrize("M,m", [(1024, 4), (512, 3)])
def test_large_M_misaligned_device_buffers_use_two_stage_kernels(M, m):
    """The fused large-M kernels stage 16-byte chunks; device buffers that start on an odd sample must take the
    two-stage kernels for the whole call (analysis and synthesis) and still match the CPU path."""
    import torch
    K = 400
    rng = np.random.default_rng(77 + M)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = _oracle_analysis(M, m, x, h=h)
    buf = torch.zeros(K * M // 2 + 1, dtype=torch.complex64, device="cuda")
    buf[1:] = torch.from_numpy(x).cuda()
    q = yb.FirPfbCh2.new(A, M, m, h)
    y = q.execute_block(buf[1:]).cpu().numpy()
    assert q.last_path() == 3
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "misaligned large-M analysis")
    refs = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(ref)
    bufs = torch.zeros(K * M + 1, dtype=torch.complex64, device="cuda")
    bufs[1:] = torch.from_numpy(ref).cuda()
    qs = yb.FirPfbCh2.new(S, M, m, h)
    ys = qs.execute_block(bufs[1:]).cpu().numpy()
    assert qs.last_path() == 3
    sc = max(1.0, np.abs(refs).max())
    assert_parity(ys / sc, refs / sc, "misaligned large-M synthesis")


def test_host_pointer_pipeline_multi_chunk():
    """Host-pointer entry point across several 32 MiB chunks (3-stream H2D / kernel / D2H pipeline):
    state must carry across chunk boundaries exactly as across calls."""
    M, m = 256, 7
    N = (1 << 23) + 5 * (M // 2)                   # 2 full chunks + a ragged tail, odd frame count
    x = stimulus.noise_plus_tones(0, N, M)
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    y = q.execute_block(x)
    ref = _oracle_analysis(M, m, x)
    assert_parity(y, ref, "host pipeline, 3 chunks")
    # pinned buffers give the same result
    hx, hy = yb.PinnedArray(N), yb.PinnedArray(2 * N)
    hx.array[:] = x
    q.reset()
    q.execute_block(hx.array, out=hy.array)
    np.testing.assert_array_equal(hy.array, y)
    hx.close(); hy.close()
    # and the synthesiser's host path (chunks of whole 32-frame rounds)
    K = 40001                                      # > 2 chunks of 16384 frames, odd count
    X = _rand_c(np.random.default_rng(5), K * M)
    qs = yb.FirPfbCh2.new_kaiser(S, M, m, 60.0)
    ys = qs.execute_block(X)
    refs = po.FirPfbCh2.new_kaiser(po.SYNTHESIZER, M, m, 60.0).execute_block(X)
    sc = max(1.0, np.abs(refs).max())
    assert_parity(ys / sc, refs / sc, "host pipeline synthesis")


def test_misaligned_device_input_falls_back_to_generic_kernels():
    """TMA needs 16-byte aligned sources; a device tensor that starts on an odd sample (8-byte aligned only)
    must still work -- through the generic kernels."""
    import torch
    M, m, K = 256, 7, 200
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    ref = _oracle_analysis(M, m, x)
    buf = torch.zeros(K * M // 2 + 1, dtype=torch.complex64, device="cuda")
    buf[1:] = torch.from_numpy(x).cuda()
    q = yb.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    y = q.execute_block(buf[1:])
    torch.cuda.synchronize()
    assert q.last_path() == 1
    assert_parity(y.cpu().numpy(), ref, "misaligned input")


def test_api_error_behaviour():
    """Argument errors surface as the reference's Error variants (src/error.rs:6-14), not as crashes."""
    import torch
    q = yb.FirPfbCh2.new_kaiser(A, 16, 3, 60.0)
    with pytest.raises(yb.ConfigError):
        q.execute(np.zeros(7, dtype=np.complex64))                  # one frame is M/2 = 8 samples
    with pytest.raises(yb.ConfigError):
        q.execute_block(np.zeros(20, dtype=np.complex64))           # not a whole number of frames
    with pytest.raises(yb.ConfigError):
        q.execute_block(np.zeros(16, dtype=np.complex64), 2, out=np.zeros(16, dtype=np.complex64))   # out too small
    with pytest.raises(yb.YagiError):
        q.execute_block(np.zeros(16, dtype=np.complex64), 2, out=np.zeros(32, dtype=np.complex128))  # wrong dtype
    with pytest.raises(yb.YagiError):
        q.execute_block(torch.zeros(16, dtype=torch.float32, device="cuda"))                          # wrong dtype on device
    with pytest.raises(yb.YagiError):
        q.set_state(np.zeros(q.state_len(), dtype=np.complex64), 2)                                  # flag must be 0/1
    with pytest.raises(yb.ConfigError):
        q.set_state(np.zeros(q.state_len() + 1, dtype=np.complex64), 0)
    # zero frames is a no-op
    assert q.execute_block(np.zeros(0, dtype=np.complex64)).size == 0
    f = yb.FirFilt.new(np.ones(5, dtype=np.float32), n_streams=3)
    with pytest.raises(yb.ConfigError):
        f.execute_block(np.zeros(10, dtype=np.complex64))           # not divisible by n_streams
    c = yb.FirPfbCh.new_kaiser(A, 8, 2, 60.0, n_streams=2)
    with pytest.raises(yb.ConfigError):
        c.execute(np.zeros(8, dtype=np.complex64))                  # needs M * n_streams samples
    assert (c.get_type(), c.get_num_channels(), c.get_p(), c.get_n_streams()) == (A, 8, 4, 2)
    assert repr(q).startswith("FirPfbCh2") and q.get_m() == 3 and q.get_num_channels() == 16


@pytest.mark.parametrize("M,m", [(64, 5), (64, 1), (64, 7), (64, 8), (128, 7), (128, 2), (128, 8)])
def test_fused_small_M_analysis(M, m):
    """firpfbch2 analysis M=64 / M=128 on the fused small-M kernel (256/M time slabs per CTA): slab
    boundaries, uneven call sizes, odd-parity starts, ragged 16-pair batches."""
    K = 9000
    rng = np.random.default_rng(640 + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = _oracle_analysis(M, m, x, h=h).reshape(K, M)
    q = yb.FirPfbCh2.new(A, M, m, h)
    cuts = [0, 256, 513, 514, 1500, 1501 + 4096, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(x[a * M // 2: b * M // 2]))
        if b - a >= 256:
            assert q.last_path() == 2, (a, b)
    y = np.concatenate(outs).reshape(K, M)
    scale = max(1.0, np.abs(ref).max())
    assert_parity(y / scale, ref / scale, "small-M M=%d m=%d" % (M, m))
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())


@pytest.mark.parametrize("M,m", [(64, 5), (64, 1), (64, 7), (128, 7), (128, 2), (128, 4)])
def test_fused_small_M_synthesis(M, m):
    """firpfbch2 synthesis M=64 / M=128 on the fused small-M kernel (256/M time slabs per CTA, warm-up batch per
    slab): slab boundaries, more slabs than batches, uneven call sizes, odd-parity starts, state hand-over."""
    K = 40000 if M == 64 else 21000                # > 148 * S slabs of several batches each
    rng = np.random.default_rng(660 + M + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    X = _rand_c(rng, K * M)
    ref = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(X).reshape(K, M // 2)
    q = yb.FirPfbCh2.new(S, M, m, h)
    cuts = [0, 256, 513, 514, 1500, 1501 + 4096, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(X[a * M: b * M]))
        if b - a >= 300:
            assert q.last_path() == 2, (a, b)
    y = np.concatenate(outs).reshape(K, M // 2)
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "small-M synthesis M=%d m=%d" % (M, m))
    hist, flag = q.get_state()
    c = yb.FirPfbCh2.new(S, M, m, h)
    c.set_state(hist, flag)
    v = _rand_c(rng, 700 * M)
    np.testing.assert_array_equal(q.execute_block(v), c.execute_block(v))


@pytest.mark.parametrize("M,m", [(16, 5), (16, 1), (16, 8), (8, 3), (8, 7), (32, 4), (32, 7), (32, 1)])
def test_fused_tiny_M_analysis(M, m):
    """firpfbch2 analysis M=8 / 16 / 32 on the fused tiny-M kernel (a FIR warp and a DFT warp per unit, 32/M time
    slabs per warp): several batches per slab, slabs of unequal length inside a warp, a ragged last batch, a call
    with fewer batches than slabs, odd-parity starts, history from the previous call."""
    slabs = 148 * 8 * (32 // M)
    K = 32 * (3 * slabs + slabs // 3) + 22
    rng = np.random.default_rng(700 + M + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    x = _rand_c(rng, K * M // 2)
    ref = _oracle_analysis(M, m, x, h=h).reshape(K, M)
    q = yb.FirPfbCh2.new(A, M, m, h)
    cuts = [0, 2500, 5001, 5002, 9000, 9001 + 40000, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(x[a * M // 2: b * M // 2]))
        if b - a >= 2100:
            assert q.last_path() == 2, (a, b)
    y = np.concatenate(outs).reshape(K, M)
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "tiny-M M=%d m=%d" % (M, m))


@pytest.mark.parametrize("M,m", [(16, 5), (16, 1), (16, 7), (8, 3), (8, 7), (32, 4), (32, 7), (32, 1)])
def test_fused_tiny_M_synthesis(M, m):
    """firpfbch2 synthesis M=8 / 16 / 32 on the fused tiny-M kernel (a DFT warp and an overlap-add warp per unit,
    32/M time slabs per warp, warm-up batch per slab): several batches per slab, slabs of unequal length inside a
    warp, a call with fewer batches than slabs, odd-parity starts (a frame pair straddling prefix | x)."""
    slabs = 148 * 8 * (32 // M)
    K = 32 * (2 * slabs + slabs // 3) + 22
    rng = np.random.default_rng(800 + M + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    X = _rand_c(rng, K * M)
    ref = po.FirPfbCh2.new(po.SYNTHESIZER, M, m, h).execute_block(X).reshape(K, M // 2)
    q = yb.FirPfbCh2.new(S, M, m, h)
    cuts = [0, 2500, 5001, 5002, 9000, 9001 + 40000, K]
    outs = []
    for a, b in zip(cuts, cuts[1:]):
        outs.append(q.execute_block(X[a * M: b * M]))
        if b - a >= 2100:
            assert q.last_path() == 2, (a, b)
    y = np.concatenate(outs).reshape(K, M // 2)
    scale = max(1.0, np.abs(ref).max())
    per_frame = np.abs(y - ref).max(axis=1) / scale
    assert per_frame.max() <= 1e-4, int(per_frame.argmax())
    assert_parity(y / scale, ref / scale, "tiny-M synthesis M=%d m=%d" % (M, m))


@pytest.mark.parametrize("type_,otype", [(A, po.ANALYZER), (S, po.SYNTHESIZER)])
@pytest.mark.parametrize("M,p,S_,Q", [(16, 14, 9, 50), (16, 2, 4, 16), (16, 16, 700, 40), (8, 14, 13, 47), (8, 6, 4, 33),
                                      (8, 16, 3000, 20), (32, 14, 5, 100), (32, 4, 1, 17), (32, 16, 400, 37)])
def test_firpfbch_fused_tiny_M(M, p, S_, Q, type_, otype):
    """Fused critically sampled channelizers for M = 8 / 16 / 32 (analysis and synthesis): groups of 32/M streams on
    the fused kernel, the remainder on the generic one; ragged 16-frame batches; history carried across calls; more
    items than units (ranges that start mid-stream, warm-up items for the synthesiser)."""
    rng = np.random.default_rng(p * 100 + S_ + M)
    h = rng.standard_normal(M * p).astype(np.float32)
    x = _rand_c(rng, S_ * Q * M).reshape(S_, Q * M)
    q = yb.FirPfbCh.new(type_, M, p, h, n_streams=S_)
    cuts = [0, Q // 2 + 3, Q // 2 + 4, Q]
    y = np.concatenate([q.execute_block(np.ascontiguousarray(x[:, a * M: b * M])).reshape(S_, -1) for a, b in zip(cuts, cuts[1:])], axis=1)
    if S_ >= 32 // M and Q - cuts[2] >= 16:
        assert q.last_path() == 2
    ref = np.stack([po.FirPfbCh.new(otype, M, p, h).execute_block(x[s]) for s in range(S_)])
    scale = max(1.0, np.abs(ref).max())
    per_stream = np.abs(y - ref).max(axis=1) / scale
    assert per_stream.max() <= 1e-4, int(per_stream.argmax())
    assert_parity(y / scale, ref / scale, "tiny firpfbch type=%d M=%d p=%d S=%d" % (int(type_), M, p, S_))
