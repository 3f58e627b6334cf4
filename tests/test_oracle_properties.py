"""Acceptance tests for the channelizer restatement.

The reference holds no implementation and no vectors for firpfbch / firpfbch2 (PARITY
UNPINNED, see oracle/yagi_oracle.h); what it does hold are the upstream test *names*
(LIQUID_COMPAT.md:1765-1798).  These tests port the properties those autotests assert, plus
the reference's own archetypes: _config (firdecim.rs:216-244), _copy (firpfb.rs:361-396),
block-vs-sample (firdecim.rs:246-279) and partitioned-stream (rresamp.rs:198-237).
"""
import numpy as np
import pytest

import closed_forms as cf
import stimulus
from oracle import pyoracle as po

A, S = po.ANALYZER, po.SYNTHESIZER


def _rand_c(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


# ------------------------------------------------------------------ firpfbch2
@pytest.mark.parametrize("M", [8, 16, 32, 64])
def test_firpfbch2_crcf_reconstruction(M):
    """autotest firpfbch2_crcf_n8/n16/n32/n64: analysis -> synthesis reconstructs with delay 2Mm - M/2 + 1."""
    m, as_ = 5, 60.0
    tol = 1e-3
    num_blocks = 8 * m * 2
    n = num_blocks * (M // 2)
    s = 1
    x = np.empty(n, dtype=np.complex64)
    for i in range(n):                       # unit-modulus pseudo-random sequence (LCG)
        s = (s * 524287) % 1031
        x[i] = np.exp(2j * np.pi * s / 1031.0)
    qa = po.FirPfbCh2.new_kaiser(A, M, m, as_)
    qs = po.FirPfbCh2.new_kaiser(S, M, m, as_)
    y = np.empty(n, dtype=np.complex64)
    for k in range(num_blocks):
        Y = qa.execute(x[k * M // 2:(k + 1) * M // 2])
        y[k * M // 2:(k + 1) * M // 2] = qs.execute(Y)
    D = 2 * M * m - M // 2 + 1
    assert np.abs(y[:D]).max() < tol
    assert np.abs(y[D:] - x[: n - D]).max() < tol


@pytest.mark.parametrize("M,m", [(2, 1), (4, 3), (6, 2), (16, 5), (12, 4), (64, 3), (256, 7)])
def test_firpfbch2_analysis_closed_form(M, m):
    rng = np.random.default_rng(M * 100 + m)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    K = 6 * m + 5
    x = _rand_c(rng, K * M // 2)
    q = po.FirPfbCh2.new(A, M, m, h)
    y = q.execute_block(x).reshape(K, M)
    ref = cf.firpfbch2_analysis(h, M, m, x)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() < 2e-5 * scale * np.sqrt(m)


@pytest.mark.parametrize("M,m", [(2, 1), (4, 3), (6, 2), (16, 5), (12, 4), (64, 3)])
def test_firpfbch2_synthesis_closed_form(M, m):
    rng = np.random.default_rng(M * 100 + m + 1)
    h = rng.standard_normal(2 * M * m).astype(np.float32)
    K = 10 * m + 3
    X = _rand_c(rng, K * M)
    q = po.FirPfbCh2.new(S, M, m, h)
    y = q.execute_block(X)
    ref = cf.firpfbch2_synthesis(h, M, m, X)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() < 2e-5 * scale * np.sqrt(m)


def test_firpfbch2_analysis_is_mix_filter_decimate():
    """y_k[c] = ((-1)^{ck}/M) sum_tau h[tau] e^{+j2pi c tau/M} s[t_k - tau]  (Appendix A.3)."""
    M, m = 8, 3
    rng = np.random.default_rng(3)
    h = rng.standard_normal(2 * M * m)
    K = 20
    x = _rand_c(rng, K * M // 2)
    y = po.FirPfbCh2.new(A, M, m, h.astype(np.float32)).execute_block(x).reshape(K, M)
    hh = h.astype(np.float32).astype(np.float64)
    tau = np.arange(2 * M * m)
    for c in range(M):
        g = hh * np.exp(2j * np.pi * c * tau / M)
        full = np.convolve(x.astype(np.complex128), g)
        for k in range(K):
            tk = (k + 1) * M // 2 - 1
            want = ((-1) ** (c * k)) * full[tk] / M
            assert abs(y[k, c] - want) < 1e-5 * np.abs(full).max()


def test_firpfbch2_crcf_copy():
    """autotest firpfbch2_crcf_copy: clone mid-stream, outputs bit-identical afterwards."""
    M, m = 16, 4
    rng = np.random.default_rng(5)
    for type_, nin in ((A, M // 2), (S, M)):
        q = po.FirPfbCh2.new_kaiser(type_, M, m, 60.0)
        for _ in range(7):                     # odd number of frames: the parity flag must be cloned too
            q.execute(_rand_c(rng, nin))
        c = q.clone()
        for _ in range(24):
            v = _rand_c(rng, nin)
            np.testing.assert_array_equal(q.execute(v), c.execute(v))


def test_firpfbch2_crcf_config():
    """autotest firpfbch2_crcf_config: invalid type / M < 2 / odd M / m < 1 are Config errors."""
    for bad in [(77, 76, 12), (A, 0, 12), (A, 17, 12), (A, 76, 0)]:
        with pytest.raises(ValueError):
            po.FirPfbCh2.new_kaiser(bad[0], bad[1], bad[2], 60.0)
    with pytest.raises(ValueError):
        po.FirPfbCh2.new(A, 8, 3, np.zeros(2 * 8 * 3 - 1, dtype=np.float32))     # prototype too short
    q = po.FirPfbCh2.new_kaiser(A, 76, 12, 60.0)
    assert (po.lib().orc_firpfbch2_crcf_get_type(q._q), po.lib().orc_firpfbch2_crcf_get_M(q._q), po.lib().orc_firpfbch2_crcf_get_m(q._q)) == (A, 76, 12)


def test_firpfbch2_block_vs_sample_and_reset():
    M, m = 32, 3
    rng = np.random.default_rng(9)
    x = _rand_c(rng, 41 * M // 2)
    q = po.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    y_block = q.execute_block(x)
    q.reset()
    y_frames = np.concatenate([q.execute(x[k * M // 2:(k + 1) * M // 2]) for k in range(41)])
    np.testing.assert_array_equal(y_block, y_frames)
    q.reset()                                   # after an odd number of frames: reset must clear the flag
    np.testing.assert_array_equal(q.execute_block(x), y_block)


def test_firpfbch2_partitioned_stream():
    """rresamp.rs:198-237 archetype: a second object primed with the halo continues the stream.
    Shard boundary at an even frame, halo = (4m-1) * M/2 samples."""
    M, m = 16, 5
    K, k0 = 60, 26
    x = stimulus.noise_plus_tones(0, K * M // 2, M)
    whole = po.FirPfbCh2.new_kaiser(A, M, m, 60.0).execute_block(x).reshape(K, M)
    halo_frames = 4 * m - 1
    assert halo_frames % 2 == 1
    # prime with halo_frames+1 frames so that priming ends on an even frame count
    q1 = po.FirPfbCh2.new_kaiser(A, M, m, 60.0)
    start = (k0 - (halo_frames + 1)) * M // 2
    q1.execute_block(x[start: k0 * M // 2])
    part = q1.execute_block(x[k0 * M // 2:]).reshape(K - k0, M)
    np.testing.assert_allclose(part, whole[k0:], atol=1e-6)


# ------------------------------------------------------------------- firpfbch
def test_firpfbch_crcf_analysis():
    """autotest firpfbch_crcf_analysis: M=4, p=5: channelizer == mix down, firfilt, decimate (phase M-1)."""
    M, p, tol = 4, 5, 1e-4
    rng = np.random.default_rng(11)
    h = rng.standard_normal(M * p).astype(np.float32)
    nsym = 12
    x = _rand_c(rng, nsym * M)
    y = po.FirPfbCh.new(A, M, p, h).execute_block(x).reshape(nsym, M)
    t = np.arange(x.size)
    for c in range(M):
        mixed = (x.astype(np.complex128) * np.exp(-2j * np.pi * c * t / M)).astype(np.complex64)
        filt = po.firfilt_crcf(h, mixed)
        np.testing.assert_allclose(y[:, c], filt[M - 1::M], atol=tol)


def test_firpfbch_crcf_synthesis():
    """autotest firpfbch_crcf_synthesis: dual: upsample by M, firfilt, mix up, sum."""
    M, p, tol = 4, 5, 1e-4
    rng = np.random.default_rng(12)
    h = rng.standard_normal(M * p).astype(np.float32)
    nsym = 12
    X = _rand_c(rng, nsym * M).reshape(nsym, M)
    y = po.FirPfbCh.new(S, M, p, h).execute_block(X.reshape(-1))
    t = np.arange(nsym * M)
    ref = np.zeros(nsym * M, dtype=np.complex128)
    for c in range(M):
        up = np.zeros(nsym * M, dtype=np.complex64)
        up[::M] = X[:, c]
        ref += po.firfilt_crcf(h, up).astype(np.complex128) * np.exp(2j * np.pi * c * t / M)
    np.testing.assert_allclose(y, ref, atol=tol)


@pytest.mark.parametrize("M,p", [(1, 3), (4, 5), (5, 4), (16, 6), (64, 14)])
def test_firpfbch_closed_forms(M, p):
    rng = np.random.default_rng(M + p)
    h = rng.standard_normal(M * p).astype(np.float32)
    Q = 3 * p + 2
    x = _rand_c(rng, Q * M)
    ya = po.FirPfbCh.new(A, M, p, h).execute_block(x).reshape(Q, M)
    ra = cf.firpfbch_analysis(h, M, p, x)
    assert np.abs(ya - ra).max() < 2e-5 * np.abs(ra).max() * np.sqrt(p)
    ys = po.FirPfbCh.new(S, M, p, h).execute_block(x)
    rs = cf.firpfbch_synthesis(h, M, p, x)
    assert np.abs(ys - rs).max() < 2e-5 * np.abs(rs).max() * np.sqrt(p)


def test_firpfbch_crcf_config_and_copy():
    for bad in [(77, 8, 4), (A, 0, 4), (A, 8, 0)]:
        with pytest.raises(ValueError):
            po.FirPfbCh.new(bad[0], bad[1], bad[2], np.zeros(64, dtype=np.float32))
    with pytest.raises(ValueError):
        po.FirPfbCh.new_kaiser(A, 8, 0, 60.0)
    rng = np.random.default_rng(13)
    for type_ in (A, S):
        q = po.FirPfbCh.new_kaiser(type_, 8, 3, 60.0)
        for _ in range(5):
            q.execute(_rand_c(rng, 8))
        c = q.clone()
        for _ in range(10):
            v = _rand_c(rng, 8)
            np.testing.assert_array_equal(q.execute(v), c.execute(v))


# -------------------------------------------------------------------- firfilt
def test_firfilt_block_state_continuity():
    rng = np.random.default_rng(14)
    h = po.fir_design_kaiser(63, 0.25, 60.0)
    x = _rand_c(rng, 1000)
    whole = po.firfilt_crcf(h, x)
    q = po.FirFilt(h)
    parts = np.concatenate([q.execute_block(x[a:b]) for a, b in ((0, 1), (1, 64), (64, 65), (65, 700), (700, 1000))])
    np.testing.assert_allclose(parts, whole, atol=1e-6)
    ref = cf.firfilt(h, x)
    assert np.abs(whole - ref).max() < 1e-5


def test_stimulus_is_counter_based():
    a = stimulus.noise_plus_tones(0, 10000, 256)
    b = stimulus.noise_plus_tones(5000, 3000, 256)
    np.testing.assert_array_equal(a[5000:8000], b)
    c = stimulus.noise_plus_tones(-100, 300, 256)
    assert np.all(c[:100] == 0)
    np.testing.assert_array_equal(c[100:], a[:200])
