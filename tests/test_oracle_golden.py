"""Pin the CPU oracle against the reference's own golden vectors.

Tolerances are the reference's: FFT 2e-4 on abs(err) plus the IFFT(FFT(x))/n
round trip (src/fft/mod.rs:125-151); firfilt / firdecim 1e-3
(src/filter/fir/firfilt.rs:851-919, firdecim.rs:297-317); dotprod 1e-3
(src/dotprod/mod.rs:455-524); firpfb 1e-4 (src/filter/fir/firpfb.rs:310-359).
"""
import numpy as np
import pytest

import golden_vectors as gv
from oracle import pyoracle as po

G = gv.load()


def test_fixture_matches_live_reference():
    live = gv.load_live()
    if live is None:
        pytest.skip("/root/reference not mounted")
    assert set(live) == set(G)
    for k in live:
        np.testing.assert_array_equal(live[k], G[k], err_msg=k)


@pytest.mark.parametrize("n", gv.FFT_SIZES)
def test_fft_forward_and_roundtrip(n):
    x = G[f"FFT_TEST_X{n}"]
    y_ref = G[f"FFT_TEST_Y{n}"]
    y = po.fft(x, backward=False)
    assert np.abs(y - y_ref).max() < 2e-4
    z = po.fft(y_ref, backward=True) / n            # Backward is unnormalised (src/fft/mod.rs:13-26)
    assert np.abs(z - x).max() < 2e-4


@pytest.mark.parametrize("case", gv.FIRFILT_CASES)
def test_firfilt_crcf(case):
    h = G[f"FIRFILT_CRCF_DATA_{case}_H"]
    x = G[f"FIRFILT_CRCF_DATA_{case}_X"]
    y_ref = G[f"FIRFILT_CRCF_DATA_{case}_Y"]
    y = po.firfilt_crcf(h, x)
    np.testing.assert_allclose(y, y_ref, atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize("M,case", gv.FIRDECIM_CASES)
def test_firdecim_crcf(M, case):
    h = G[f"FIRDECIM_CRCF_DATA_{case}_H"]
    x = G[f"FIRDECIM_CRCF_DATA_{case}_X"]
    y_ref = G[f"FIRDECIM_CRCF_DATA_{case}_Y"]
    y = po.firdecim_crcf(M, h, x)
    np.testing.assert_allclose(y, y_ref, atol=1e-3, rtol=1e-3)


def test_dotprod_crcf_known_answers():
    for tag in ("RAND01", "RAND02"):
        h, x, t = G[f"DOTPROD_CRCF_{tag}_H"], G[f"DOTPROD_CRCF_{tag}_X"], G[f"DOTPROD_CRCF_{tag}_Y"][0]
        y = po.dotprod_rcc(h, x)
        assert abs(y - t) < 1e-3
    y = po.dotprod_rcc(G["DOTPROD_CRCF_RAND01_H"][::-1], G["DOTPROD_CRCF_RAND01_X"])
    assert abs(y - G["DOTPROD_CRCF_RAND01_YREV"][0]) < 1e-3


def test_dotprod_crcf_vs_naive():
    # src/dotprod/mod.rs:526-546: n = 1..512 vs the naive sum
    rng = np.random.default_rng(7)
    for n in list(range(1, 40)) + [64, 100, 255, 256, 511, 512]:
        h = rng.random(n).astype(np.float32)
        x = (rng.random(n) + 1j * rng.random(n)).astype(np.complex64)
        y = po.dotprod_rcc(h, x)
        t = np.sum(h.astype(np.float64) * x.astype(np.complex128))
        assert abs(y - t) < 1e-3 * max(1.0, abs(t))


def test_firpfb_sub_filter_layout():
    f = po.FirPfbRrrf(4, G["FIRPFB_IMPULSE_H"])
    for v in G["FIRPFB_IMPULSE_NOISE"]:
        f.push(v)
    for i, t in enumerate(G["FIRPFB_IMPULSE_TEST"]):
        assert abs(f.execute(i) - t) < 1e-4 * max(1.0, abs(t))
    with pytest.raises(ValueError):
        f.execute(4)


def test_window_semantics():
    # src/buffer/window.rs:109-186 (push / read order / wrap) and Appendix B sizes
    assert po.Window(10).allocated == 16 + 10 - 1
    assert po.Window(14).allocated == 16 + 14 - 1
    assert po.Window(16).allocated == 32 + 16 - 1
    w = po.Window(10)
    np.testing.assert_array_equal(w.read(), np.zeros(10, dtype=np.complex64))
    for i in range(1, 5):
        w.push(complex(i, 0))
    np.testing.assert_array_equal(w.read().real, [0, 0, 0, 0, 0, 0, 1, 2, 3, 4])
    for i in range(100):
        w.push(complex(i, -i))
    np.testing.assert_array_equal(w.read().real, np.arange(90, 100))     # oldest first
    w.reset()
    np.testing.assert_array_equal(w.read(), np.zeros(10, dtype=np.complex64))
    with pytest.raises(ValueError):
        po.Window(0)


def test_kaiser_design_matches_f64_formula():
    # No numeric pin in the reference for the taps (only spectral masks); check the f32
    # restatement against an independent f64 evaluation of the same formula.
    for n, fc, as_ in [(161, 1.0 / 16, 60.0), (63, 0.25, 60.0), (3585, 1.0 / 256, 60.0), (41, 0.1, 80.0)]:
        h = po.fir_design_kaiser(n, fc, as_, 0.0)
        beta = 0.1102 * (as_ - 8.7)
        t = np.arange(n) - (n - 1) / 2.0
        r = 2 * t / (n - 1)
        ref = np.sinc(2 * fc * t) * np.i0(beta * np.sqrt(1 - r * r)) / np.i0(beta)
        assert np.abs(h - ref).max() < 5e-6
    for bad in [(0, 0.1, 60, 0), (10, 0.0, 60, 0), (10, 0.6, 60, 0), (10, 0.1, 0, 0), (10, 0.1, 60, -0.5), (10, 0.1, 60, 0.6)]:
        with pytest.raises(ValueError):
            po.fir_design_kaiser(*bad)
