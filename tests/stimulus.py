"""Seeded synthetic stimulus: complex noise plus tones -- shared by tests and bench.

yagi's own randnf is unseeded (src/random/normal.rs:9-22), so the stimulus is ours: a
counter-based generator keyed by (seed, t) so any shard can regenerate any index range
bit-identically.  numpy's Philox bit generator is counter based; we advance it to the
requested offset (4 x 64-bit words per 256-bit block => jump by blocks).
"""
import numpy as np

SEED = 0x5EED0001
TONES = ((1.0, 3.0, 0.0), (0.5, -17.25, 0.7))      # (amplitude, cycles per M samples, phase)
TONES_ABS = ((0.25, 0.123, 1.9), (0.1, -0.377, 2.6))  # (amplitude, cycles/sample, phase)
SIGMA = 0.1

_BLOCK = 4096          # samples per independently keyed block


def _noise_block(b: int) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(key=SEED, counter=[0, 0, 0, b]))
    v = rng.standard_normal(2 * _BLOCK, dtype=np.float32)
    return (v[0::2] + 1j * v[1::2]).astype(np.complex64)


def noise_plus_tones(t0: int, n: int, M: int = 256, stream: int = 0) -> np.ndarray:
    """cf32 samples s[t0 .. t0+n) of stream `stream`; t < 0 gives zeros (pre-reset history)."""
    out = np.zeros(n, dtype=np.complex64)
    lo = max(t0, 0)
    hi = t0 + n
    if hi <= lo:
        return out
    b0, b1 = lo // _BLOCK, (hi - 1) // _BLOCK
    parts = [_noise_block(b + stream * (1 << 40)) for b in range(b0, b1 + 1)]
    noise = np.concatenate(parts)[lo - b0 * _BLOCK: hi - b0 * _BLOCK]
    t = np.arange(lo, hi, dtype=np.float64)
    sig = np.zeros(hi - lo, dtype=np.complex128)
    for a, cyc, ph in TONES:
        sig += a * np.exp(1j * (2 * np.pi * (cyc / M) * t + ph))
    for a, f, ph in TONES_ABS:
        sig += a * np.exp(1j * (2 * np.pi * f * t + ph))
    out[lo - t0:] = (SIGMA * noise.astype(np.complex128) + sig).astype(np.complex64)
    return out
