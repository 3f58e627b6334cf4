#!/usr/bin/env python
"""tools/sanitize_small.py -- the smallest run that touches every fused kernel once (for compute-sanitizer memcheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

import yagi_b200 as yb

rng = np.random.default_rng(0)


def c(n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


paths = {}
for M, m, K in ((256, 7, 131), (64, 3, 523), (128, 2, 515), (1024, 2, 131), (512, 2, 131)):
    q = yb.FirPfbCh2.new_kaiser(yb.ANALYZER, M, m, 60.0)
    y = q.execute_block(c(K * M // 2))
    paths["ana M=%d" % M] = q.last_path()
    assert np.isfinite(y).all()
for M, m, K in ((256, 7, 161), (1024, 2, 161), (512, 2, 161)):
    q = yb.FirPfbCh2.new_kaiser(yb.SYNTHESIZER, M, m, 60.0)
    y = q.execute_block(c(K * M))
    paths["syn M=%d" % M] = q.last_path()
    assert np.isfinite(y).all()
for t in (yb.ANALYZER, yb.SYNTHESIZER):
    q = yb.FirPfbCh.new_kaiser(t, 64, 7, 60.0, n_streams=9)
    y = q.execute_block(c(9 * 37 * 64))
    assert np.isfinite(y).all()
f = yb.FirFilt.new_kaiser(63, 0.25, 60.0, 0.0, n_streams=3)
y = f.execute_block(c(3 * 5000))
assert np.isfinite(y).all()
print("sanitize_small ok", paths)
