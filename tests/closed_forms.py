"""Stateless f64 closed forms of the channelizers -- tests only.

These are what the CUDA kernels compute; tests check closed form == stateful oracle, so
that any disagreement between GPU and oracle can be attributed (re-association noise vs a
real indexing/parity bug).  All take the full stream since reset; history before t = 0 is 0.
"""
import numpy as np


def firpfbch2_analysis(h, M, m, s, n_frames=None, k0=0):
    """y[k][c] for frames k0..k0+n_frames-1; `s` is the stream since reset (s[t<0] = 0)."""
    h = np.asarray(h, dtype=np.float64)[: 2 * M * m]
    s = np.asarray(s, dtype=np.complex128)
    M2 = M // 2
    if n_frames is None:
        n_frames = s.size // M2 - k0
    L = 2 * M * m
    sp = np.concatenate([np.zeros(L, dtype=np.complex128), s])
    out = np.empty((n_frames, M), dtype=np.complex128)
    for f in range(n_frames):
        k = k0 + f
        tk = (k + 1) * M2 - 1
        seg = sp[L + tk - L + 1: L + tk + 1][::-1]          # seg[tau] = s[tk - tau]
        V = (h * seg).reshape(2 * m, M).sum(axis=0)          # V[b] = sum_n h[b+nM] s[tk-b-nM]
        V = np.roll(V, (k & 1) * M2)
        out[f] = np.fft.ifft(V)                              # (1/M) * unnormalised backward
    return out


def firpfbch2_synthesis(h, M, m, X, k0=0, U_hist=None):
    """y[k*M/2 + i] for input frames X[k][c]; u_{k<0} = 0 unless U_hist (4m-1 frames) is given."""
    h = np.asarray(h, dtype=np.float64)[: 2 * M * m]
    X = np.asarray(X, dtype=np.complex128).reshape(-1, M)
    M2 = M // 2
    K = X.shape[0]
    U = 0.5 * M * np.fft.ifft(X, axis=1)                     # u = 1/2 * unnormalised backward
    nh = 4 * m - 1
    hist = np.zeros((nh, M), dtype=np.complex128) if U_hist is None else np.asarray(U_hist, dtype=np.complex128)
    Up = np.concatenate([hist, U], axis=0)
    y = np.zeros(K * M2, dtype=np.complex128)
    i = np.arange(M2)
    for f in range(K):
        k = k0 + f
        pi_k = (k & 1) * M2
        acc = np.zeros(M2, dtype=np.complex128)
        for l in range(4 * m):
            acc += h[i + l * M2] * Up[nh + f - l, (i + pi_k) % M]
        y[f * M2:(f + 1) * M2] = acc
    return y


def firpfbch_analysis(h, M, p, s):
    """y[q][c] = e^{+j 2 pi c / M} * IDFT_unnorm(V_q)[c],  V_q[b] = sum_n h[b+nM] s[qM+M-1-b-nM]."""
    h = np.asarray(h, dtype=np.float64)[: M * p]
    s = np.asarray(s, dtype=np.complex128)
    Q = s.size // M
    L = M * p
    sp = np.concatenate([np.zeros(L, dtype=np.complex128), s])
    out = np.empty((Q, M), dtype=np.complex128)
    rot = np.exp(2j * np.pi * np.arange(M) / M)
    for q in range(Q):
        tq = q * M + M - 1
        seg = sp[L + tq - L + 1: L + tq + 1][::-1]
        V = (h * seg).reshape(p, M).sum(axis=0)
        out[q] = rot * (M * np.fft.ifft(V))
    return out


def firpfbch_synthesis(h, M, p, X):
    """y[qM + i] = sum_n h[i + nM] * IDFT_unnorm(X_{q-n})[i]."""
    h = np.asarray(h, dtype=np.float64)[: M * p]
    X = np.asarray(X, dtype=np.complex128).reshape(-1, M)
    Q = X.shape[0]
    U = M * np.fft.ifft(X, axis=1)
    Up = np.concatenate([np.zeros((p - 1, M), dtype=np.complex128), U], axis=0)
    y = np.zeros(Q * M, dtype=np.complex128)
    i = np.arange(M)
    for q in range(Q):
        acc = np.zeros(M, dtype=np.complex128)
        for n in range(p):
            acc += h[i + n * M] * Up[p - 1 + q - n, i]
        y[q * M:(q + 1) * M] = acc
    return y


def firfilt(h, x):
    """y[n] = sum_k h[k] x[n-k], zero initial state, len(y) = len(x) (Appendix B)."""
    return np.convolve(np.asarray(x, dtype=np.complex128), np.asarray(h, dtype=np.float64))[: len(x)]
