"""Host-side mirror of yagi's `multichannel` module slot (src/lib.rs:27-28), over the C ABI.

`FirPfbCh2` / `FirPfbCh` follow the object protocol of every yagi filter struct
(`new*`, `reset`, `execute`, `execute_block`, getters, `Clone`; e.g.
src/filter/fir/firdecim.rs:38-57,124-126,179-205) so parity tests read like the
reference's own tests.  All arithmetic happens in libyagi_b200.so on the GPU.

Inputs may be numpy arrays (host path: copies inside the call) or torch CUDA tensors
(device path: asynchronous on torch's current stream).
"""
from __future__ import annotations

import ctypes as C
from enum import IntEnum

import numpy as np

from . import _buffers as B
from . import _lib


class FirPfbChType(IntEnum):
    """Upstream LIQUID_ANALYZER / LIQUID_SYNTHESIZER (2-variant enum idiom of msresamp2.rs:27-30)."""
    Analyzer = 0
    Synthesizer = 1


ANALYZER = FirPfbChType.Analyzer
SYNTHESIZER = FirPfbChType.Synthesizer


class FirPfbCh2:
    """firpfbch2_crcf: 2x oversampled polyphase filterbank channelizer.

    Analyzer: `num_channels/2` samples in -> `num_channels` out per frame; Synthesizer the reverse.
    """

    _P = "yg_firpfbch2_crcf_"

    def __init__(self, handle: C.c_void_p):
        self._q = handle
        L = _lib.lib()
        t, M, m = C.c_int32(), C.c_uint32(), C.c_uint32()
        _lib.check(L.yg_firpfbch2_crcf_get_type(self._q, C.byref(t)))
        _lib.check(L.yg_firpfbch2_crcf_get_M(self._q, C.byref(M)))
        _lib.check(L.yg_firpfbch2_crcf_get_m(self._q, C.byref(m)))
        self._type, self._M, self._m = FirPfbChType(t.value), M.value, m.value
        d = C.c_int32()
        _lib.check(L.yg_firpfbch2_crcf_get_device(self._q, C.byref(d)))
        self._dev = d.value

    # -- constructors ------------------------------------------------------------
    @classmethod
    def new(cls, type_, num_channels: int, m: int, h) -> "FirPfbCh2":
        """`new(type, M, m, h)`: M even >= 2, m >= 1, len(h) >= 2*M*m; else ConfigError."""
        h = np.ascontiguousarray(h, dtype=np.float32)
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_create(int(type_), _u32(num_channels), _u32(m), B.ptr(h), h.size, C.byref(q)))
        return cls(q)

    @classmethod
    def new_kaiser(cls, type_, num_channels: int, m: int, as_: float = 60.0) -> "FirPfbCh2":
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_create_kaiser(int(type_), _u32(num_channels), _u32(m), float(as_), C.byref(q)))
        return cls(q)

    def clone(self) -> "FirPfbCh2":
        """`#[derive(Clone)]`: deep copy including the stream state."""
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_clone(self._q, C.byref(q)))
        return FirPfbCh2(q)

    def __del__(self):
        q = getattr(self, "_q", None)
        if q is not None and q.value:
            try:
                _lib.lib().yg_firpfbch2_crcf_destroy(q)
            except Exception:
                pass
            self._q = None

    # -- getters -----------------------------------------------------------------
    def get_type(self) -> FirPfbChType:
        return self._type

    def get_num_channels(self) -> int:
        return self._M

    def get_m(self) -> int:
        return self._m

    def get_taps(self) -> np.ndarray:
        h = np.empty(2 * self._M * self._m, dtype=np.float32)
        _lib.check(_lib.lib().yg_firpfbch2_crcf_get_taps(self._q, B.ptr(h)))
        return h

    @property
    def samples_in_per_frame(self) -> int:
        return self._M // 2 if self._type == ANALYZER else self._M

    @property
    def samples_out_per_frame(self) -> int:
        return self._M if self._type == ANALYZER else self._M // 2

    # -- state -------------------------------------------------------------------
    def reset(self) -> None:
        _lib.check(_lib.lib().yg_firpfbch2_crcf_reset(self._q))

    def state_len(self) -> int:
        n = C.c_size_t()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_state_len(self._q, C.byref(n)))
        return n.value

    def get_state(self):
        """(history, flag): history oldest-first; analyser: last (4m-1)*M/2 input samples."""
        hist = np.empty(self.state_len(), dtype=np.complex64)
        flag = C.c_int32()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_get_state(self._q, B.ptr(hist), C.byref(flag)))
        return hist, flag.value

    def set_state(self, hist, flag: int) -> None:
        hist = B.host_in(hist, self.state_len(), "state")
        _lib.check(_lib.lib().yg_firpfbch2_crcf_set_state(self._q, B.ptr(hist), int(flag)))

    # -- execution ---------------------------------------------------------------
    def execute(self, x, out=None):
        """One frame (`execute(&mut self, x, y)`)."""
        return self.execute_block(x, 1, out)

    def execute_block(self, x, n: int | None = None, out=None):
        """`n` consecutive frames.  numpy in -> numpy out (host path); CUDA tensor in -> CUDA tensor out."""
        nin, nout = self.samples_in_per_frame, self.samples_out_per_frame
        L = _lib.lib()
        if B.is_torch_cuda(x):
            if n is None:
                n = _frames(x.numel(), nin)
            x = B.dev_in(x, n * nin, "input")
            B.check_device(x, self._dev)
            y = B.dev_out(out, n * nout, x)
            _lib.check(L.yg_firpfbch2_crcf_execute_block_dev(self._q, C.c_void_p(x.data_ptr()), n, C.c_void_p(y.data_ptr()), B.cur_stream(x)))
            return y
        if n is None:
            n = _frames(np.asarray(x).size, nin)
        xa = B.host_in(x, n * nin, "input")
        y = B.host_out(out, n * nout)
        _lib.check(L.yg_firpfbch2_crcf_execute_block(self._q, B.ptr(xa), n, B.ptr(y)))
        return y

    def last_path(self) -> int:
        """0 none, 1 generic kernels, 2 fused fast kernel (for tests / bench bookkeeping)."""
        p = C.c_int32()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_last_path(self._q, C.byref(p)))
        return p.value

    def get_device(self) -> int:
        """Index of the CUDA device this object is bound to (the device current at construction)."""
        return self._dev

    def set_kernel_timing(self, enable: bool) -> None:
        """Turn the two CUDA events recorded around the dominant kernel of each call off / on (default on)."""
        _lib.check(_lib.lib().yg_firpfbch2_crcf_set_kernel_timing(self._q, 1 if enable else 0))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_last_kernel_ms(self._q, C.byref(ms)))
        return ms.value

    def kernel_times_ms(self, cap: int = 64) -> np.ndarray:
        """Device durations (CUDA events on the launching stream) of the dominant kernel of the most
        recent execute_block calls on device pointers, oldest first (at most 64 are kept)."""
        ms = np.zeros(cap, dtype=np.float32)
        n = C.c_size_t()
        _lib.check(_lib.lib().yg_firpfbch2_crcf_kernel_times(self._q, B.ptr(ms), cap, C.byref(n)))
        return ms[: n.value].copy()

    def __repr__(self):
        return "FirPfbCh2 { type: %s, num_channels: %d, m: %d }" % (self._type.name, self._M, self._m)


class FirPfbCh:
    """firpfbch_crcf: critically sampled channelizer (Appendix A.2), batched over `n_streams`
    independent streams that share the taps.  Layout x[stream][frame][M]."""

    def __init__(self, handle: C.c_void_p):
        self._q = handle
        L = _lib.lib()
        t, M, p, s = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _lib.check(L.yg_firpfbch_crcf_get_type(self._q, C.byref(t)))
        _lib.check(L.yg_firpfbch_crcf_get_M(self._q, C.byref(M)))
        _lib.check(L.yg_firpfbch_crcf_get_p(self._q, C.byref(p)))
        _lib.check(L.yg_firpfbch_crcf_get_n_streams(self._q, C.byref(s)))
        self._type, self._M, self._p, self._S = FirPfbChType(t.value), M.value, p.value, s.value
        d = C.c_int32()
        _lib.check(L.yg_firpfbch_crcf_get_device(self._q, C.byref(d)))
        self._dev = d.value

    @classmethod
    def new(cls, type_, num_channels: int, p: int, h, n_streams: int = 1) -> "FirPfbCh":
        h = np.ascontiguousarray(h, dtype=np.float32)
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firpfbch_crcf_create(int(type_), _u32(num_channels), _u32(p), B.ptr(h), h.size, _u32(n_streams), C.byref(q)))
        return cls(q)

    @classmethod
    def new_kaiser(cls, type_, num_channels: int, m: int, as_: float = 60.0, n_streams: int = 1) -> "FirPfbCh":
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firpfbch_crcf_create_kaiser(int(type_), _u32(num_channels), _u32(m), float(as_), _u32(n_streams), C.byref(q)))
        return cls(q)

    def clone(self) -> "FirPfbCh":
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firpfbch_crcf_clone(self._q, C.byref(q)))
        return FirPfbCh(q)

    def last_path(self) -> int:
        """0 none, 1 generic kernels, 2 fused kernel for the bulk of the streams (for tests / bench bookkeeping)."""
        p = C.c_int32()
        _lib.check(_lib.lib().yg_firpfbch_crcf_last_path(self._q, C.byref(p)))
        return p.value

    def __del__(self):
        q = getattr(self, "_q", None)
        if q is not None and q.value:
            try:
                _lib.lib().yg_firpfbch_crcf_destroy(q)
            except Exception:
                pass
            self._q = None

    def get_type(self) -> FirPfbChType:
        return self._type

    def get_num_channels(self) -> int:
        return self._M

    def get_p(self) -> int:
        return self._p

    def get_n_streams(self) -> int:
        return self._S

    def get_device(self) -> int:
        return self._dev

    def get_taps(self) -> np.ndarray:
        h = np.empty(self._M * self._p, dtype=np.float32)
        _lib.check(_lib.lib().yg_firpfbch_crcf_get_taps(self._q, B.ptr(h)))
        return h

    def reset(self) -> None:
        _lib.check(_lib.lib().yg_firpfbch_crcf_reset(self._q))

    def execute(self, x, out=None):
        return self.execute_block(x, 1, out)

    def execute_block(self, x, n: int | None = None, out=None):
        per = self._M * self._S
        L = _lib.lib()
        if B.is_torch_cuda(x):
            if n is None:
                n = _frames(x.numel(), per)
            x = B.dev_in(x, n * per, "input")
            B.check_device(x, self._dev)
            y = B.dev_out(out, n * per, x)
            _lib.check(L.yg_firpfbch_crcf_execute_block_dev(self._q, C.c_void_p(x.data_ptr()), n, C.c_void_p(y.data_ptr()), B.cur_stream(x)))
            return y
        if n is None:
            n = _frames(np.asarray(x).size, per)
        xa = B.host_in(x, n * per, "input")
        y = B.host_out(out, n * per)
        _lib.check(L.yg_firpfbch_crcf_execute_block(self._q, B.ptr(xa), n, B.ptr(y)))
        return y

    def __repr__(self):
        return "FirPfbCh { type: %s, num_channels: %d, p: %d, n_streams: %d }" % (self._type.name, self._M, self._p, self._S)


def _u32(v) -> int:
    v = int(v)
    if v < 0 or v > 0xFFFFFFFF:
        from .error import ConfigError
        raise ConfigError("argument out of range: %d" % v)
    return v


def _frames(total: int, per: int) -> int:
    if per == 0 or total % per:
        from .error import ConfigError
        raise ConfigError("input length (%d) is not a whole number of frames of %d samples" % (total, per))
    return total // per
