"""Host-side mirror of the filter pieces of the channelizer path, over the C ABI.

`fir_design_kaiser` (src/filter/fir/design/kaiser.rs:16-51) and `FirFilt` =
`FirFilter<Complex32, f32>` (src/filter/fir/firfilt.rs), batched over independent streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _buffers as B
from . import _lib
from .multichannel import _frames, _u32


def fir_design_kaiser(n: int, fc: float, as_: float = 60.0, mu: float = 0.0) -> np.ndarray:
    """`fir_design_kaiser(n, fc, as_, mu) -> Result<Vec<f32>>`."""
    h = np.zeros(max(int(n), 0), dtype=np.float32)
    _lib.check(_lib.lib().yg_fir_design_kaiser(_u32(n), float(fc), float(as_), float(mu), B.ptr(h) if h.size else None))
    return h


class FirFilt:
    """firfilt_crcf over `n_streams` independent streams sharing the taps; layout x[stream][n]."""

    def __init__(self, handle: C.c_void_p):
        self._q = handle
        n = C.c_size_t()
        _lib.check(_lib.lib().yg_firfilt_crcf_get_len(self._q, C.byref(n)))
        self._h_len = n.value
        self._S = None
        d = C.c_int32()
        _lib.check(_lib.lib().yg_firfilt_crcf_get_device(self._q, C.byref(d)))
        self._dev = d.value

    @classmethod
    def new(cls, h, n_streams: int = 1) -> "FirFilt":
        h = np.ascontiguousarray(h, dtype=np.float32)
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firfilt_crcf_create(B.ptr(h) if h.size else None, h.size, _u32(n_streams), C.byref(q)))
        f = cls(q)
        f._S = int(n_streams)
        return f

    @classmethod
    def new_kaiser(cls, n: int, fc: float, as_: float = 60.0, mu: float = 0.0, n_streams: int = 1) -> "FirFilt":
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firfilt_crcf_create_kaiser(_u32(n), float(fc), float(as_), float(mu), _u32(n_streams), C.byref(q)))
        f = cls(q)
        f._S = int(n_streams)
        return f

    def clone(self) -> "FirFilt":
        q = C.c_void_p()
        _lib.check(_lib.lib().yg_firfilt_crcf_clone(self._q, C.byref(q)))
        f = FirFilt(q)
        f._S = self._S
        return f

    def __del__(self):
        q = getattr(self, "_q", None)
        if q is not None and q.value:
            try:
                _lib.lib().yg_firfilt_crcf_destroy(q)
            except Exception:
                pass
            self._q = None

    def reset(self) -> None:
        _lib.check(_lib.lib().yg_firfilt_crcf_reset(self._q))

    def set_scale(self, scale: float) -> None:
        _lib.check(_lib.lib().yg_firfilt_crcf_set_scale(self._q, float(scale)))

    def get_scale(self) -> float:
        s = C.c_float()
        _lib.check(_lib.lib().yg_firfilt_crcf_get_scale(self._q, C.byref(s)))
        return s.value

    def len(self) -> int:
        return self._h_len

    def get_device(self) -> int:
        return self._dev

    def last_path(self) -> int:
        """0 none, 1 generic kernel, 2 register-blocked FFMA2 kernel, 4 tensor-core (tcgen05 3xTF32) kernel."""
        p = C.c_int32()
        _lib.check(_lib.lib().yg_firfilt_crcf_last_path(self._q, C.byref(p)))
        return p.value

    def execute_block(self, x, out=None):
        """x[stream][n] -> y[stream][n] (`execute_block(&mut self, x, y)`, firfilt.rs:267-278)."""
        L = _lib.lib()
        if B.is_torch_cuda(x):
            n = _frames(x.numel(), self._S)
            x = B.dev_in(x, n * self._S, "input")
            B.check_device(x, self._dev)
            y = B.dev_out(out, n * self._S, x)
            _lib.check(L.yg_firfilt_crcf_execute_block_dev(self._q, C.c_void_p(x.data_ptr()), n, C.c_void_p(y.data_ptr()), B.cur_stream(x)))
            return y
        n = _frames(np.asarray(x).size, self._S)
        xa = B.host_in(x, n * self._S, "input")
        y = B.host_out(out, n * self._S)
        _lib.check(L.yg_firfilt_crcf_execute_block(self._q, B.ptr(xa), n, B.ptr(y)))
        return y
