"""ctypes bindings for oracle/libyagi_oracle.so -- TEST INFRASTRUCTURE ONLY.

See oracle/yagi_oracle.h for provenance and the parity status of each function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libyagi_oracle.so")

ANALYZER, SYNTHESIZER = 0, 1
OK, EINTERNAL, ECONFIG, EVALUE, ERANGE, EMODE, ENOCONV = range(7)


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("core.c", "filters.c", "channelizer.c", "cpu_bench.c", "yagi_oracle.h")]
    def stale():
        return (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale():
        # several ranks of one job may get here at once: build under a lock, and re-check once it is held
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.check_call(["make", "-C", _HERE, "-B", "libyagi_oracle.so"], stdout=subprocess.DEVNULL)
            finally:
                fcntl.flock(lk, fcntl.LOCK_UN)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u32, f32, sz, i32 = C.c_void_p, C.c_uint32, C.c_float, C.c_size_t, C.c_int
        L.orc_lngammaf.restype = f32; L.orc_lngammaf.argtypes = [f32]
        L.orc_besseli0f.restype = f32; L.orc_besseli0f.argtypes = [f32]
        L.orc_sincf.restype = f32; L.orc_sincf.argtypes = [f32]
        L.orc_kaiser_beta_as.restype = f32; L.orc_kaiser_beta_as.argtypes = [f32]
        L.orc_kaiser.restype = i32; L.orc_kaiser.argtypes = [u32, u32, f32, vp]
        L.orc_fir_design_kaiser.restype = i32; L.orc_fir_design_kaiser.argtypes = [u32, f32, f32, f32, vp]
        L.orc_window_create.restype = vp; L.orc_window_create.argtypes = [u32]
        L.orc_window_destroy.argtypes = [vp]
        L.orc_window_reset.argtypes = [vp]
        L.orc_window_push.argtypes = [vp, C.c_uint64]       # ocf32 by value == 8 bytes in one INTEGER-class reg? no: see push()
        L.orc_window_read.restype = vp; L.orc_window_read.argtypes = [vp]
        L.orc_window_len.restype = u32; L.orc_window_len.argtypes = [vp]
        L.orc_window_allocated.restype = u32; L.orc_window_allocated.argtypes = [vp]
        L.orc_fft_create.restype = vp; L.orc_fft_create.argtypes = [u32, i32]
        L.orc_fft_destroy.argtypes = [vp]
        L.orc_fft_run.argtypes = [vp, vp, vp]
        L.orc_firfilt_crcf_create.restype = i32; L.orc_firfilt_crcf_create.argtypes = [vp, sz, vp]
        L.orc_firfilt_crcf_destroy.argtypes = [vp]
        L.orc_firfilt_crcf_reset.argtypes = [vp]
        L.orc_firfilt_crcf_set_scale.argtypes = [vp, f32]
        L.orc_firfilt_crcf_execute_block.restype = i32; L.orc_firfilt_crcf_execute_block.argtypes = [vp, vp, sz, vp]
        L.orc_firdecim_crcf_create.restype = i32; L.orc_firdecim_crcf_create.argtypes = [u32, vp, sz, vp]
        L.orc_firdecim_crcf_destroy.argtypes = [vp]
        L.orc_firpfb_rrrf_create.restype = i32; L.orc_firpfb_rrrf_create.argtypes = [u32, vp, sz, vp]
        L.orc_firpfb_rrrf_destroy.argtypes = [vp]
        L.orc_firpfb_rrrf_push.argtypes = [vp, f32]
        L.orc_firpfb_rrrf_execute.restype = i32; L.orc_firpfb_rrrf_execute.argtypes = [vp, u32, vp]
        for fam in ("firpfbch2", "firpfbch"):
            g = lambda n: getattr(L, f"orc_{fam}_crcf_{n}")
            g("create").restype = i32; g("create").argtypes = [i32, u32, u32, vp, sz, vp]
            g("create_kaiser").restype = i32; g("create_kaiser").argtypes = [i32, u32, u32, f32, vp]
            g("clone").restype = i32; g("clone").argtypes = [vp, vp]
            g("destroy").argtypes = [vp]
            g("reset").argtypes = [vp]
            g("execute").restype = i32; g("execute").argtypes = [vp, vp, vp]
            g("execute_block").restype = i32; g("execute_block").argtypes = [vp, vp, sz, vp]
            g("taps").restype = vp; g("taps").argtypes = [vp, vp]
        L.orc_bench_firpfbch2_analysis.restype = C.c_double
        L.orc_bench_firpfbch2_analysis.argtypes = [u32, u32, f32, vp, sz, u32, u32, vp]
        _lib = L
    return _lib


def _cf(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- L0 design
def fir_design_kaiser(n: int, fc: float, as_: float = 60.0, mu: float = 0.0) -> np.ndarray:
    h = np.zeros(n, dtype=np.float32)
    rc = lib().orc_fir_design_kaiser(n, fc, as_, mu, _p(h))
    if rc:
        raise ValueError(f"orc_fir_design_kaiser -> {rc}")
    return h


def kaiser(i: int, wlen: int, beta: float) -> float:
    out = C.c_float()
    rc = lib().orc_kaiser(i, wlen, beta, C.byref(out))
    if rc:
        raise ValueError(f"orc_kaiser -> {rc}")
    return out.value


# ---------------------------------------------------------------- L1
class Window:
    """src/buffer/window.rs"""

    def __init__(self, n: int):
        self._h = lib().orc_window_create(n)
        if not self._h:
            raise ValueError("window size must be greater than zero")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_window_destroy(self._h)
            self._h = None

    def push(self, v: complex):
        # struct {float,float} by value is passed in one SSE register on x86-64 SysV;
        # ctypes cannot express that portably through c_uint64, so use a tiny struct type.
        _push(self._h, v)

    def read(self) -> np.ndarray:
        n = lib().orc_window_len(self._h)
        ptr = lib().orc_window_read(self._h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(2 * n,)).copy().view(np.complex64)

    def reset(self):
        lib().orc_window_reset(self._h)

    @property
    def allocated(self) -> int:
        return lib().orc_window_allocated(self._h)


class _OCF32(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


def _push(h, v: complex):
    L = lib()
    f = L.orc_window_push
    f.argtypes = [C.c_void_p, _OCF32]
    f.restype = None
    f(h, _OCF32(float(np.float32(v.real)), float(np.float32(v.imag))))


def fft(x, backward: bool = False) -> np.ndarray:
    """Fft::run (src/fft/mod.rs:45-48): unnormalised in both directions."""
    x = _cf(x)
    y = np.empty_like(x)
    L = lib()
    p = L.orc_fft_create(x.size, 1 if backward else 0)
    L.orc_fft_run(p, _p(x), _p(y))
    L.orc_fft_destroy(p)
    return y


def dotprod_rcc(h, x) -> complex:
    L = lib()
    f = L.orc_dotprod_rcc
    f.restype = _OCF32
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    h = np.ascontiguousarray(h, dtype=np.float32)
    x = _cf(x)
    r = f(_p(h), _p(x), h.size)
    return complex(r.re, r.im)


# ---------------------------------------------------------------- L2
def firfilt_crcf(h, x, scale: float = 1.0) -> np.ndarray:
    """FirFilter::execute_block from a fresh (zero-state) object."""
    L = lib()
    h = np.ascontiguousarray(h, dtype=np.float32)
    x = _cf(x)
    y = np.empty_like(x)
    q = C.c_void_p()
    rc = L.orc_firfilt_crcf_create(_p(h), h.size, C.byref(q))
    if rc:
        raise ValueError(f"orc_firfilt_crcf_create -> {rc}")
    L.orc_firfilt_crcf_set_scale(q, scale)
    L.orc_firfilt_crcf_execute_block(q, _p(x), x.size, _p(y))
    L.orc_firfilt_crcf_destroy(q)
    return y


class FirFilt:
    def __init__(self, h, scale: float = 1.0):
        L = lib()
        h = np.ascontiguousarray(h, dtype=np.float32)
        self._q = C.c_void_p()
        rc = L.orc_firfilt_crcf_create(_p(h), h.size, C.byref(self._q))
        if rc:
            raise ValueError(f"orc_firfilt_crcf_create -> {rc}")
        L.orc_firfilt_crcf_set_scale(self._q, scale)

    def __del__(self):
        if getattr(self, "_q", None):
            lib().orc_firfilt_crcf_destroy(self._q)
            self._q = None

    def reset(self):
        lib().orc_firfilt_crcf_reset(self._q)

    def execute_block(self, x) -> np.ndarray:
        x = _cf(x)
        y = np.empty_like(x)
        lib().orc_firfilt_crcf_execute_block(self._q, _p(x), x.size, _p(y))
        return y


def firdecim_crcf(M: int, h, x) -> np.ndarray:
    L = lib()
    f = L.orc_firdecim_crcf_execute
    f.restype = _OCF32
    f.argtypes = [C.c_void_p, C.c_void_p]
    h = np.ascontiguousarray(h, dtype=np.float32)
    x = _cf(x)
    q = C.c_void_p()
    rc = L.orc_firdecim_crcf_create(M, _p(h), h.size, C.byref(q))
    if rc:
        raise ValueError(f"orc_firdecim_crcf_create -> {rc}")
    n = x.size // M
    y = np.empty(n, dtype=np.complex64)
    for k in range(n):
        blk = x[k * M:(k + 1) * M]
        r = f(q, _p(blk))
        y[k] = complex(r.re, r.im)
    L.orc_firdecim_crcf_destroy(q)
    return y


class FirPfbRrrf:
    def __init__(self, num_filters: int, h):
        h = np.ascontiguousarray(h, dtype=np.float32)
        self._q = C.c_void_p()
        rc = lib().orc_firpfb_rrrf_create(num_filters, _p(h), h.size, C.byref(self._q))
        if rc:
            raise ValueError(f"orc_firpfb_rrrf_create -> {rc}")

    def __del__(self):
        if getattr(self, "_q", None):
            lib().orc_firpfb_rrrf_destroy(self._q)
            self._q = None

    def push(self, x: float):
        lib().orc_firpfb_rrrf_push(self._q, float(x))

    def execute(self, i: int) -> float:
        y = C.c_float()
        rc = lib().orc_firpfb_rrrf_execute(self._q, i, C.byref(y))
        if rc:
            raise ValueError(f"filterbank index ({i}) exceeds maximum")
        return y.value


# ---------------------------------------------------------------- L3
class _Channelizer:
    _fam = ""

    def __init__(self, type_: int, M: int, m_or_p: int, h=None, as_: float | None = None, _handle=None):
        L = lib()
        self._q = C.c_void_p()
        self.type, self.M = type_, M
        if _handle is not None:
            self._q = _handle
            return
        if h is not None:
            h = np.ascontiguousarray(h, dtype=np.float32)
            rc = getattr(L, f"orc_{self._fam}_crcf_create")(type_, M, m_or_p, _p(h), h.size, C.byref(self._q))
        else:
            rc = getattr(L, f"orc_{self._fam}_crcf_create_kaiser")(type_, M, m_or_p, as_, C.byref(self._q))
        if rc:
            self._q = None
            raise ValueError(f"{self._fam} create -> status {rc}")

    def __del__(self):
        if getattr(self, "_q", None):
            getattr(lib(), f"orc_{self._fam}_crcf_destroy")(self._q)
            self._q = None

    def clone(self):
        out = C.c_void_p()
        rc = getattr(lib(), f"orc_{self._fam}_crcf_clone")(self._q, C.byref(out))
        if rc:
            raise RuntimeError("clone failed")
        c = type(self).__new__(type(self))
        c._q, c.type, c.M = out, self.type, self.M
        return c

    def reset(self):
        getattr(lib(), f"orc_{self._fam}_crcf_reset")(self._q)

    def taps(self) -> np.ndarray:
        n = C.c_size_t()
        ptr = getattr(lib(), f"orc_{self._fam}_crcf_taps")(self._q, C.byref(n))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n.value,)).copy()

    def _io(self):
        raise NotImplementedError

    def execute(self, x) -> np.ndarray:
        nin, nout = self._io()
        x = _cf(x)
        assert x.size == nin, (x.size, nin)
        y = np.empty(nout, dtype=np.complex64)
        getattr(lib(), f"orc_{self._fam}_crcf_execute")(self._q, _p(x), _p(y))
        return y

    def execute_block(self, x, n_frames: int | None = None) -> np.ndarray:
        nin, nout = self._io()
        x = _cf(x)
        if n_frames is None:
            assert x.size % nin == 0
            n_frames = x.size // nin
        y = np.empty(n_frames * nout, dtype=np.complex64)
        getattr(lib(), f"orc_{self._fam}_crcf_execute_block")(self._q, _p(x), n_frames, _p(y))
        return y


class FirPfbCh2(_Channelizer):
    """firpfbch2_crcf."""
    _fam = "firpfbch2"

    def _io(self):
        return (self.M // 2, self.M) if self.type == ANALYZER else (self.M, self.M // 2)

    @classmethod
    def new(cls, type_, M, m, h):
        return cls(type_, M, m, h=h)

    @classmethod
    def new_kaiser(cls, type_, M, m, as_=60.0):
        return cls(type_, M, m, as_=as_)


class FirPfbCh(_Channelizer):
    """firpfbch_crcf."""
    _fam = "firpfbch"

    def _io(self):
        return (self.M, self.M)

    @classmethod
    def new(cls, type_, M, p, h):
        return cls(type_, M, p, h=h)

    @classmethod
    def new_kaiser(cls, type_, M, m, as_=60.0):
        return cls(type_, M, m, as_=as_)


def bench_firpfbch2_analysis(M: int, m: int, as_: float, x: np.ndarray, n_per_thread: int, n_threads: int, passes: int = 1) -> float:
    """Seconds (slowest thread) for `passes` passes of one analyser per thread."""
    x = _cf(x)
    assert x.size >= n_per_thread * n_threads
    return lib().orc_bench_firpfbch2_analysis(M, m, as_, _p(x), n_per_thread, n_threads, passes, None)
