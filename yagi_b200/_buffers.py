"""Argument plumbing shared by the host-side mirrors: numpy host arrays or torch CUDA tensors."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .error import ConfigError, ValueError_


def is_torch_cuda(x) -> bool:
    mod = type(x).__module__
    if not mod.startswith("torch"):
        return False
    return bool(getattr(x, "is_cuda", False))


def host_in(x, n: int, what: str) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.complex64).reshape(-1)
    if a.size != n:
        raise ConfigError("%s length (%d) does not match the expected %d samples" % (what, a.size, n))
    return a


def host_out(out, n: int) -> np.ndarray:
    if out is None:
        return np.empty(n, dtype=np.complex64)
    if not (isinstance(out, np.ndarray) and out.dtype == np.complex64 and out.flags.c_contiguous):
        raise ValueError_("out must be a C-contiguous complex64 numpy array")
    if out.size != n:
        raise ConfigError("output length (%d) does not match the expected %d samples" % (out.size, n))
    return out


def ptr(a: np.ndarray) -> C.c_void_p:
    return a.ctypes.data_as(C.c_void_p)


def dev_in(x, n: int, what: str):
    import torch
    if x.dtype != torch.complex64 or not x.is_contiguous():
        raise ValueError_("%s must be a contiguous complex64 CUDA tensor" % what)
    if x.numel() != n:
        raise ConfigError("%s length (%d) does not match the expected %d samples" % (what, x.numel(), n))
    return x


def dev_out(out, n: int, like):
    import torch
    if out is None:
        return torch.empty(n, dtype=torch.complex64, device=like.device)
    if out.dtype != torch.complex64 or not out.is_contiguous() or out.device != like.device:
        raise ValueError_("out must be a contiguous complex64 CUDA tensor on the input's device")
    if out.numel() != n:
        raise ConfigError("output length (%d) does not match the expected %d samples" % (out.numel(), n))
    return out


def check_device(x, dev: int, what: str = "input") -> None:
    """A handle is bound to the CUDA device it was created on; tensors from another device would be
    dereferenced there (or fail with an opaque CUDA error), so refuse them up front."""
    idx = x.device.index
    if idx is None:
        import torch
        idx = torch.cuda.current_device()
    if idx != dev:
        raise ValueError_("%s lives on cuda:%d but this object is bound to cuda:%d" % (what, idx, dev))


def cur_stream(x) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


class PinnedArray:
    """Page-locked complex64 host buffer from yg_host_alloc, exposed as a numpy view."""

    def __init__(self, n: int):
        from . import _lib
        self._p = C.c_void_p()
        _lib.check(_lib.lib().yg_host_alloc(C.byref(self._p), n * 8))
        buf = (C.c_float * (2 * n)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.complex64)

    def close(self):
        if getattr(self, "_p", None) is not None and self._p.value:
            from . import _lib
            self.array = None
            _lib.lib().yg_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
