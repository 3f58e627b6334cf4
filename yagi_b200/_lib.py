"""ctypes binding of the C ABI declared in include/yagi_b200.h.

The shared library is the product; this module only loads it.  If it is missing it is built
in-tree with nvcc; if that is impossible the import fails loudly -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build
from .error import from_status

_lib = None

SYMBOLS = """
yg_version yg_last_error yg_device_count yg_launch_count yg_channel_major_dev yg_host_alloc yg_host_free yg_fir_design_kaiser
yg_firpfbch2_crcf_create yg_firpfbch2_crcf_create_kaiser yg_firpfbch2_crcf_clone yg_firpfbch2_crcf_destroy
yg_firpfbch2_crcf_reset yg_firpfbch2_crcf_execute yg_firpfbch2_crcf_execute_block
yg_firpfbch2_crcf_execute_block_dev yg_firpfbch2_crcf_sync yg_firpfbch2_crcf_get_type yg_firpfbch2_crcf_get_M
yg_firpfbch2_crcf_get_m yg_firpfbch2_crcf_get_taps yg_firpfbch2_crcf_state_len yg_firpfbch2_crcf_get_state
yg_firpfbch2_crcf_set_state yg_firpfbch2_crcf_last_path yg_firpfbch2_crcf_last_kernel_ms
yg_firpfbch2_crcf_kernel_times yg_firpfbch2_crcf_set_kernel_timing yg_firpfbch2_crcf_get_device
yg_firpfbch_crcf_create yg_firpfbch_crcf_create_kaiser yg_firpfbch_crcf_clone yg_firpfbch_crcf_destroy
yg_firpfbch_crcf_reset yg_firpfbch_crcf_execute yg_firpfbch_crcf_execute_block yg_firpfbch_crcf_execute_block_dev
yg_firpfbch_crcf_sync yg_firpfbch_crcf_get_type yg_firpfbch_crcf_get_M yg_firpfbch_crcf_get_p
yg_firpfbch_crcf_get_n_streams yg_firpfbch_crcf_get_taps yg_firpfbch_crcf_last_path yg_firpfbch_crcf_get_device
yg_firfilt_crcf_create yg_firfilt_crcf_create_kaiser yg_firfilt_crcf_clone yg_firfilt_crcf_destroy
yg_firfilt_crcf_reset yg_firfilt_crcf_set_scale yg_firfilt_crcf_get_scale yg_firfilt_crcf_get_len
yg_firfilt_crcf_execute_block yg_firfilt_crcf_execute_block_dev yg_firfilt_crcf_sync yg_firfilt_crcf_get_device yg_firfilt_crcf_last_path
""".split()


def path() -> str:
    return _build.SO


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    if not os.path.exists(so):
        so = _build.build()
    L = C.CDLL(so)
    vp, u32, i32, f32, sz = C.c_void_p, C.c_uint32, C.c_int32, C.c_float, C.c_size_t
    for name in SYMBOLS:
        getattr(L, name).restype = i32
    L.yg_last_error.restype = C.c_char_p
    L.yg_last_error.argtypes = []
    L.yg_version.argtypes = []
    L.yg_device_count.argtypes = [vp]
    L.yg_launch_count.argtypes = [vp]
    L.yg_channel_major_dev.argtypes = [vp, sz, u32, vp, vp]
    L.yg_host_alloc.argtypes = [vp, sz]
    L.yg_host_free.argtypes = [vp]
    L.yg_fir_design_kaiser.argtypes = [u32, f32, f32, f32, vp]
    # firpfbch2
    L.yg_firpfbch2_crcf_create.argtypes = [i32, u32, u32, vp, sz, vp]
    L.yg_firpfbch2_crcf_create_kaiser.argtypes = [i32, u32, u32, f32, vp]
    L.yg_firpfbch2_crcf_clone.argtypes = [vp, vp]
    L.yg_firpfbch2_crcf_destroy.argtypes = [vp]
    L.yg_firpfbch2_crcf_reset.argtypes = [vp]
    L.yg_firpfbch2_crcf_execute.argtypes = [vp, vp, vp]
    L.yg_firpfbch2_crcf_execute_block.argtypes = [vp, vp, sz, vp]
    L.yg_firpfbch2_crcf_execute_block_dev.argtypes = [vp, vp, sz, vp, vp]
    L.yg_firpfbch2_crcf_sync.argtypes = [vp]
    for n in ("get_type", "get_M", "get_m", "get_taps", "state_len", "last_path", "last_kernel_ms"):
        getattr(L, "yg_firpfbch2_crcf_" + n).argtypes = [vp, vp]
    L.yg_firpfbch2_crcf_get_state.argtypes = [vp, vp, vp]
    L.yg_firpfbch2_crcf_set_state.argtypes = [vp, vp, i32]
    L.yg_firpfbch2_crcf_kernel_times.argtypes = [vp, vp, sz, vp]
    L.yg_firpfbch2_crcf_set_kernel_timing.argtypes = [vp, i32]
    for n in ("yg_firpfbch2_crcf_get_device", "yg_firpfbch_crcf_get_device", "yg_firfilt_crcf_get_device", "yg_firfilt_crcf_last_path"):
        getattr(L, n).argtypes = [vp, vp]
    # firpfbch
    L.yg_firpfbch_crcf_create.argtypes = [i32, u32, u32, vp, sz, u32, vp]
    L.yg_firpfbch_crcf_create_kaiser.argtypes = [i32, u32, u32, f32, u32, vp]
    L.yg_firpfbch_crcf_clone.argtypes = [vp, vp]
    L.yg_firpfbch_crcf_destroy.argtypes = [vp]
    L.yg_firpfbch_crcf_reset.argtypes = [vp]
    L.yg_firpfbch_crcf_execute.argtypes = [vp, vp, vp]
    L.yg_firpfbch_crcf_execute_block.argtypes = [vp, vp, sz, vp]
    L.yg_firpfbch_crcf_execute_block_dev.argtypes = [vp, vp, sz, vp, vp]
    L.yg_firpfbch_crcf_sync.argtypes = [vp]
    for n in ("get_type", "get_M", "get_p", "get_n_streams", "get_taps", "last_path"):
        getattr(L, "yg_firpfbch_crcf_" + n).argtypes = [vp, vp]
    # firfilt
    L.yg_firfilt_crcf_create.argtypes = [vp, sz, u32, vp]
    L.yg_firfilt_crcf_create_kaiser.argtypes = [u32, f32, f32, f32, u32, vp]
    L.yg_firfilt_crcf_clone.argtypes = [vp, vp]
    L.yg_firfilt_crcf_destroy.argtypes = [vp]
    L.yg_firfilt_crcf_reset.argtypes = [vp]
    L.yg_firfilt_crcf_set_scale.argtypes = [vp, f32]
    L.yg_firfilt_crcf_get_scale.argtypes = [vp, vp]
    L.yg_firfilt_crcf_get_len.argtypes = [vp, vp]
    L.yg_firfilt_crcf_execute_block.argtypes = [vp, vp, sz, vp]
    L.yg_firfilt_crcf_execute_block_dev.argtypes = [vp, vp, sz, vp, vp]
    L.yg_firfilt_crcf_sync.argtypes = [vp]
    _lib = L
    return L


def launch_count() -> int:
    """Kernels launched by libyagi_b200.so in this process so far."""
    n = C.c_uint64()
    check(lib().yg_launch_count(C.byref(n)))
    return n.value


def check(status: int) -> None:
    """Map a status code to the matching exception (Rust: `Result<()>`)."""
    if status != 0:
        msg = lib().yg_last_error()
        raise from_status(status, msg.decode("utf-8", "replace") if msg else "status %d" % status)
