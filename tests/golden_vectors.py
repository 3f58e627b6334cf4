"""Loader for the reference's golden vectors (tests only).

Prefers parsing /root/reference live (and checks the committed fixture agrees);
falls back to tests/golden/reference_vectors.npz where the reference is absent
(the GPU box).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_NPZ = os.path.join(_HERE, "golden", "reference_vectors.npz")
_cache = None


def load():
    global _cache
    if _cache is None:
        z = np.load(_NPZ)
        _cache = {k: z[k] for k in z.files}
    return _cache


def load_live():
    """Parse the reference directly; None when it is not mounted."""
    if not os.path.isdir("/root/reference/src"):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("extract_reference_vectors", os.path.join(_HERE, "golden", "extract_reference_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.collect()


FFT_SIZES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 16, 17, 20, 21, 22, 24, 26, 30, 32, 35, 36, 43, 48, 63, 64, 79, 92, 96,
             120, 130, 157, 192, 317, 509]
FIRFILT_CASES = ["H4X8", "H7X16", "H13X32", "H23X64"]
FIRDECIM_CASES = [(2, "M2H4X20"), (3, "M3H7X30"), (4, "M4H13X40"), (5, "M5H23X50")]
