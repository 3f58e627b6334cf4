"""Parity metrics (BASELINE.json north_star): rel-RMS <= 1e-5 and max-abs <= 1e-4 per output sample, f32."""
import numpy as np

REL_RMS_TOL = 1e-5
MAX_ABS_TOL = 1e-4


def errors(y, ref):
    y = np.asarray(y).reshape(-1).astype(np.complex128)
    ref = np.asarray(ref).reshape(-1).astype(np.complex128)
    assert y.shape == ref.shape, (y.shape, ref.shape)
    d = y - ref
    den = np.sqrt(np.sum(np.abs(ref) ** 2))
    rel = float(np.sqrt(np.sum(np.abs(d) ** 2)) / den) if den > 0 else float(np.sqrt(np.sum(np.abs(d) ** 2)))
    return rel, float(np.abs(d).max()) if d.size else 0.0


def assert_parity(y, ref, what="", rel_tol=REL_RMS_TOL, abs_tol=MAX_ABS_TOL):
    rel, mx = errors(y, ref)
    assert rel <= rel_tol and mx <= abs_tol, "%s: rel-RMS %.3e (tol %.1e), max-abs %.3e (tol %.1e)" % (what, rel, rel_tol, mx, abs_tol)
    return rel, mx
