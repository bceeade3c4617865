"""Shared helpers for the parity tests."""
import numpy as np


def relerr(a, b):
    """Relative L2 error (the norm `north_star` states its tolerances in)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def isapprox(a, b, rtol, atol=0.0):
    """Julia `isapprox` for arrays."""
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) <= max(atol, rtol * max(np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel())))


# tolerances from BASELINE.json north_star
TOL_STEP = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
TOL_FFT = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}
