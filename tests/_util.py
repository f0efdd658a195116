"""Shared helpers for the test-suite."""
import numpy as np
import scipy.sparse as sp


def f(lst):
    """JSON list with 'nan'/'inf'/'-inf' strings -> float64 array."""
    return np.array([float(x) for x in lst], dtype=np.float64)


def random_csr(rng, n, m, density, dtype=np.float32, empty_rows=(), empty_cols=(), integer=True):
    """Canonical CSR (sorted, unique indices; no explicit zeros) with optional empty lines."""
    mask = rng.random((n, m)) < density
    mask[list(empty_rows), :] = False
    mask[:, list(empty_cols)] = False
    if integer:
        vals = rng.integers(1, 50, size=(n, m)).astype(np.float64)
    else:
        vals = rng.uniform(0.01, 50.0, size=(n, m))
    a = sp.csr_matrix(np.where(mask, vals, 0.0).astype(dtype))
    a.sort_indices()
    return a


def sign_align(a, ref):
    """Flip the sign of each column of `a` to best match `ref` (singular vectors are sign-ambiguous)."""
    a = np.array(a, dtype=np.float64, copy=True)
    s = np.sign(np.sum(a * ref, axis=0))
    s[s == 0] = 1.0
    return a * s


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / den) if a.size else 0.0
