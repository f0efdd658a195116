"""NumPy/SciPy ORACLE of filter_cells / filter_genes — TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).

Restates /root/reference/src/memory/processing/mod.rs:16-299 independently of the product's mirror: per-line stored-entry
counts and f64 sums (helper/csr.rs), ndarray-stats Linear quantiles (:148-174), the nine-way FlexValue match
(:32-83, :193-243) written out arm by arm, then a SciPy row/column slice standing in for IMAnnData::subset.
PARITY UNPINNED: the reference's own tests only assert "filtered count < original" (:385-417)."""
from __future__ import annotations

import numpy as np

F64_MIN, F64_MAX = -np.finfo(np.float64).max, np.finfo(np.float64).max


def linear_quantile(values, q):
    """ndarray_stats::interpolate::Linear: index = q (n-1); lower + (higher - lower) * frac."""
    v = np.sort(np.asarray(values, dtype=np.float64))
    pos = q * (v.size - 1)
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    return v[lo] + (v[hi] - v[lo]) * (pos - lo)


def mask(counts, sums, lower, upper):
    """lower / upper: ('Absolute', u32) | ('Relative', f64) | ('None', None)."""
    lp = linear_quantile(sums, lower[1]) if lower[0] == "Relative" else F64_MIN
    up = linear_quantile(sums, upper[1]) if upper[0] == "Relative" else F64_MAX
    out = np.zeros(len(sums), dtype=bool)
    for i in range(len(sums)):
        n, s = (int(counts[i]) if counts is not None else None), float(sums[i])
        k = (lower[0], upper[0])
        if k == ("Absolute", "Absolute"):
            out[i] = n >= lower[1] and n <= upper[1]
        elif k == ("Relative", "Relative"):
            out[i] = s >= lp and s <= up
        elif k == ("Absolute", "Relative"):
            out[i] = n >= lower[1] and s <= up
        elif k == ("Relative", "Absolute"):
            out[i] = s >= lp and n <= upper[1]
        elif k == ("Absolute", "None"):
            out[i] = n >= lower[1]
        elif k == ("None", "Absolute"):
            out[i] = n <= upper[1]
        elif k == ("Relative", "None"):
            out[i] = s >= lp
        elif k == ("None", "Relative"):
            out[i] = s <= up
        else:
            out[i] = True
    return out


def filter_matrix(a_csr, lower, upper, axis):
    """axis 0 = filter cells (rows), 1 = filter genes (columns). Returns (filtered scipy matrix, mask)."""
    import scipy.sparse as sp
    a = sp.csr_matrix(a_csr)
    nz = a.copy()
    nz.data = np.ones_like(nz.data)
    counts = np.asarray(nz.sum(axis=1 - axis)).ravel().astype(np.int64)   # stored entries (test matrices hold no explicit zeros)
    sums = np.asarray(a.astype(np.float64).sum(axis=1 - axis)).ravel()
    m = mask(counts, sums, lower, upper)
    return (a[m] if axis == 0 else a[:, m]), m
