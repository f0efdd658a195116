"""Minimal in-memory container standing in for anndata_memory::IMAnnData on the Python side of the boundary.

X lives on the device (DeviceMatrix); obs / var are dicts of NumPy columns; obsm / varm dicts of arrays. Only what
the hot path touches is modelled (x(), n_obs, n_vars, obs/var columns written by qc_vars_inplace, obsm["X_pca"])."""
from __future__ import annotations

import numpy as np

from . import _ffi


class IMAnnData:
    def __init__(self, x: _ffi.DeviceMatrix, obs=None, var=None):
        self._x = x
        self.obs = dict(obs or {})
        self.var = dict(var or {})
        self.obsm, self.varm = {}, {}

    @classmethod
    def from_scipy(cls, ctx: _ffi.Context, m, index_dtype=np.uint64):
        return cls(_ffi.DeviceMatrix.from_scipy(ctx, m, index_dtype=index_dtype))

    def x(self) -> _ffi.DeviceMatrix:
        return self._x

    @property
    def n_obs(self):
        return self._x.shape[0]

    @property
    def n_vars(self):
        return self._x.shape[1]

    def deep_clone(self) -> "IMAnnData":
        """IMAnnData::deep_clone (used by normalize_total / log1p_transform): copy-on-write on the device."""
        c = IMAnnData(self._x.clone(), {k: np.copy(v) for k, v in self.obs.items()}, {k: np.copy(v) for k, v in self.var.items()})
        c.obsm = {k: np.copy(v) for k, v in self.obsm.items()}
        c.varm = {k: np.copy(v) for k, v in self.varm.items()}
        return c


class BackedAnnData:
    """Stand-in for anndata::AnnData<B: Backend>: X is only reachable through a row-chunk iterator
    (ArrayElemOp::iter, called at src/shared/statistics/mod.rs:24,66) or as a whole."""

    def __init__(self, scipy_csr_or_csc):
        import scipy.sparse as sp
        self._m = scipy_csr_or_csc
        self.is_csr = sp.isspmatrix_csr(self._m) or isinstance(self._m, sp.csr_array)

    @property
    def n_obs(self):
        return self._m.shape[0]

    @property
    def n_vars(self):
        return self._m.shape[1]

    def iter_chunks(self, chunk_size: int):
        """Yields (chunk, start, end) like anndata's chunk iterator: row chunks for CSR, column chunks for CSC."""
        n = self.n_obs if self.is_csr else self.n_vars
        for s in range(0, n, chunk_size):
            e = min(s + chunk_size, n)
            yield (self._m[s:e] if self.is_csr else self._m[:, s:e]), s, e

    def whole(self):
        return self._m
