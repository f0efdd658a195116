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
    (ArrayElemOp::iter, called at src/shared/statistics/mod.rs:24,66) or as a whole.

    Two stores: a SciPy matrix in memory, or an on-disk *chunk store* (`write_store` / `open_store`): a directory with
    `meta.txt` ("csr|csc nrows ncols"), `indptr.npy`, `indices.npy`, `data.npy`. The arrays are memory-mapped, so a chunk
    is read from disk only when the iterator reaches it (there is no HDF5 library in this image; the layout is what an
    .h5ad X group holds: data / indices / indptr)."""

    def __init__(self, scipy_csr_or_csc):
        import scipy.sparse as sp
        self._m = scipy_csr_or_csc
        self.is_csr = sp.isspmatrix_csr(self._m) or isinstance(self._m, sp.csr_array)

    @staticmethod
    def write_store(path, m, index_dtype=np.uint64):
        """Write a SciPy CSR / CSC matrix as a chunk store (canonical form: indices sorted within each line)."""
        import os
        import scipy.sparse as sp
        is_csr = sp.isspmatrix_csr(m) or isinstance(m, sp.csr_array)
        m = m.copy()
        m.sort_indices()
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "meta.txt"), "w") as f:
            f.write(f"{'csr' if is_csr else 'csc'} {m.shape[0]} {m.shape[1]}\n")
        np.save(os.path.join(path, "indptr.npy"), m.indptr.astype(index_dtype))
        np.save(os.path.join(path, "indices.npy"), m.indices.astype(index_dtype))
        np.save(os.path.join(path, "data.npy"), np.ascontiguousarray(m.data))

    @classmethod
    def open_store(cls, path):
        import os
        fmt, nrows, ncols = open(os.path.join(path, "meta.txt")).read().split()
        if fmt not in ("csr", "csc"):
            raise ValueError(f"chunk store format must be csr or csc, not {fmt!r}")
        self = cls.__new__(cls)
        self._m = None
        self.is_csr = fmt == "csr"
        self._shape = (int(nrows), int(ncols))
        self._indptr = np.load(os.path.join(path, "indptr.npy"), mmap_mode="r")
        self._indices = np.load(os.path.join(path, "indices.npy"), mmap_mode="r")
        self._data = np.load(os.path.join(path, "data.npy"), mmap_mode="r")
        nmajor = self._shape[0] if self.is_csr else self._shape[1]
        if self._indptr.shape != (nmajor + 1,) or self._indices.shape != self._data.shape or int(self._indptr[-1]) != self._data.shape[0]:
            raise ValueError("chunk store arrays do not match meta.txt")
        return self

    @property
    def n_obs(self):
        return self._m.shape[0] if self._m is not None else self._shape[0]

    @property
    def n_vars(self):
        return self._m.shape[1] if self._m is not None else self._shape[1]

    def _slice_store(self, s, e):
        """Lines [s, e) of the on-disk store as a SciPy matrix (this is the disk read)."""
        import scipy.sparse as sp
        a, b = int(self._indptr[s]), int(self._indptr[e])
        indptr = np.asarray(self._indptr[s:e + 1]).astype(np.int64) - a
        indices, data = np.asarray(self._indices[a:b]).astype(np.int64), np.asarray(self._data[a:b])
        if self.is_csr:
            return sp.csr_matrix((data, indices, indptr), shape=(e - s, self._shape[1]))
        return sp.csc_matrix((data, indices, indptr), shape=(self._shape[0], e - s))

    def iter_chunks(self, chunk_size: int):
        """Yields (chunk, start, end) like anndata's chunk iterator: row chunks for CSR, column chunks for CSC."""
        if chunk_size <= 0:
            raise ValueError("chunk size must be positive")
        n = self.n_obs if self.is_csr else self.n_vars
        for s in range(0, n, chunk_size):
            e = min(s + chunk_size, n)
            if self._m is None:
                yield self._slice_store(s, e), s, e
            else:
                yield (self._m[s:e] if self.is_csr else self._m[:, s:e]), s, e

    def whole(self):
        if self._m is None:
            return self._slice_store(0, self._shape[0] if self.is_csr else self._shape[1])
        return self._m
