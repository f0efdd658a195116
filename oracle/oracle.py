"""ctypes front-end of the CPU ORACLE (oracle/srb_oracle.c) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED by the reference (it has no golden vectors for this path); pinned by
tests/golden/kat_4x5.json (hand-evaluated from the cited reference lines) and tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. The product package singlerust_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ROW, COLUMN = 0, 1  # src/shared/mod.rs:39-42


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "srb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_select_var_threshold.restype = C.c_uint64
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Compressed:
    """Host-layout compressed matrix exactly as nalgebra-sparse holds it: usize offsets/indices.

    fmt 'csr': major = row; fmt 'csc': major = column.
    """

    def __init__(self, fmt, nrows, ncols, offsets, indices, values):
        assert fmt in ("csr", "csc")
        self.fmt, self.nrows, self.ncols = fmt, int(nrows), int(ncols)
        self.offsets, self.indices = _u64(offsets), _u64(indices)
        values = np.ascontiguousarray(values)
        if values.dtype not in (np.float32, np.float64):
            # f64::from(v) is exact for every dtype the reference accepts (shared/mod.rs:110-129)
            values = values.astype(np.float64)
        self.values = values

    @property
    def nmajor(self):
        return self.nrows if self.fmt == "csr" else self.ncols

    @property
    def nminor(self):
        return self.ncols if self.fmt == "csr" else self.nrows

    @property
    def nnz(self):
        return int(self.offsets[-1])

    def along_major(self, direction):
        return int((direction == ROW) == (self.fmt == "csr"))

    def out_len(self, direction):
        return self.nrows if direction == ROW else self.ncols

    def _suf(self):
        return "f32" if self.values.dtype == np.float32 else "f64"

    def _vt(self):
        return C.c_float if self.values.dtype == np.float32 else C.c_double

    def _args(self):
        return (C.c_uint64(self.nmajor), C.c_uint64(self.nminor), _p(self.offsets, C.c_uint64),
                _p(self.indices, C.c_uint64), _p(self.values, self._vt()))

    @classmethod
    def from_scipy(cls, m):
        import scipy.sparse as sp
        if sp.isspmatrix_csr(m) or isinstance(m, sp.csr_array):
            fmt = "csr"
        else:
            fmt = "csc"
        return cls(fmt, m.shape[0], m.shape[1], m.indptr, m.indices, m.data)


def number(m: Compressed, direction) -> np.ndarray:
    out = np.zeros(m.out_len(direction), dtype=np.uint32)
    lib().orc_number(C.c_uint64(m.nmajor), C.c_uint64(m.nminor), _p(m.offsets, C.c_uint64),
                     _p(m.indices, C.c_uint64), C.c_int(m.along_major(direction)), _p(out, C.c_uint32))
    return out


def _stat(name, m, direction):
    out = np.zeros(m.out_len(direction), dtype=np.float64)
    getattr(lib(), f"orc_{name}_{m._suf()}")(*m._args(), C.c_int(m.along_major(direction)), _p(out, C.c_double))
    return out


def sum_(m, direction):
    return _stat("sum", m, direction)


def variance(m, direction):
    return _stat("variance", m, direction)


def std_dev(m, direction):
    return _stat("std_dev", m, direction)


def min_max(m, direction):
    n = m.out_len(direction)
    mn, mx = np.zeros(n), np.zeros(n)
    getattr(lib(), f"orc_min_max_{m._suf()}")(*m._args(), C.c_int(m.along_major(direction)),
                                             _p(mn, C.c_double), _p(mx, C.c_double))
    return mn, mx


def normalize_total(m: Compressed, target: float, direction) -> Compressed:
    """scale_row / scale_col: returns a new matrix whose values are f64 (scale/mod.rs:82)."""
    out = np.zeros(m.nnz, dtype=np.float64)
    getattr(lib(), f"orc_normalize_total_{m._suf()}")(*m._args(), C.c_int(m.along_major(direction)),
                                                     C.c_double(target), _p(out, C.c_double))
    return Compressed(m.fmt, m.nrows, m.ncols, m.offsets, m.indices, out)


def log1p(m: Compressed) -> Compressed:
    v = m.values.copy()
    if v.dtype == np.float32:
        lib().orc_log1p_f32(C.c_uint64(v.size), _p(v, C.c_float))
    else:
        lib().orc_log1p_f64(C.c_uint64(v.size), _p(v, C.c_double))
    return Compressed(m.fmt, m.nrows, m.ncols, m.offsets, m.indices, v)


def select_hvg(variances: np.ndarray, n_top: int) -> np.ndarray:
    v = np.ascontiguousarray(variances, dtype=np.float64)
    out = np.zeros(min(n_top, v.size), dtype=np.uint64)
    rc = lib().orc_select_hvg(_p(v, C.c_double), C.c_uint64(v.size), C.c_uint64(n_top), _p(out, C.c_uint64))
    if rc != 0:
        raise ValueError("NaN variance: the reference panics in sort_by(partial_cmp().unwrap())")
    return out


def select_var_threshold(variances, t):
    v = np.ascontiguousarray(variances, dtype=np.float64)
    out = np.zeros(v.size, dtype=np.uint64)
    n = lib().orc_select_var_threshold(_p(v, C.c_double), C.c_uint64(v.size), C.c_double(t), _p(out, C.c_uint64))
    return out[:n].copy()


def densify_selected(m: Compressed, row_sel, col_sel) -> np.ndarray:
    rs, cs = _u64(row_sel), _u64(col_sel)
    dense = np.zeros((rs.size, cs.size), dtype=np.float64)
    getattr(lib(), f"orc_densify_selected_{m._suf()}")(
        *m._args(), C.c_int(m.fmt == "csc"), _p(rs, C.c_uint64), C.c_uint64(rs.size), _p(cs, C.c_uint64),
        C.c_uint64(cs.size), _p(dense, C.c_double))
    return dense


def number_chunk(chunk: Compressed, direction, reference: np.ndarray, major_offset: int = 0):
    """number_chunk_helper; major_offset=0 reproduces the reference (chunk-local index)."""
    lib().orc_number_chunk(C.c_uint64(chunk.nmajor), _p(chunk.offsets, C.c_uint64), _p(chunk.indices, C.c_uint64),
                           C.c_int(chunk.along_major(direction)), C.c_uint64(major_offset),
                           _p(reference, C.c_uint32), C.c_uint64(reference.size))


def sum_chunk(chunk: Compressed, direction, reference: np.ndarray, major_offset: int = 0):
    getattr(lib(), f"orc_sum_chunk_{chunk._suf()}")(
        C.c_uint64(chunk.nmajor), _p(chunk.offsets, C.c_uint64), _p(chunk.indices, C.c_uint64),
        _p(chunk.values, chunk._vt()), C.c_int(chunk.along_major(direction)), C.c_uint64(major_offset),
        _p(reference, C.c_double), C.c_uint64(reference.size))


def norm_log1p_genevar_omp(m: Compressed, target: float):
    """Threaded CPU baseline of normalise(Row)+log1p+per-gene moments (timing leg only)."""
    assert m.fmt == "csr" and m.values.dtype == np.float32
    out = np.zeros(m.nnz, dtype=np.float64)
    gs, gq, gv = np.zeros(m.ncols), np.zeros(m.ncols), np.zeros(m.ncols)
    gc = np.zeros(m.ncols, dtype=np.uint32)
    lib().orc_norm_log1p_genevar_omp_f32(C.c_uint64(m.nrows), C.c_uint64(m.ncols), _p(m.offsets, C.c_uint64),
                                         _p(m.indices, C.c_uint64), _p(m.values, C.c_float), C.c_double(target),
                                         _p(out, C.c_double), _p(gs, C.c_double), _p(gq, C.c_double),
                                         _p(gc, C.c_uint32), _p(gv, C.c_double))
    return Compressed("csr", m.nrows, m.ncols, m.offsets, m.indices, out), gs, gq, gc, gv


def synth_csr(seed: int, nrows: int, ncols: int, thr: np.ndarray, amp: np.ndarray, row0: int = 0,
              skew: bool = False) -> Compressed:
    """CPU twin of the device generator (singlerust_b200/csrc/synth.cu); bit-identical by construction."""
    thr = np.ascontiguousarray(thr, dtype=np.uint32)
    amp = np.ascontiguousarray(amp, dtype=np.uint32)
    off = np.zeros(nrows + 1, dtype=np.uint64)
    lib().orc_synth_count(C.c_uint32(seed), C.c_int(int(skew)), C.c_uint64(row0), C.c_uint64(nrows),
                          C.c_uint32(ncols), _p(thr, C.c_uint32), _p(amp, C.c_uint32), _p(off, C.c_uint64))
    nnz = int(off[-1])
    idx = np.zeros(nnz, dtype=np.uint64)
    val = np.zeros(nnz, dtype=np.float32)
    lib().orc_synth_fill(C.c_uint32(seed), C.c_int(int(skew)), C.c_uint64(row0), C.c_uint64(nrows),
                         C.c_uint32(ncols), _p(thr, C.c_uint32), _p(amp, C.c_uint32), _p(off, C.c_uint64),
                         _p(idx, C.c_uint64), _p(val, C.c_float))
    return Compressed("csr", nrows, ncols, off, idx, val)


def num_threads() -> int:
    return int(lib().orc_num_threads())
