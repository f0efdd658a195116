"""src/backed/statistics/mod.rs:5-45 — compute_number / compute_sum over a backed AnnData, Whole or Chunked(n).

Chunked mode drives the device chunk accumulator (srb_stream_*), which replaces shared::statistics::{number,sum}::
chunked (src/shared/statistics/mod.rs:17-41, 59-83). Unlike the reference, Row-direction results are placed at the
chunk's global offset (the reference ignores it — SURVEY §10 — and returns garbage for Direction::Row)."""
from __future__ import annotations

import numpy as np

from .. import _ffi
from ..anndata import BackedAnnData
from ..shared import ComputationMode, Direction


def _stream(ctx, adata: BackedAnnData, chunk: int) -> _ffi.ChunkStream:
    fmt = _ffi.CSR if adata.is_csr else _ffi.CSC
    st = _ffi.ChunkStream(ctx, fmt, adata.n_obs, adata.n_vars)
    for ch, _s, _e in adata.iter_chunks(chunk):
        ch = ch.tocsr() if adata.is_csr else ch.tocsc()
        ch.sort_indices()
        st.push(ch.indptr, ch.indices, ch.data)
    return st


def compute_number(ctx: _ffi.Context, adata: BackedAnnData, direction: Direction, mode: ComputationMode) -> np.ndarray:
    if mode.is_whole:
        m = _ffi.DeviceMatrix.from_scipy(ctx, adata.whole())
        return m.number(int(direction))
    return _stream(ctx, adata, mode.chunk).number(int(direction))


def compute_sum(ctx: _ffi.Context, adata: BackedAnnData, direction: Direction, mode: ComputationMode) -> np.ndarray:
    if mode.is_whole:
        m = _ffi.DeviceMatrix.from_scipy(ctx, adata.whole())
        return m.sum(int(direction))
    return _stream(ctx, adata, mode.chunk).sum(int(direction))


def compute_variance(ctx: _ffi.Context, adata: BackedAnnData, direction: Direction, mode: ComputationMode) -> np.ndarray:
    """Not in the reference (only number and sum exist for backed data); same accumulator, one more output."""
    if mode.is_whole:
        m = _ffi.DeviceMatrix.from_scipy(ctx, adata.whole())
        return m.variance(int(direction))
    return _stream(ctx, adata, mode.chunk).variance(int(direction))
