"""Synthetic count-matrix tables: per-gene detection thresholds and value amplitudes.

Entry (cell i, gene j) of the synthetic matrix exists iff hash32(seed, i, j) < thr[j]; its value is a small
positive integer whose spread grows with amp[j]. The hash itself is evaluated by the device generator
(csrc/synth.cu, `srb_synth_*`) — this module only builds the two per-gene tables it consumes, so that any
other generator fed the same tables produces the same matrix bit for bit.

Detection rates are log-uniform (heavy-tailed, like real scRNA-seq): p_j = p_max * 2^(-t_j), t_j ~ U[0, T),
rescaled so that mean(p_j) equals `mean_density`, clipped to [1e-4, 0.6].
"""
from __future__ import annotations

import numpy as np

P_MAX = 0.6
P_MIN = 1e-4


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    m = np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def gene_tables(n_genes: int, seed: int = 0, mean_density: float = 0.05, log2_span: float = 12.0):
    """Returns (thr u32[n_genes], amp u32[n_genes])."""
    j = np.arange(n_genes, dtype=np.uint64)
    u = _mix32((j * np.uint64(0x9E3779B1) + np.uint64(seed & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF))
    a = _mix32(u ^ np.uint64(0x2545F491))
    t = (u.astype(np.float64) / 2.0**32) * log2_span
    p = P_MAX * np.exp2(-t)
    for _ in range(30):  # rescale to the requested mean under the clip
        p = np.clip(p * (mean_density / p.mean()), P_MIN, P_MAX)
    thr = np.minimum(np.rint(p * 2.0**32), 2.0**32 - 1).astype(np.uint32)
    amp = (a & np.uint64(0xFF)).astype(np.uint32)
    return thr, amp
