"""Host-side logic of the cell-row-sharded job (one process per GPU).

The data path has exactly two exchanges, both issued by the library on its own NCCL communicator:
  1. sum-allreduce of the per-gene integer moment limbs (6 x n_genes u64) — preceded by a 5-double MAX-allreduce that
     makes the fixed-point scale 2^F and the code-path decision identical on every rank;
  2. sum-allreduce of the d x d Gram matrix.
torch.distributed is only the rendezvous that carries the 128-byte ncclUniqueId.

The limb arithmetic is restated here in NumPy / Python integers (`fexp_from_bound`, `limbs_from_values`,
`finalize_limbs`) so the property the design rests on — integer sums are associative, hence the per-gene moments are
bit-identical for any sharding — can be tested on CPU with the gloo backend (tests/test_parallel_cpu.py)."""
from __future__ import annotations

import math

import numpy as np


def shard_rows(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Rank r owns rows [r*n/G, (r+1)*n/G); the remainder goes to the first ranks."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def comm_init_from_torch(ctx, group=None):
    """Broadcast rank 0's ncclUniqueId over an initialised torch.distributed group and join the library's communicator."""
    import torch.distributed as dist
    from . import _ffi
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    obj = [_ffi.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    ctx.comm_init(obj[0], rank, world)
    return rank, world


def fexp_from_bound(bound: float, do_log1p: bool) -> int:
    """Mirror of fexp_kernel (csrc/minor_moments.cu): F with rint(bound' * 2^F) < 2^28."""
    b = math.log1p(bound) if do_log1p else bound
    b *= 1.0 + 1e-6
    e = 0
    if b > 0.0 and math.isfinite(b):
        e = math.frexp(b)[1]  # b < 2^e  (ilogb(b) + 1)
    return max(-900, min(100, 28 - e))


def limbs_from_values(cols: np.ndarray, vals: np.ndarray, n_genes: int, F: int) -> np.ndarray:
    """What the fused kernel accumulates for one shard: int64[6, n_genes] = cnt, sumA, sumB, sqA, sqB, sqC."""
    q = np.rint(np.asarray(vals, dtype=np.float64) * math.ldexp(1.0, F)).astype(np.uint64)
    cols = np.asarray(cols, dtype=np.int64)
    acc = np.zeros((6, n_genes), dtype=np.int64)
    np.add.at(acc[0], cols, 1)
    np.add.at(acc[1], cols, (q & np.uint64(0xFFFFFFFF)).astype(np.int64))
    np.add.at(acc[2], cols, (q >> np.uint64(32)).astype(np.int64))
    q2_lo = (q * q) & np.uint64(0xFFFFFFFFFFFFFFFF)           # q < 2^28 => q*q < 2^56 fits
    np.add.at(acc[3], cols, (q2_lo & np.uint64(0xFFFFFFFF)).astype(np.int64))
    np.add.at(acc[4], cols, (q2_lo >> np.uint64(32)).astype(np.int64))
    return acc


def finalize_limbs(acc: np.ndarray, F: int):
    """Mirror of finalize_exact_kernel + minor_variance_exact_kernel: (count u64, sum f64, sumsq f64, variance f64)."""
    n = acc.shape[1]
    cnt = acc[0].astype(np.uint64)
    s_out, q_out, v_out = np.zeros(n), np.zeros(n), np.zeros(n)
    for j in range(n):
        c = int(acc[0, j])
        S = int(acc[1, j]) + (int(acc[2, j]) << 32)
        Q = int(acc[3, j]) + (int(acc[4, j]) << 32) + (int(acc[5, j]) << 64)
        s_out[j] = math.ldexp(float(S), -F)
        q_out[j] = math.ldexp(float(Q), -2 * F)
        if c > 0:
            N = c * Q - S * S
            v_out[j] = math.ldexp(float(max(N, 0)) / (float(c) * float(c)), -2 * F)
    return cnt, s_out, q_out, v_out


def tri_index(i: int, j: int, d: int) -> int:
    """Position of G[i][j] (i <= j) in the packed upper triangle the ranks allreduce (csrc/pca.cu: tri_pack_kernel)."""
    assert 0 <= i <= j < d
    return i * d - i * (i - 1) // 2 + (j - i)


def tri_pack(G: np.ndarray) -> np.ndarray:
    d = G.shape[0]
    return G[np.triu_indices(d)]


def tri_unpack(T: np.ndarray, d: int) -> np.ndarray:
    G = np.zeros((d, d), dtype=T.dtype)
    iu = np.triu_indices(d)
    G[iu] = T
    G.T[iu] = T
    return G
