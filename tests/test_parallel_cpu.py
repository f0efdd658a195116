"""CPU tests (gloo, world_size 2) of the host-side logic of the row-sharded job, plus the C-ABI export check."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    from singlerust_b200 import _ffi
    from singlerust_b200 import build as b
    b.build()
    lib = ctypes.CDLL(_ffi.lib_path())
    hdr = open(os.path.join(ROOT, "include", "srb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    _ffi.lib()  # prototypes resolve


def test_bindings_declare_every_header_symbol():
    """The Rust extern block (bindings/single_rust_b200.rs, mirrored in INTEGRATION.md) and the ctypes table bind the
    same symbol set as include/srb200.h — none missing, none invented."""
    from singlerust_b200 import _ffi
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "srb200.h")).read(), flags=re.S)
    names = set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "bindings", "single_rust_b200.rs")).read()
    rs_names = set(re.findall(r"pub fn (srb_[a-z0-9_]+)\s*\(", rs))
    assert rs_names == names, (sorted(names - rs_names), sorted(rs_names - names))
    lib = _ffi.lib()
    untyped = [n for n in names if getattr(lib, n).argtypes is None and n not in
               ("srb_version", "srb_last_error_message", "srb_kernel_launch_count")]
    assert not untyped, untyped


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a device the product path raises (this container has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from singlerust_b200 import _ffi
    with pytest.raises(_ffi.SrbError) as e:
        _ffi.Context(0)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "singlerust_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_shard_rows_partition():
    from singlerust_b200.parallel import shard_rows
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_enums_match_reference_discriminants():
    from singlerust_b200.shared import ComputationMode, Direction, FeatureSelection, FlexValue
    assert int(Direction.Row) == 0 and int(Direction.Column) == 1 and Direction.Row.is_row()
    assert ComputationMode.Whole().is_whole and ComputationMode.Chunked(1000).chunk == 1000
    assert FeatureSelection.HighlyVariable(25).value == 25
    assert FlexValue.Absolute(200).is_absolute() and FlexValue.None_().is_none()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from singlerust_b200 import synth
    from singlerust_b200.parallel import fexp_from_bound, finalize_limbs, limbs_from_values, shard_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, m = 600, 500
    thr, amp = synth.gene_tables(m, seed=3, mean_density=0.08)
    a, b = shard_rows(n, world, rank)
    shard = O.synth_csr(0x5EED0001, b - a, m, thr, amp, row0=a, skew=True)
    ln = O.log1p(O.normalize_total(shard, 1e4, O.ROW))           # per-cell ops are local to the shard
    vals32 = ln.values.astype(np.float32)                        # COMPACT storage
    # exchange 0: global bound -> identical F on every rank
    bound = torch.tensor([float(np.max(O.normalize_total(shard, 1e4, O.ROW).values, initial=0.0))], dtype=torch.float64)
    dist.all_reduce(bound, op=dist.ReduceOp.MAX)
    F = fexp_from_bound(float(bound.item()), True)
    # exchange 1: integer limbs
    acc = torch.from_numpy(limbs_from_values(ln.indices, vals32, m, F))
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    cnt, s, sq, var = finalize_limbs(acc.numpy(), F)
    if rank == 0:
        q.put((F, cnt, s, sq, var))
    dist.destroy_process_group()


def _run(world, port):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_sharded_gene_moments_are_bit_identical_and_match_oracle():
    """world_size 2 (gloo) == world_size 1, bit for bit; both within 1e-5 of the oracle on the whole matrix."""
    from oracle import oracle as O
    from singlerust_b200 import synth
    F1, cnt1, s1, sq1, var1 = _run(1, 29611)
    F2, cnt2, s2, sq2, var2 = _run(2, 29613)
    assert F1 == F2
    np.testing.assert_array_equal(cnt1, cnt2)
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(sq1, sq2)
    np.testing.assert_array_equal(var1, var2)
    thr, amp = synth.gene_tables(500, seed=3, mean_density=0.08)
    whole = O.synth_csr(0x5EED0001, 600, 500, thr, amp, skew=True)
    ln = O.log1p(O.normalize_total(whole, 1e4, O.ROW))
    np.testing.assert_array_equal(cnt1, O.number(ln, O.COLUMN))
    np.testing.assert_allclose(s1, O.sum_(ln, O.COLUMN), rtol=1e-6)
    want = O.variance(ln, O.COLUMN)
    ex2 = sq1 / np.maximum(cnt1, 1)
    # backward-error bound of f32 storage: |dvar| <= 1e-5 var + 4e-7 E[x^2]
    assert np.all(np.abs(var1 - want) <= 1e-5 * want + 4e-7 * ex2)


def _worker_gram(rank, world, port, q):
    """The PCA exchange of a row-sharded job restated on the CPU: global per-gene moments (allreduce 1), each rank's
    Gram matrix of its standardised rows, packed upper triangle summed over ranks (allreduce 2), replicated eigensolve."""
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from singlerust_b200 import synth
    from singlerust_b200.parallel import shard_rows, tri_index, tri_pack, tri_unpack
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, m, d, k = 500, 120, 40, 4
    thr, amp = synth.gene_tables(m, seed=5, mean_density=0.3)
    a, b = shard_rows(n, world, rank)
    shard = O.synth_csr(0x5EED0004, b - a, m, thr, amp, row0=a, skew=True)
    ln = O.log1p(O.normalize_total(shard, 1e4, O.ROW))
    sel = np.arange(d, dtype=np.uint64)
    X = O.densify_selected(ln, np.arange(b - a, dtype=np.uint64), sel)
    mom = torch.from_numpy(np.stack([X.sum(axis=0), (X * X).sum(axis=0)]))
    dist.all_reduce(mom, op=dist.ReduceOp.SUM)
    mean = mom[0].numpy() / n
    std = np.sqrt(mom[1].numpy() / n - mean * mean)
    Z = (X - mean) / std
    T = torch.from_numpy(tri_pack(Z.T @ Z))
    assert T.numel() == d * (d + 1) // 2 and tri_index(3, 7, d) == int(np.flatnonzero((np.triu_indices(d)[0] == 3) & (np.triu_indices(d)[1] == 7))[0])
    dist.all_reduce(T, op=dist.ReduceOp.SUM)
    G = tri_unpack(T.numpy(), d)
    w, V = np.linalg.eigh(G)
    if rank == 0:
        q.put((G, w[::-1][:k] / np.trace(G), V[:, ::-1][:, :k]))
    dist.destroy_process_group()


def test_sharded_gram_triangle_exchange_matches_oracle_pca():
    """world_size 2 (gloo): the packed-triangle Gram exchange reproduces the whole-matrix PCA of the oracle."""
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from oracle import pca_oracle as P
    from singlerust_b200 import synth
    from tests._util import sign_align
    out = {}
    for world, port in ((1, 29631), (2, 29633)):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker_gram, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        out[world] = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    np.testing.assert_allclose(out[2][0], out[1][0], rtol=1e-12, atol=1e-9)
    thr, amp = synth.gene_tables(120, seed=5, mean_density=0.3)
    whole = O.log1p(O.normalize_total(O.synth_csr(0x5EED0004, 500, 120, thr, amp, skew=True), 1e4, O.ROW))
    want = P.pca_pipeline(whole, 40, 4, selection=np.arange(40, dtype=np.uint64))
    np.testing.assert_allclose(out[2][1], want["explained_variance_ratio"], rtol=1e-9)
    np.testing.assert_allclose(sign_align(out[2][2], want["components"]), want["components"], atol=1e-8)


def test_filter_mask_host_logic_matches_oracle_all_nine_arms():
    """create_filter_mask / calculate_percentiles (processing/mod.rs:32-83, 148-174): host logic vs the arm-by-arm oracle."""
    from oracle import filter_oracle as FO
    from singlerust_b200.memory.processing import calculate_percentiles, create_filter_mask
    from singlerust_b200.shared import FlexValue
    rng = np.random.default_rng(5)
    counts = rng.integers(0, 60, size=500).astype(np.uint32)
    sums = rng.uniform(0, 1000, size=500)
    np.testing.assert_allclose(FO.linear_quantile(sums, 0.37), np.quantile(sums, 0.37), rtol=1e-15)

    def fv(t):
        return {"Absolute": FlexValue.Absolute, "Relative": FlexValue.Relative}[t[0]](t[1]) if t[0] != "None" else FlexValue.None_()

    for lo in (("Absolute", 10), ("Relative", 0.25), ("None", None)):
        for up in (("Absolute", 45), ("Relative", 0.8), ("None", None)):
            lp, upc = calculate_percentiles(sums, fv(lo), fv(up))
            got = create_filter_mask(500, counts, sums, fv(lo), fv(up), lp, upc)
            np.testing.assert_array_equal(got, FO.mask(counts, sums, lo, up))
    with pytest.raises(ValueError):
        calculate_percentiles(np.array([1.0, np.nan]), FlexValue.Relative(0.5), FlexValue.None_())
