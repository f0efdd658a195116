"""The reference's own tests, restated against the mirrored API (same names / arguments), with the assertions the
reference tests lack added against the oracle:
  - src/memory/processing/mod.rs:420-481  test_normalize_total (line sums == target_sum)
  - tests/test_basic_stats.rs:21-85       compute_number Row/Column in memory and backed Chunked(1000)
  - tests/test_basic_load.rs:109-171      pca_inplace(HighlyVariable(25), 5 components)
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle import pca_oracle as P
from tests._util import random_csr, sign_align

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def env():
    from singlerust_b200 import _ffi
    from singlerust_b200 import backed, memory
    from singlerust_b200.anndata import BackedAnnData, IMAnnData
    from singlerust_b200.shared import ComputationMode, Direction, FeatureSelection
    ctx = _ffi.Context(0, value_mode=_ffi.VALUES_FAITHFUL)
    yield dict(ffi=_ffi, ctx=ctx, memory=memory, backed=backed, IMAnnData=IMAnnData, BackedAnnData=BackedAnnData,
               Direction=Direction, ComputationMode=ComputationMode, FeatureSelection=FeatureSelection)
    ctx.close()


def create_large_test_data(rows, cols, sparsity, seed):
    """processing/mod.rs:343-376: random COO with duplicates summed, values U(0,50), nnz = rows*cols/sparsity."""
    rng = np.random.default_rng(seed)
    nnz = int(rows * cols / sparsity)
    coo = sp.coo_matrix((rng.uniform(0.0, 50.0, nnz), (rng.integers(0, rows, nnz), rng.integers(0, cols, nnz))), shape=(rows, cols))
    a = coo.tocsr()  # duplicates are summed like CsrMatrix::from(&coo)
    a.sort_indices()
    return a


def test_normalize_total(env):
    """check_row_sums / check_column_sums with the reference's tolerance (1e-6 absolute on 1e4), skipping the empty
    lines the reference test forgets about."""
    D, mem = env["Direction"], env["memory"]
    a = create_large_test_data(1000, 100, 10.0, seed=1)
    adata = env["IMAnnData"].from_scipy(env["ctx"], a)
    target_sum = 1e4
    normalized = mem.processing.normalize_total(adata, target_sum, D.Row)
    _, _, v = normalized.x().download()
    rows = np.add.reduceat(v, a.indptr[:-1][np.diff(a.indptr) > 0])
    assert np.all(np.abs(rows - target_sum) < 1e-6)
    # the non-inplace form leaves the original untouched
    _, _, v0 = adata.x().download()
    np.testing.assert_array_equal(v0, a.data)
    mem.processing.normalize_total_inplace(adata, target_sum, D.Column)
    off, idx, v = adata.x().download()
    cols = np.bincount(idx.astype(np.int64), weights=v, minlength=100)
    nonempty = np.bincount(idx.astype(np.int64), minlength=100) > 0
    assert np.all(np.abs(cols[nonempty] - target_sum) < 1e-6)


def test_n_genes_in_memory_and_backed_chunked(env):
    """tests/test_basic_stats.rs: compute_number Row/Column, in memory and backed Chunked(1000) — with assertions."""
    D, CM = env["Direction"], env["ComputationMode"]
    a = random_csr(np.random.default_rng(3), 4500, 300, 0.05, dtype=np.float32)
    o = O.Compressed.from_scipy(a)
    adata = env["IMAnnData"].from_scipy(env["ctx"], a)
    res = env["memory"].statistics.compute_number(adata, D.Row)
    assert len(res) == 4500 and res.dtype == np.uint32
    np.testing.assert_array_equal(res, O.number(o, O.ROW))
    np.testing.assert_array_equal(env["memory"].statistics.compute_number(adata, D.Column), O.number(o, O.COLUMN))
    b = env["BackedAnnData"](a)
    for d, od in ((D.Row, O.ROW), (D.Column, O.COLUMN)):
        np.testing.assert_array_equal(env["backed"].statistics.compute_number(env["ctx"], b, d, CM.Chunked(1000)), O.number(o, od))
        np.testing.assert_array_equal(env["backed"].statistics.compute_number(env["ctx"], b, d, CM.Whole()), O.number(o, od))
        np.testing.assert_allclose(env["backed"].statistics.compute_sum(env["ctx"], b, d, CM.Chunked(1000)), O.sum_(o, od), rtol=1e-12)
    # CSC-backed data streams column chunks
    bc = env["BackedAnnData"](a.tocsc())
    np.testing.assert_array_equal(env["backed"].statistics.compute_number(env["ctx"], bc, D.Row, CM.Chunked(64)), O.number(o, O.ROW))
    np.testing.assert_allclose(env["backed"].statistics.compute_sum(env["ctx"], bc, D.Column, CM.Chunked(64)), O.sum_(o, O.COLUMN), rtol=1e-12)


def test_qc_vars_inplace_column_names(env):
    a = random_csr(np.random.default_rng(4), 300, 80, 0.1, dtype=np.float32)
    adata = env["IMAnnData"].from_scipy(env["ctx"], a)
    env["memory"].statistics.qc_vars_inplace(adata)
    assert set(adata.obs) == {"num_genes_per_cell", "sum_expr_per_cell", "var_expr_per_cell", "std_dev_per_cell"}
    assert set(adata.var) == {"num_cells_per_gene", "sum_expr_per_gene", "var_expr_per_gene", "std_dev_per_gene"}
    o = O.Compressed.from_scipy(a)
    np.testing.assert_array_equal(adata.var["num_cells_per_gene"], O.number(o, O.COLUMN))
    np.testing.assert_allclose(adata.obs["std_dev_per_cell"], O.std_dev(o, O.ROW), rtol=1e-9, equal_nan=True)


def test_pca_inplace_hvg25(env):
    """tests/test_basic_load.rs:143-161: pca_inplace(Some(5), center, scale, threads, HighlyVariable(25), svd)."""
    from tests.test_gpu_parity import clustered_counts
    FS, D, mem = env["FeatureSelection"], env["Direction"], env["memory"]
    a = clustered_counts(np.random.default_rng(8), 2500, 200)
    adata = env["IMAnnData"].from_scipy(env["ctx"], a)
    mem.processing.normalize_total_inplace(adata, 1e4, D.Row)
    mem.processing.log1p_transform_inplace(adata)
    mem.processing.pca_inplace(adata, 5, True, True, 32, FS.HighlyVariable(25), svd_mode="LapackSVD")
    assert adata.obsm["X_pca"].shape == (2500, 5)
    ol = O.log1p(O.normalize_total(O.Compressed.from_scipy(a), 1e4, O.ROW))
    want = P.pca_pipeline(ol, 25, 5)
    got = sign_align(adata.obsm["X_pca"], want["scores"])
    for j in range(3):
        rms = np.linalg.norm(want["scores"][:, j]) / np.sqrt(2500)
        assert np.max(np.abs(got[:, j] - want["scores"][:, j])) <= 1e-4 * rms
    np.testing.assert_allclose(adata.explained_variance_ratio, want["explained_variance_ratio"], rtol=1e-5)
    # defaults: n_components=None -> 2 (dim_red/mod.rs:52)
    mem.processing.pca_inplace(adata, None, None, None, None, FS.VarianceThreshold(0.05))
    assert adata.obsm["X_pca"].shape[1] == 2


def test_two_gpu_row_sharding_matches_single_gpu():
    """Runs tests/multigpu_check.py under torchrun on 2 GPUs when the box has them."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29655", os.path.join(ROOT, "tests", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU OK" in r.stdout


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_backed_chunked_pipeline_equals_in_memory(env, fmt):
    """BASELINE config 5 in miniature: chunk-streamed (resident) X gives bit-identical statistics and the same PCA as
    the whole-matrix upload; chunk boundaries must not matter."""
    from tests.test_gpu_parity import clustered_counts
    D, CM, FS, mem, backed = env["Direction"], env["ComputationMode"], env["FeatureSelection"], env["memory"], env["backed"]
    a = clustered_counts(np.random.default_rng(12), 1700, 240)
    src = a if fmt == "csr" else a.tocsc()
    whole = env["IMAnnData"].from_scipy(env["ctx"], src)
    b = env["BackedAnnData"](src)
    for chunk in (1, 333, 5000):
        dev = backed.processing.load_resident(env["ctx"], b, CM.Chunked(chunk))
        off, idx, val = dev.x().download()
        off0, idx0, val0 = whole.x().download()
        np.testing.assert_array_equal(off, off0)
        np.testing.assert_array_equal(idx, idx0)
        np.testing.assert_array_equal(val, val0)
    res = backed.processing.normalize_hvg_pca(env["ctx"], b, CM.Chunked(400), 1e4, 30, 5)
    ref = whole.deep_clone()
    mem.processing.normalize_total_inplace(ref, 1e4, D.Row)
    mem.processing.log1p_transform_inplace(ref)
    mem.processing.pca_inplace(ref, 5, True, True, None, FS.HighlyVariable(30))
    np.testing.assert_allclose(res.explained_variance_ratio, ref.explained_variance_ratio, rtol=1e-9)
    np.testing.assert_allclose(sign_align(res.obsm["X_pca"], ref.obsm["X_pca"]), ref.obsm["X_pca"], atol=1e-6)


LIMS = [("Absolute", 12), ("Relative", 0.2), ("None", None)]
UPS = [("Absolute", 30), ("Relative", 0.9), ("None", None)]


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_filter_cells_and_genes_all_flexvalue_arms(env, fmt):
    """filter_cells / filter_genes (processing/mod.rs:86-146, 245-299) for all nine FlexValue combinations: the mask
    and the compacted matrix are bit-identical to the oracle's; plus the reference's own assertion (fewer lines)."""
    from oracle import filter_oracle as FO
    from singlerust_b200.shared import FlexValue
    mem = env["memory"]
    a = random_csr(np.random.default_rng(17), 400, 90, 0.25, dtype=np.float32, empty_rows=(3,), empty_cols=(8,))
    src = a if fmt == "csr" else a.tocsc()

    def fv(t):
        return {"Absolute": FlexValue.Absolute, "Relative": FlexValue.Relative}[t[0]](t[1]) if t[0] != "None" else FlexValue.None_()

    for lo in LIMS:
        for up in UPS:
            for axis, fn in ((0, mem.processing.filter_cells), (1, mem.processing.filter_genes)):
                lo_a, up_a = lo, up
                if axis == 1 and lo[0] == "Absolute":
                    lo_a, up_a = ("Absolute", 80), (up if up[0] != "Absolute" else ("Absolute", 120))
                adata = env["IMAnnData"].from_scipy(env["ctx"], src)
                adata.obs["tag"] = np.arange(400)
                adata.var["tag"] = np.arange(90)
                got = fn(adata, fv(lo_a), fv(up_a))
                want, m = FO.filter_matrix(a, lo_a, up_a, axis)
                off, idx, val = got.x().download()
                w = want if fmt == "csr" else want.tocsc()
                w.sort_indices()
                assert got.x().shape == want.shape
                np.testing.assert_array_equal(off, w.indptr)
                np.testing.assert_array_equal(idx, w.indices)
                np.testing.assert_array_equal(val, w.data)
                tags = got.obs["tag"] if axis == 0 else got.var["tag"]
                np.testing.assert_array_equal(tags, np.nonzero(m)[0])
                assert adata.x().shape == (400, 90)          # non-inplace leaves the input untouched
    # the reference's own tests (processing/mod.rs:385-417): filtered count < original, Absolute and Relative limits
    adata = env["IMAnnData"].from_scipy(env["ctx"], src)
    n0 = adata.n_obs
    mem.processing.filter_cells_inplace(adata, FlexValue.Absolute(20), FlexValue.None_())
    assert adata.n_obs < n0
    g0 = adata.n_vars
    mem.processing.filter_genes_inplace(adata, FlexValue.Relative(0.1), FlexValue.Relative(0.9))
    assert adata.n_vars < g0
    # statistics keep working on the compacted matrix
    off, idx, val = adata.x().download()
    assert np.array_equal(adata.x().number(0), np.diff(off.astype(np.int64)) if fmt == "csr" else np.bincount(idx.astype(np.int64), minlength=adata.n_obs))
