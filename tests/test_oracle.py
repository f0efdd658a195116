"""Pins the CPU oracle: hand-derived known answers + an independent NumPy/SciPy statement.

The reference has no golden vectors for this path (parity unpinned); these are ours.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle import pca_oracle as P
from tests._util import f, random_csr, sign_align


def kat_matrix(kat, dtype=np.float64, fmt="csr"):
    m = sp.csr_matrix((np.array(kat["data"], dtype=dtype), kat["indices"], kat["indptr"]), shape=kat["shape"])
    if fmt == "csc":
        m = m.tocsc()
    return O.Compressed.from_scipy(m)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.uint8])
def test_kat_stats_csr(kat, dtype):
    m = kat_matrix(kat, dtype)
    for d, key in ((O.ROW, "row"), (O.COLUMN, "column")):
        assert O.number(m, d).tolist() == kat["number"][key]
        np.testing.assert_array_equal(O.sum_(m, d), f(kat["sum"][key]))
        np.testing.assert_allclose(O.variance(m, d), f(kat["variance"][key]), rtol=0, atol=2e-15, equal_nan=True)
        np.testing.assert_allclose(O.std_dev(m, d), f(kat["std_dev"][key]), rtol=0, atol=2e-15, equal_nan=True)
        mn, mx = O.min_max(m, d)
        np.testing.assert_array_equal(mn, f(kat["min"][key]))
        np.testing.assert_array_equal(mx, f(kat["max"][key]))


def test_kat_csc_mirror(kat):
    """csc.rs is the mirror of csr.rs: the two-pass/one-pass variance forms swap directions."""
    m = kat_matrix(kat, fmt="csc")
    assert O.number(m, O.ROW).tolist() == kat["number"]["row"]
    assert O.number(m, O.COLUMN).tolist() == kat["number"]["column"]
    np.testing.assert_array_equal(O.sum_(m, O.ROW), f(kat["sum"]["row"]))
    np.testing.assert_array_equal(O.sum_(m, O.COLUMN), f(kat["sum"]["column"]))
    # CSC Column = major two-pass: empty gene 4 -> NaN (csc.rs:161-171); CSC Row = minor one-pass: empty -> 0
    vc = O.variance(m, O.COLUMN)
    assert np.isnan(vc[4]) and vc[1] == 0.0
    np.testing.assert_allclose(vc[[0, 2, 3]], [14 / 9, 0.0, 2.25], atol=2e-15)
    vr = O.variance(m, O.ROW)
    assert vr[2] == 0.0
    np.testing.assert_allclose(vr[[0, 1, 3]], [2 / 3, 0.25, 32 / 9], atol=2e-15)


def test_kat_normalize_log1p_hvg(kat):
    m = kat_matrix(kat)
    nm = O.normalize_total(m, 10.0, O.ROW)
    assert nm.values.dtype == np.float64
    np.testing.assert_allclose(nm.values, kat["normalize_total_row_target10"]["values"], rtol=0, atol=1e-15)
    lm = O.log1p(nm)
    np.testing.assert_allclose(lm.values, kat["log1p_after_normalize"], rtol=0, atol=4e-16)
    gv = O.variance(lm, O.COLUMN)
    np.testing.assert_allclose(gv, kat["gene_variance_after_log1p"], rtol=0, atol=1e-15)
    assert O.select_hvg(gv, 3).tolist() == kat["hvg_top3"]
    assert O.select_hvg(gv, 5).tolist() == kat["hvg_full_order"]
    dense = O.densify_selected(lm, np.arange(4), np.array(kat["hvg_top3"]))
    assert dense.shape == (4, 3)
    assert np.all(dense[2] == 0)
    np.testing.assert_array_equal(dense[:, 1], [lm.values[1], 0.0, 0.0, lm.values[6]])


def test_f32_input_paths(kat):
    m = kat_matrix(kat, np.float32)
    nm = O.normalize_total(m, 10.0, O.ROW)
    assert nm.values.dtype == np.float64  # scale/mod.rs:82
    lm32 = O.log1p(m)
    assert lm32.values.dtype == np.float32  # transform/mod.rs:43-46
    np.testing.assert_array_equal(lm32.values, np.log1p(np.array(kat["data"], dtype=np.float32)))


def test_select_hvg_ties_and_nan():
    v = np.array([1.0, 3.0, 3.0, 0.0, 3.0, 2.0])
    assert O.select_hvg(v, 4).tolist() == [1, 2, 4, 5]
    assert O.select_var_threshold(v, 1.0).tolist() == [1, 2, 4, 5]
    with pytest.raises(ValueError):
        O.select_hvg(np.array([1.0, np.nan]), 1)


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_against_numpy_statement(fmt, dtype):
    """Independent NumPy/SciPy statement of the same formulas on a random matrix with empty lines."""
    rng = np.random.default_rng(7)
    a = random_csr(rng, 300, 70, 0.1, dtype=dtype, empty_rows=(3, 17), empty_cols=(5,))
    m = O.Compressed.from_scipy(a if fmt == "csr" else a.tocsc())
    d = a.toarray().astype(np.float64)
    nz = (a != 0).toarray()  # generator stores no explicit zeros
    for direction, axis in ((O.ROW, 1), (O.COLUMN, 0)):
        cnt = nz.sum(axis=axis)
        s = d.sum(axis=axis)
        assert O.number(m, direction).tolist() == cnt.tolist()
        np.testing.assert_allclose(O.sum_(m, direction), s, rtol=1e-13)
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = s / cnt
            var = (np.where(nz, (d - np.expand_dims(mean, axis)) ** 2, 0)).sum(axis=axis) / cnt
        got = O.variance(m, direction)
        major = m.along_major(direction)
        empty = cnt == 0
        if major:
            assert np.all(np.isnan(got[empty]))
        else:
            assert np.all(got[empty] == 0.0)
        np.testing.assert_allclose(got[~empty], var[~empty], rtol=1e-9, atol=1e-12)
        mn, mx = O.min_max(m, direction)
        dm = np.where(nz, d, np.inf).min(axis=axis)
        dM = np.where(nz, d, -np.inf).max(axis=axis)
        np.testing.assert_array_equal(mn, dm)
        np.testing.assert_array_equal(mx, dM)
        # normalise: every non-empty line sums to target (reference test, processing/mod.rs:420-481, 1e-6 abs)
        nm = O.normalize_total(m, 1e4, direction)
        sums = O.sum_(nm, direction)
        np.testing.assert_allclose(sums[~empty], 1e4, rtol=0, atol=1e-6)
        assert np.all(sums[empty] == 0)


def test_chunked_matches_whole_in_minor_direction_and_documents_major_defect():
    rng = np.random.default_rng(11)
    a = random_csr(rng, 100, 30, 0.2, dtype=np.float32)
    m = O.Compressed.from_scipy(a)
    cnt = np.zeros(30, dtype=np.uint32)
    sm = np.zeros(30)
    cnt_row_fixed = np.zeros(100, dtype=np.uint32)
    cnt_row_faithful = np.zeros(100, dtype=np.uint32)
    sum_row_fixed = np.zeros(100)
    for s in range(0, 100, 32):
        ch = O.Compressed.from_scipy(a[s:s + 32])
        O.number_chunk(ch, O.COLUMN, cnt)
        O.sum_chunk(ch, O.COLUMN, sm)
        O.number_chunk(ch, O.ROW, cnt_row_fixed, major_offset=s)
        O.number_chunk(ch, O.ROW, cnt_row_faithful)
        O.sum_chunk(ch, O.ROW, sum_row_fixed, major_offset=s)
    assert cnt.tolist() == O.number(m, O.COLUMN).tolist()
    np.testing.assert_allclose(sm, O.sum_(m, O.COLUMN), rtol=1e-13)
    assert cnt_row_fixed.tolist() == O.number(m, O.ROW).tolist()
    np.testing.assert_array_equal(sum_row_fixed, O.sum_(m, O.ROW))
    # reference defect (shared/statistics/mod.rs:24, csr.rs:56-61): chunk-local index => rows >= 32 stay 0
    assert cnt_row_faithful[32:].sum() == 0 and cnt_row_faithful[:32].sum() == a.nnz


def test_pca_oracle_against_sklearn():
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(3)
    X = rng.normal(size=(200, 12)) @ rng.normal(size=(12, 12)) + rng.normal(size=12)
    res = P.pca_fit_transform(X, 4, center=True, scale=True)
    Z = StandardScaler().fit_transform(X)
    sk = PCA(n_components=4, svd_solver="full").fit(Z)
    np.testing.assert_allclose(res["explained_variance_ratio"], sk.explained_variance_ratio_, rtol=1e-10)
    comps = sign_align(res["components"], sk.components_.T)
    np.testing.assert_allclose(comps, sk.components_.T, atol=1e-9)
    np.testing.assert_allclose(sign_align(res["scores"], sk.transform(Z)), sk.transform(Z), atol=1e-8)


def test_synth_generator_is_deterministic_and_sharded():
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(400, seed=5, mean_density=0.05)
    whole = O.synth_csr(0x5EED0001, 300, 400, thr, amp)
    a = O.synth_csr(0x5EED0001, 100, 400, thr, amp, row0=0)
    b = O.synth_csr(0x5EED0001, 200, 400, thr, amp, row0=100)
    assert whole.nnz == a.nnz + b.nnz
    np.testing.assert_array_equal(whole.indices, np.concatenate([a.indices, b.indices]))
    np.testing.assert_array_equal(whole.values, np.concatenate([a.values, b.values]))
    dens = whole.nnz / (300 * 400)
    assert 0.02 < dens < 0.09
    # indices sorted and unique within each row (nalgebra-sparse invariant)
    for i in range(300):
        r = whole.indices[int(whole.offsets[i]):int(whole.offsets[i + 1])]
        assert np.all(np.diff(r.astype(np.int64)) > 0)
