"""Edge cases of the path the reference's semantics make observable: explicitly stored zeros,
duplicate / out-of-order feature selections, degenerate shapes, warp-boundary line lengths."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle import pca_oracle as P
from tests._util import random_csr, sign_align

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def ffi():
    from singlerust_b200 import _ffi
    return _ffi


@pytest.fixture(scope="module")
def ctx(ffi):
    c = ffi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_faithful(ffi):
    c = ffi.Context(0, value_mode=ffi.VALUES_FAITHFUL)
    yield c
    c.close()


def close(got, want, rtol=RTOL, atol=0.0):
    got, want = np.asarray(got), np.asarray(want)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf])
    ok = ~(np.isnan(want) | inf)
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=atol)


def close_compact_variance(got, ol, direction):
    """COMPACT (f32) storage after normalise / log1p: the documented backward-error bound of DESIGN.md §4,
    |dvar| <= 1e-5 var + 4e-7 E[x^2] per line (a 2e-7 relative perturbation of every stored value)."""
    want = O.variance(ol, direction)
    sq = O.sum_(O.Compressed(ol.fmt, ol.nrows, ol.ncols, ol.offsets, ol.indices, ol.values ** 2), direction)
    cnt = np.maximum(O.number(ol, direction), 1)
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    ok = ~np.isnan(want)
    assert np.all(np.abs(got[ok] - want[ok]) <= 1e-5 * np.abs(want[ok]) + 4e-7 * sq[ok] / cnt[ok] + 1e-12)


def csr_with_explicit_zeros(rng, n, m, density, zero_frac, dtype=np.float32):
    """Canonical CSR whose stored entries include zeros (anndata keeps them; nalgebra-sparse does not prune)."""
    a = random_csr(rng, n, m, density, dtype=dtype, empty_rows=(2,), empty_cols=(3,))
    data = a.data.copy()
    data[rng.random(data.size) < zero_frac] = 0
    # one line whose stored entries are ALL zero: sum 0 -> scale 0, min = max = 0, variance 0 (not NaN: count > 0)
    r = 5
    data[a.indptr[r]:a.indptr[r + 1]] = 0
    out = sp.csr_matrix((data, a.indices.copy(), a.indptr.copy()), shape=a.shape)
    assert out.nnz == a.nnz  # scipy keeps explicit zeros until eliminate_zeros()
    return out


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint8])
def test_explicit_zeros_count_as_stored_entries(ffi, ctx, fmt, dtype):
    """number counts stored entries incl. zeros (csr.rs:21-36); the nonzero-only variance population is the STORED
    population (csr.rs:161-167,179-183); min/max see the zeros (csr.rs:200-220)."""
    rng = np.random.default_rng(31)
    a = csr_with_explicit_zeros(rng, 600, 150, 0.15, 0.2, dtype)
    if fmt == "csc":
        nnz = a.nnz
        a = sp.csc_matrix(a)
        a.sort_indices()
        assert a.nnz == nnz
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    o = O.Compressed.from_scipy(a)
    for d in (ffi.ROW, ffi.COLUMN):
        np.testing.assert_array_equal(m.number(d), O.number(o, d))
        np.testing.assert_array_equal(m.sum(d), O.sum_(o, d))
        close(m.variance(d), O.variance(o, d), atol=1e-9)
        close(m.std_dev(d), O.std_dev(o, d), atol=1e-6)
        mn, mx = m.min_max(d)
        wmn, wmx = O.min_max(o, d)
        np.testing.assert_array_equal(mn, wmn)
        np.testing.assert_array_equal(mx, wmx)
    assert m.number(ffi.ROW).sum() == a.nnz and (a.data == 0).sum() > 0


@pytest.mark.parametrize("mode", ["compact", "faithful"])
def test_explicit_zeros_through_normalize_log1p_hvg(ffi, ctx, ctx_faithful, mode):
    c = ctx if mode == "compact" else ctx_faithful
    rng = np.random.default_rng(32)
    a = csr_with_explicit_zeros(rng, 900, 400, 0.1, 0.25)
    m = ffi.DeviceMatrix.from_scipy(c, a)
    o = O.Compressed.from_scipy(a)
    m.normalize_total_inplace(1e4, ffi.ROW)
    m.log1p_inplace()
    ol = O.log1p(O.normalize_total(o, 1e4, O.ROW))
    _, _, v = m.download()
    close(v, ol.values, rtol=6e-7 if mode == "compact" else 1e-13)
    assert np.all(v[a.data == 0] == 0)                       # log1p(0 * scale) = 0 exactly
    s = m.sum(ffi.ROW)
    assert s[2] == 0 and s[5] == 0                            # empty line and all-zero line: scale 0 (scale/mod.rs:9-15)
    gv, want = m.variance(ffi.COLUMN), O.variance(ol, O.COLUMN)
    if mode == "compact":
        close_compact_variance(gv, ol, O.COLUMN)
    else:
        close(gv, want, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(m.number(ffi.COLUMN), O.number(ol, O.COLUMN))
    if mode == "faithful":
        np.testing.assert_array_equal(m.select_hvg(50), O.select_hvg(want, 50))


def test_densify_selection_order_and_duplicates(ffi, ctx):
    """Output column j = j-th entry of the selection list; a duplicated gene fills only its LAST position
    (HashMap insert, shared/mod.rs:241-256); unselected genes vanish; order need not be ascending."""
    rng = np.random.default_rng(33)
    a = random_csr(rng, 300, 60, 0.3)
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    o = O.Compressed.from_scipy(a)
    for sel in ([7, 3, 59, 0], [5, 9, 5, 2], [4, 4, 4], list(range(59, -1, -1)), [11]):
        got = m.densify_selected(sel)
        want = O.densify_selected(o, np.arange(300), sel)
        np.testing.assert_array_equal(got, want)
    got = m.densify_selected([5, 9, 5, 2])
    assert np.all(got[:, 0] == 0) and np.array_equal(got[:, 2], a[:, 5].toarray().ravel())


def test_pca_on_a_permuted_selection_permutes_the_loadings(ffi, ctx_faithful):
    rng = np.random.default_rng(34)
    a = random_csr(rng, 2000, 120, 0.2, integer=False)
    m = ffi.DeviceMatrix.from_scipy(ctx_faithful, a)
    sel = np.arange(10, 90, dtype=np.uint64)
    perm = rng.permutation(sel.size)
    r1 = m.pca(sel, 6, gram_mode=ffi.GRAM_FP64)
    r2 = m.pca(sel[perm], 6, gram_mode=ffi.GRAM_FP64)
    np.testing.assert_allclose(r1["explained_variance_ratio"], r2["explained_variance_ratio"], rtol=1e-10)
    np.testing.assert_allclose(sign_align(r2["components"], r1["components"][perm]), r1["components"][perm], atol=1e-8)
    np.testing.assert_allclose(sign_align(r2["scores"], r1["scores"]), r1["scores"], rtol=1e-7, atol=1e-7)
    want = P.pca_fit_transform(O.densify_selected(O.Compressed.from_scipy(a), np.arange(2000), sel), 6, True, True)
    np.testing.assert_allclose(r1["explained_variance_ratio"], want["explained_variance_ratio"], rtol=RTOL)


def test_selection_degenerate_requests(ffi, ctx):
    rng = np.random.default_rng(35)
    a = random_csr(rng, 200, 30, 0.3, empty_cols=(0, 29))
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    gv = O.variance(O.Compressed.from_scipy(a), O.COLUMN)
    np.testing.assert_array_equal(m.select_hvg(1000), O.select_hvg(gv, 1000))      # n_top > n_vars: all, in rank order
    assert m.select_hvg(1000).size == 30 and m.select_hvg(0).size == 0
    assert set(m.select_hvg(1000)[-2:].tolist()) == {0, 29}                          # empty genes (variance 0) rank last, by index
    assert m.select_var_threshold(1e300).size == 0                                    # nothing passes
    np.testing.assert_array_equal(m.select_var_threshold(-1.0), np.arange(30))        # everything passes, index order
    np.testing.assert_array_equal(m.select_var_threshold(50.0), O.select_var_threshold(gv, 50.0))


@pytest.mark.parametrize("shape", [(1, 1), (1, 500), (500, 1), (3, 70000)])
def test_degenerate_shapes(ffi, ctx, shape):
    rng = np.random.default_rng(36)
    n, mcols = shape
    a = random_csr(rng, n, mcols, 0.5 if n * mcols < 1000 else 0.01)
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    o = O.Compressed.from_scipy(a)
    for d in (ffi.ROW, ffi.COLUMN):
        np.testing.assert_array_equal(m.number(d), O.number(o, d))
        np.testing.assert_array_equal(m.sum(d), O.sum_(o, d))
        close(m.variance(d), O.variance(o, d), atol=1e-9)
    m.normalize_total_inplace(100.0, ffi.ROW)
    m.log1p_inplace()
    ol = O.log1p(O.normalize_total(o, 100.0, O.ROW))
    close(m.download()[2], ol.values, rtol=6e-7)
    close_compact_variance(m.variance(ffi.COLUMN), ol, O.COLUMN)


def test_line_lengths_around_the_warp_and_batch_boundaries(ffi, ctx):
    """Rows of exactly 0, 1, 7, 8, 9, 31, 32, 33, 255, 256, 257, 511, 512, 513 stored entries (lane groups of 8 / 32 and
    the 16-deep load batches of K1 / K4)."""
    lens = [0, 1, 7, 8, 9, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1024, 1500] * 3
    mcols = 2000
    rng = np.random.default_rng(37)
    indptr = np.concatenate([[0], np.cumsum(lens)])
    indices = np.concatenate([np.sort(rng.choice(mcols, L, replace=False)) for L in lens] + [np.zeros(0, np.int64)]).astype(np.int64)
    data = rng.integers(1, 30, size=indices.size).astype(np.float32)
    a = sp.csr_matrix((data, indices, indptr), shape=(len(lens), mcols))
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    o = O.Compressed.from_scipy(a)
    np.testing.assert_array_equal(m.number(ffi.ROW), np.array(lens, np.uint32))
    for d in (ffi.ROW, ffi.COLUMN):
        np.testing.assert_array_equal(m.sum(d), O.sum_(o, d))
        close(m.variance(d), O.variance(o, d), atol=1e-9)
        mn, mx = m.min_max(d)
        wmn, wmx = O.min_max(o, d)
        np.testing.assert_array_equal(mn, wmn)
        np.testing.assert_array_equal(mx, wmx)
    m.normalize_total_inplace(1e4, ffi.ROW)
    m.log1p_inplace()
    ol = O.log1p(O.normalize_total(o, 1e4, O.ROW))
    close(m.download()[2], ol.values, rtol=6e-7)
    close_compact_variance(m.variance(ffi.COLUMN), ol, O.COLUMN)
    close_compact_variance(m.variance(ffi.ROW), ol, O.ROW)


def test_handles_may_outlive_their_context(ffi):
    """A garbage-collected host frees handles in any order: a matrix freed after srb_ctx_destroy must neither crash nor
    leave a CUDA error behind for the next call (round-1 bug: 'invalid device ordinal' in an unrelated upload)."""
    a = random_csr(np.random.default_rng(3), 50, 20, 0.3)
    c1 = ffi.Context(0)
    m = ffi.DeviceMatrix.from_scipy(c1, a)
    m2 = m.clone()
    c1.close()
    with pytest.raises(ffi.SrbError) as e:   # using it is an error, not a crash
        m.sum(ffi.ROW)
    assert e.value.code == -1
    m.free(), m2.free()
    c2 = ffi.Context(0)
    try:
        w = ffi.DeviceMatrix.from_scipy(c2, a)   # would have reported the stale error
        np.testing.assert_array_equal(w.sum(ffi.ROW), O.sum_(O.Compressed.from_scipy(a), O.ROW))
        w.free()
    finally:
        c2.close()


def test_offsets_must_start_at_zero(ffi, ctx):
    """offsets[0] = p > 0 leaves p entries that belong to no line; nalgebra-sparse rejects it, so does the upload."""
    a = random_csr(np.random.default_rng(4), 30, 12, 0.4)
    off = a.indptr.astype(np.uint64).copy()
    off[0] = 2
    for mode in (ffi.UPLOAD_DEVICE_NARROW, ffi.UPLOAD_HOST_PACK, ffi.UPLOAD_HOST_PACK_DELTA):
        ctx.set_upload_mode(mode)
        try:
            with pytest.raises(ffi.SrbError) as e:
                ffi.DeviceMatrix.upload(ctx, ffi.CSR, 30, 12, off, a.indices.astype(np.uint64), a.data, nnz=a.nnz)
            assert e.value.code in (-1, -8), e.value   # INVALID_ARG or UNSUPPORTED (non-canonical)
        finally:
            ctx.set_upload_mode(ffi.UPLOAD_AUTO)


def test_pca_stream_fit_requires_every_cell(ffi, ctx):
    """srb_pca_stream_fit refuses a stream that has not seen exactly ncells_total cells."""
    rng = np.random.default_rng(5)
    a = random_csr(rng, 400, 60, 0.3)
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    cnt, s, q = m.gene_moments()
    sel = np.arange(20, dtype=np.uint64)
    ps = ffi.PcaStream(ctx, 60, 800, s, q, sel, 3)   # announces 800 cells, pushes 400
    ps.push_gram(m)
    with pytest.raises(ffi.SrbError) as e:
        ps.fit()
    assert e.value.code == -1
    ps.free()
    m.free()


def test_bulk_staged_row_reduce_irregular_long_rows(ffi, ctx):
    """K1 in its cp.async.bulk form (lines of >= 128 stored entries on average): rows shorter, equal to and longer than
    the 2048-value staging tile, empty rows, rows that straddle tile and CTA boundaries — sums bit-exact on integer data,
    and the range / sign / integrality flags it feeds to the fixed-point moments must lead to the oracle's per-gene
    statistics after normalise + log1p."""
    rng = np.random.default_rng(77)
    m = 9000
    lens = [0, 1, 3, 127, 128, 2047, 2048, 2049, 4096, 4100, 0, 0, 6000, 5, 8999, 9000] + list(rng.integers(0, 6000, 150))
    n = len(lens)
    rows, cols, vals = [], [], []
    for r, ln in enumerate(lens):
        c = np.sort(rng.choice(m, size=int(ln), replace=False))
        rows.append(np.full(c.size, r)), cols.append(c), vals.append(rng.integers(1, 40, c.size).astype(np.float32))
    a = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, m), dtype=np.float32)
    a.sort_indices()
    assert a.nnz / n >= 128
    o = O.Compressed.from_scipy(a)
    mt = ffi.DeviceMatrix.from_scipy(ctx, a)
    np.testing.assert_array_equal(mt.sum(ffi.ROW), O.sum_(o, O.ROW))
    np.testing.assert_array_equal(mt.number(ffi.ROW), O.number(o, O.ROW))
    mn, mx = mt.min_max(ffi.ROW)
    omn, omx = O.min_max(o, O.ROW)
    np.testing.assert_array_equal(mn, omn), np.testing.assert_array_equal(mx, omx)
    np.testing.assert_array_equal(mt.sum(ffi.COLUMN), O.sum_(o, O.COLUMN))          # exact fixed-point path (integer data)
    mt.normalize_total_inplace(1e4, ffi.ROW)
    mt.log1p_inplace()
    ol = O.log1p(O.normalize_total(o, 1e4, O.ROW))
    close(mt.sum(ffi.ROW), O.sum_(ol, O.ROW), rtol=1e-6)
    close_compact_variance(mt.variance(ffi.COLUMN), ol, O.COLUMN)
    # fractional / negative data must raise the flags (no fixed-point path): results still match the oracle
    b = a.copy().astype(np.float64)
    b.data = b.data * 0.37 - 3.0
    mb = ffi.DeviceMatrix.from_scipy(ctx, b.astype(np.float32))
    ob = O.Compressed.from_scipy(b.astype(np.float32))
    close(mb.sum(ffi.ROW), O.sum_(ob, O.ROW), rtol=1e-6)
    close(mb.variance(ffi.COLUMN), O.variance(ob, O.COLUMN), rtol=1e-5, atol=1e-9)
    mt.free(), mb.free()
