"""GPU parity tests: every call goes through the C ABI (libsrb200.so) and is compared with the CPU oracle on the
same seeded inputs. Integer / index outputs are bit-exact; float outputs within the tolerance written in each test
(north_star: 1e-5 relative). Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle import pca_oracle as P
from tests._util import f, random_csr, rel_err, sign_align

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for float statistics / PCA


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


def upload(ffi, ctx, a, index_dtype=np.uint64):
    return ffi.DeviceMatrix.from_scipy(ctx, a, index_dtype=index_dtype)


def assert_close(got, want, rtol=RTOL, atol=0.0):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), "NaN pattern differs"
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf])
    ok = ~(nan_w | inf)
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=atol)


# ------------------------------------------------------------------------------------------------------
# known answers through the ABI
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int8, np.uint16, np.int32, np.uint32])
def test_kat_stats(ffi, ctx, kat, fmt, dtype):
    a = sp.csr_matrix((np.array(kat["data"], dtype=dtype), kat["indices"], kat["indptr"]), shape=kat["shape"])
    if fmt == "csc":
        a = a.tocsc()
    m = upload(ffi, ctx, a)
    o = O.Compressed.from_scipy(a)
    for d, key in ((ffi.ROW, "row"), (ffi.COLUMN, "column")):
        assert m.number(d).tolist() == kat["number"][key]
        np.testing.assert_array_equal(m.sum(d), f(kat["sum"][key]))
        assert_close(m.variance(d), O.variance(o, d), rtol=1e-12, atol=1e-15)
        assert_close(m.std_dev(d), O.std_dev(o, d), rtol=1e-12, atol=1e-15)
        mn, mx = m.min_max(d)
        np.testing.assert_array_equal(mn, f(kat["min"][key]))
        np.testing.assert_array_equal(mx, f(kat["max"][key]))
    if fmt == "csr":
        assert_close(m.variance(ffi.ROW), f(kat["variance"]["row"]), rtol=1e-12)
        assert_close(m.variance(ffi.COLUMN), f(kat["variance"]["column"]), rtol=1e-12, atol=1e-15)


def test_kat_normalize_log1p_hvg_densify(ffi, ctx_faithful, kat):
    a = sp.csr_matrix((np.array(kat["data"], dtype=np.float64), kat["indices"], kat["indptr"]), shape=kat["shape"])
    m = upload(ffi, ctx_faithful, a)
    m.normalize_total_inplace(10.0, ffi.ROW)
    _, _, v = m.download()
    np.testing.assert_allclose(v, kat["normalize_total_row_target10"]["values"], rtol=1e-15)
    m.log1p_inplace()
    _, _, v = m.download()
    np.testing.assert_allclose(v, kat["log1p_after_normalize"], rtol=1e-14)
    gv = m.variance(ffi.COLUMN)
    np.testing.assert_allclose(gv, kat["gene_variance_after_log1p"], rtol=1e-6, atol=1e-9)
    assert m.select_hvg(3).tolist() == kat["hvg_top3"]
    assert m.select_hvg(5).tolist() == kat["hvg_full_order"]
    d = m.densify_selected(kat["hvg_top3"])
    want = O.densify_selected(O.log1p(O.normalize_total(O.Compressed.from_scipy(a), 10.0, O.ROW)), np.arange(4), kat["hvg_top3"])
    np.testing.assert_allclose(d, want, rtol=1e-14)
    assert np.all(d[2] == 0)


# ------------------------------------------------------------------------------------------------------
# randomised differential tests vs the oracle
# ------------------------------------------------------------------------------------------------------
CASES = [
    # n, m, density, dtype, integer
    (257, 61, 0.20, np.float32, True),
    (1000, 300, 0.05, np.float64, False),
    (64, 2000, 0.08, np.float32, False),   # long rows -> 32 lanes per line
    (5000, 40, 0.30, np.uint8, True),
    (300, 12000, 0.02, np.float32, True),  # > one stripe of genes (fused kernel column striping)
    (33, 35000, 0.01, np.int16, True),     # 4 stripes
]


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1]}-{np.dtype(c[3]).name}" for c in CASES])
def test_stats_random(ffi, ctx, fmt, case):
    n, mcols, dens, dtype, integer = case
    rng = np.random.default_rng(n * 7 + mcols)
    a = random_csr(rng, n, mcols, dens, dtype=dtype, empty_rows=(1, n - 1), empty_cols=(0, mcols // 2), integer=integer)
    if fmt == "csc":
        a = a.tocsc()
        a.sort_indices()
    m = upload(ffi, ctx, a)
    o = O.Compressed.from_scipy(a)
    for d in (ffi.ROW, ffi.COLUMN):
        np.testing.assert_array_equal(m.number(d), O.number(o, d))        # bit-exact
        assert_close(m.sum(d), O.sum_(o, d), rtol=1e-7 if not integer else 0)
        assert_close(m.variance(d), O.variance(o, d), rtol=RTOL, atol=1e-9)
        assert_close(m.std_dev(d), O.std_dev(o, d), rtol=RTOL, atol=1e-6)
        mn, mx = m.min_max(d)
        wmn, wmx = O.min_max(o, d)
        np.testing.assert_array_equal(mn, wmn)                            # comparisons are exact
        np.testing.assert_array_equal(mx, wmx)
    q = m.qc_all()
    np.testing.assert_array_equal(q["num_per_cell"], O.number(o, O.ROW))
    np.testing.assert_array_equal(q["num_per_gene"], O.number(o, O.COLUMN))
    assert_close(q["variance_per_gene"], O.variance(o, O.COLUMN), rtol=RTOL, atol=1e-9)
    assert_close(q["std_dev_per_cell"], O.std_dev(o, O.ROW), rtol=RTOL, atol=1e-6)


def test_index_width_32_and_64_agree(ffi, ctx):
    rng = np.random.default_rng(5)
    a = random_csr(rng, 500, 80, 0.1)
    m64, m32 = upload(ffi, ctx, a, np.uint64), upload(ffi, ctx, a, np.uint32)
    np.testing.assert_array_equal(m64.sum(ffi.COLUMN), m32.sum(ffi.COLUMN))
    off, idx, val = m32.download()
    np.testing.assert_array_equal(off, a.indptr)
    np.testing.assert_array_equal(idx, a.indices)
    np.testing.assert_array_equal(val, a.data)


@pytest.mark.parametrize("mode", ["compact", "faithful"])
@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_normalize_log1p_random(ffi, ctx, ctx_faithful, mode, direction, fmt):
    c = ctx if mode == "compact" else ctx_faithful
    rng = np.random.default_rng(21 + direction)
    a = random_csr(rng, 700, 12000 if fmt == "csr" else 150, 0.03 if fmt == "csr" else 0.1, dtype=np.float32,
                   empty_rows=(5,), empty_cols=(7,))
    if fmt == "csc":
        a = a.tocsc()
        a.sort_indices()
    m = upload(ffi, c, a)
    o = O.Compressed.from_scipy(a)
    m.normalize_total_inplace(1e4, direction)
    on = O.normalize_total(o, 1e4, direction)
    tol = 3e-7 if mode == "compact" else 1e-14
    assert m.info()["value_dtype"] == (ffi.F32 if mode == "compact" else ffi.F64)
    _, _, v = m.download()
    assert_close(v, on.values, rtol=tol)
    # reference invariant (processing/mod.rs:420-481): non-empty line sums == target
    s = m.sum(direction)
    nonempty = O.number(o, direction) > 0
    np.testing.assert_allclose(s[nonempty], 1e4, rtol=2e-7 if mode == "compact" else 1e-12)
    assert np.all(s[~nonempty] == 0)
    m.log1p_inplace()
    ol = O.log1p(on)
    _, _, v = m.download()
    assert_close(v, ol.values, rtol=6e-7 if mode == "compact" else 1e-13)
    for d in (ffi.ROW, ffi.COLUMN):
        assert_close(m.variance(d), O.variance(ol, d), rtol=RTOL, atol=1e-10)
        assert_close(m.sum(d), O.sum_(ol, d), rtol=1e-6)
        np.testing.assert_array_equal(m.number(d), O.number(ol, d))


def test_log1p_only_keeps_f32_and_clone_is_copy_on_write(ffi, ctx):
    rng = np.random.default_rng(2)
    a = random_csr(rng, 200, 50, 0.2, dtype=np.float32)
    m = upload(ffi, ctx, a)
    cl = m.clone()
    cl.log1p_inplace()
    _, _, v = cl.download(values="f32")
    np.testing.assert_allclose(v, np.log1p(a.data), rtol=3e-7)
    assert cl.info()["value_dtype"] == ffi.F32           # transform/mod.rs:43-46
    _, _, v0 = m.download(values="f32")
    np.testing.assert_array_equal(v0, a.data)             # the original is untouched (deep_clone semantics)


def test_general_path_negative_values(ffi, ctx):
    """Negative / non-count data takes the fp64-atomic moments path; same results within tolerance."""
    rng = np.random.default_rng(9)
    a = random_csr(rng, 400, 90, 0.15, dtype=np.float64, integer=False)
    a.data -= 20.0
    m = upload(ffi, ctx, a)
    o = O.Compressed.from_scipy(a)
    for d in (ffi.ROW, ffi.COLUMN):
        assert_close(m.sum(d), O.sum_(o, d), rtol=1e-9, atol=1e-9)
        assert_close(m.variance(d), O.variance(o, d), rtol=RTOL, atol=1e-9)
        np.testing.assert_array_equal(m.number(d), O.number(o, d))
    mn, mx = m.min_max(ffi.COLUMN)
    wmn, wmx = O.min_max(o, O.COLUMN)
    np.testing.assert_array_equal(mn, wmn)
    np.testing.assert_array_equal(mx, wmx)


def test_errors(ffi, ctx):
    a = random_csr(np.random.default_rng(1), 20, 10, 0.3)
    bad = a.indices.astype(np.uint64).copy()
    bad[3] = 10
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, 10, a.indptr.astype(np.uint64), bad, a.data)
    assert e.value.code == -3
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, 10, a.indptr.astype(np.uint64), a.indices.astype(np.uint64), a.data.astype(np.int64))
    assert e.value.code == -2   # the reference panics for I64 (shared/mod.rs:117)
    unsorted = a.indices.astype(np.uint64).copy()
    r0, r1 = int(a.indptr[0]), int(a.indptr[1])
    if r1 - r0 >= 2:
        unsorted[r0], unsorted[r0 + 1] = unsorted[r0 + 1], unsorted[r0]
        with pytest.raises(ffi.SrbError) as e:
            ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, 10, a.indptr.astype(np.uint64), unsorted, a.data)
        assert e.value.code == -8
    m = upload(ffi, ctx, a)
    with pytest.raises(ffi.SrbError):
        m.sum(2)
    with pytest.raises(ffi.SrbError) as e:
        m.densify_selected([3, 11])
    assert e.value.code == -3   # "Index out of bounds", shared/utils/mod.rs:8-13


def test_empty_matrix(ffi, ctx):
    a = sp.csr_matrix((5, 7), dtype=np.float32)
    m = upload(ffi, ctx, a)
    assert m.number(ffi.ROW).tolist() == [0] * 5 and m.number(ffi.COLUMN).tolist() == [0] * 7
    assert np.all(np.isnan(m.variance(ffi.ROW))) and np.all(m.variance(ffi.COLUMN) == 0)
    mn, mx = m.min_max(ffi.COLUMN)
    assert np.all(np.isposinf(mn)) and np.all(np.isneginf(mx))
    m.normalize_total_inplace(1e4, ffi.ROW)
    m.log1p_inplace()
    assert np.all(m.sum(ffi.ROW) == 0)


# ------------------------------------------------------------------------------------------------------
# synthetic generator: device == CPU twin, bit for bit
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("skew", [False, True])
def test_synth_device_equals_cpu(ffi, ctx, skew):
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(3000, seed=11, mean_density=0.05)
    m = ffi.DeviceMatrix.synth(ctx, 0x5EED0001, 700, 3000, thr, amp, row0=123, skew=skew)
    o = O.synth_csr(0x5EED0001, 700, 3000, thr, amp, row0=123, skew=skew)
    off, idx, val = m.download(values="f32")
    np.testing.assert_array_equal(off, o.offsets)
    np.testing.assert_array_equal(idx, o.indices)
    np.testing.assert_array_equal(val, o.values)


# ------------------------------------------------------------------------------------------------------
# config S of BASELINE.json: per-gene mean/variance on 10k x 2k, 5 % nnz (tests/test_basic_stats path)
# ------------------------------------------------------------------------------------------------------
def test_config_S_gene_mean_variance(ffi, ctx):
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(2000, seed=0x5EED0001, mean_density=0.05)
    m = ffi.DeviceMatrix.synth(ctx, 0x5EED0001, 10000, 2000, thr, amp)
    o = O.synth_csr(0x5EED0001, 10000, 2000, thr, amp)
    cnt, s = m.number(ffi.COLUMN), m.sum(ffi.COLUMN)
    np.testing.assert_array_equal(cnt, O.number(o, O.COLUMN))
    np.testing.assert_array_equal(s, O.sum_(o, O.COLUMN))  # integer counts: sums are exact on both sides
    assert_close(m.variance(ffi.COLUMN), O.variance(o, O.COLUMN), rtol=1e-9, atol=1e-12)
    assert_close(m.variance(ffi.ROW), O.variance(o, O.ROW), rtol=1e-12)
    np.testing.assert_array_equal(m.number(ffi.ROW), O.number(o, O.ROW))


# ------------------------------------------------------------------------------------------------------
# chunked accumulation (backed path)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("chunk", [1, 37, 1000])
def test_chunk_stream_matches_whole(ffi, ctx, chunk):
    rng = np.random.default_rng(4)
    a = random_csr(rng, 230, 64, 0.15, dtype=np.float32)
    o = O.Compressed.from_scipy(a)
    st = ffi.ChunkStream(ctx, ffi.CSR, 230, 64)
    for s in range(0, 230, chunk):
        ch = a[s:s + chunk]
        st.push(ch.indptr, ch.indices, ch.data)
    for d in (ffi.ROW, ffi.COLUMN):
        np.testing.assert_array_equal(st.number(d), O.number(o, d))
        assert_close(st.sum(d), O.sum_(o, d), rtol=1e-12)
        assert_close(st.variance(d), O.variance(o, d), rtol=RTOL, atol=1e-10)


# ------------------------------------------------------------------------------------------------------
# HVG + PCA
# ------------------------------------------------------------------------------------------------------
def clustered_counts(rng, n, m, groups=6, dens=0.08):
    """Count matrix with planted group structure so the leading principal components are well separated."""
    g = rng.integers(0, groups, size=n)
    base = rng.uniform(0.3, 1.7, size=(groups, m)) ** 3
    rate = dens * base[g] * rng.uniform(0.6, 1.6, size=(n, 1))
    x = rng.poisson(rate * 6.0) * (rng.random((n, m)) < np.clip(rate * 4, 0, 0.9))
    a = sp.csr_matrix(x.astype(np.float32))
    a.sort_indices()
    return a


def well_separated(eigvals, k, rel_gap=1e-3):
    ev = np.asarray(eigvals)
    ok = []
    for j in range(k):
        gaps = []
        if j > 0:
            gaps.append(ev[j - 1] - ev[j])
        if j + 1 < ev.size:
            gaps.append(ev[j] - ev[j + 1])
        ok.append(min(gaps) > rel_gap * ev[j])
    return np.array(ok)


@pytest.mark.parametrize("center,scale", [(True, True), (True, False), (False, True), (False, False)])
def test_pca_fp64_path_matches_oracle(ffi, ctx_faithful, center, scale):
    rng = np.random.default_rng(31)
    a = clustered_counts(rng, 3000, 400)
    m = upload(ffi, ctx_faithful, a)
    o = O.Compressed.from_scipy(a)
    m.normalize_total_inplace(1e4, ffi.ROW)
    m.log1p_inplace()
    ol = O.log1p(O.normalize_total(o, 1e4, O.ROW))
    sel = O.select_hvg(O.variance(ol, O.COLUMN), 100)
    got_sel = m.select_hvg(100)
    np.testing.assert_array_equal(got_sel, sel)          # index bookkeeping: bit-exact, incl. order
    want = P.pca_pipeline(ol, 100, 10, center, scale, selection=sel)
    got = m.pca(sel, 10, center, scale, gram_mode=ffi.GRAM_FP64)
    np.testing.assert_allclose(got["explained_variance_ratio"], want["explained_variance_ratio"], rtol=RTOL)
    good = well_separated(want["eigenvalues"], 10)
    assert good[:5].all()
    comps = sign_align(got["components"], want["components"])
    scores = sign_align(got["scores"], want["scores"])
    for j in np.nonzero(good)[0]:
        scale_c = np.linalg.norm(want["components"][:, j])
        assert np.max(np.abs(comps[:, j] - want["components"][:, j])) <= RTOL * scale_c
        scale_s = np.linalg.norm(want["scores"][:, j]) / np.sqrt(a.shape[0])
        assert np.max(np.abs(scores[:, j] - want["scores"][:, j])) <= 10 * RTOL * scale_s
    # densify parity (selection order, zeros elsewhere)
    np.testing.assert_allclose(m.densify_selected(sel[:17]), O.densify_selected(ol, np.arange(3000), sel[:17]), rtol=1e-13)


def test_pipeline_call_equals_separate_calls(ffi, ctx):
    rng = np.random.default_rng(32)
    a = clustered_counts(rng, 2000, 300)
    m1, m2 = upload(ffi, ctx, a), upload(ffi, ctx, a)
    r1 = m1.pipeline_normalize_hvg_pca(1e4, 64, 8, gram_mode=ffi.GRAM_FP64)
    m2.normalize_total_inplace(1e4, ffi.ROW)
    m2.log1p_inplace()
    sel = m2.select_hvg(64)
    r2 = m2.pca(sel, 8, gram_mode=ffi.GRAM_FP64)
    np.testing.assert_array_equal(r1["selection"], sel)
    np.testing.assert_allclose(r1["explained_variance_ratio"], r2["explained_variance_ratio"], rtol=1e-12)
    np.testing.assert_allclose(r1["scores"], r2["scores"], rtol=1e-9, atol=1e-9)
    st = ctx.last_stage_ms()
    assert st["gram"] > 0 and st["eig"] > 0


@pytest.mark.parametrize("n,m,n_top,k", [(3000, 400, 100, 10), (20000, 900, 600, 20), (1000, 300, 257, 5)])
def test_pca_tensor_core_gram_matches_fp64_and_oracle(ffi, ctx, n, m, n_top, k):
    """tcgen05 Gram (split-fp16 operands, fp32 TMEM chunks, fp64 across chunks) vs the fp64 CUDA-core path and the
    oracle's exact SVD. Tolerance: north_star 1e-5 relative on loadings / explained variance."""
    rng = np.random.default_rng(77 + n)
    a = clustered_counts(rng, n, m)
    mt, mf = upload(ffi, ctx, a), upload(ffi, ctx, a)
    for mm in (mt, mf):
        mm.normalize_total_inplace(1e4, ffi.ROW)
        mm.log1p_inplace()
    sel = mt.select_hvg(n_top)
    rt = mt.pca(sel, k, gram_mode=ffi.GRAM_TENSOR)
    rf = mf.pca(sel, k, gram_mode=ffi.GRAM_FP64)
    print("evr rel diff tensor vs fp64:", np.max(np.abs(rt["explained_variance_ratio"] / rf["explained_variance_ratio"] - 1)))
    # oracle on the device's (f32-stored) values so the comparison isolates the PCA stage
    off, idx, val = mt.download()
    ol = O.Compressed("csr", n, m, off, idx, val)
    want = P.pca_pipeline(ol, n_top, k, selection=sel)
    good = well_separated(want["eigenvalues"], k)
    comps = sign_align(rt["components"], want["components"])
    scores = sign_align(rt["scores"], want["scores"])
    cerr = [float(np.max(np.abs(comps[:, j] - want["components"][:, j]))) for j in range(k)]
    serr = [float(np.max(np.abs(scores[:, j] - want["scores"][:, j])) / (np.linalg.norm(want["scores"][:, j]) / np.sqrt(n))) for j in range(k)]
    print("component max-abs errors:", ["%.1e" % e for e in cerr], "score errors / rms:", ["%.1e" % e for e in serr], "good:", good)
    np.testing.assert_allclose(rt["explained_variance_ratio"], rf["explained_variance_ratio"], rtol=RTOL)
    np.testing.assert_allclose(rt["explained_variance_ratio"], want["explained_variance_ratio"], rtol=RTOL)
    for j in np.nonzero(good)[0]:
        assert cerr[j] <= RTOL, (j, cerr)
        assert serr[j] <= 10 * RTOL, (j, serr)
    assert good[:3].all()


def test_config_L_slice_against_the_oracle(ffi, ctx):
    """The bench configuration itself — 30 k genes, HVG 2000, 50 PCs, the default K8 solver (ChFSI at d >= 1024) and the
    tcgen05 Gram / scores — on the first 16 384 cells of the L matrix (seed 0x5EED0002), against the oracle's SVD PCA.
    The spectrum of this synthetic block is flat (no cell programmes), so individual loadings are only compared where
    the eigengap allows; every component is additionally checked through its eigen-residual on the ORACLE's matrix,
    which does not depend on the gaps: || Z^T Z v - lambda v || <= 1e-5 lambda."""
    from singlerust_b200 import synth
    n, m, d, k = 16_384, 30_000, 2_000, 50
    thr, amp = synth.gene_tables(m, seed=0x5EED0002, mean_density=0.05)
    mat = ffi.DeviceMatrix.synth(ctx, 0x5EED0002, n, m, thr, amp)
    cpu = O.synth_csr(0x5EED0002, n, m, thr, amp)
    off, idx, val = mat.download(values="f32")
    assert np.array_equal(off, cpu.offsets) and np.array_equal(idx, cpu.indices) and np.array_equal(val, cpu.values)
    res = mat.pipeline_normalize_hvg_pca(1e4, d, k)
    assert ctx.last_eig()["solver"] in ("chfsi", "chfsi->syevd")
    # HVG list: bit-exact against the oracle run on the device's stored (f32) values, set-equal against the f64 oracle
    off, idx, val = mat.download()
    ol = O.Compressed("csr", n, m, off, idx, val)
    np.testing.assert_array_equal(res["selection"], O.select_hvg(O.variance(ol, O.COLUMN), d))
    ol64 = O.log1p(O.normalize_total(cpu, 1e4, O.ROW))
    want_sel = O.select_hvg(O.variance(ol64, O.COLUMN), d)
    assert len(set(res["selection"].tolist()) ^ set(want_sel.tolist())) <= 2   # f32 storage may swap a near-tie at the cut
    want = P.pca_pipeline(ol, d, k, selection=res["selection"])
    np.testing.assert_allclose(res["explained_variance_ratio"], want["explained_variance_ratio"], rtol=RTOL)
    V = res["components"]
    np.testing.assert_allclose(V.T @ V, np.eye(k), atol=1e-9)
    # gap-independent: residual of every returned pair on the oracle's standardised block
    dense = O.densify_selected(ol, np.arange(n, dtype=np.uint64), res["selection"])
    Z = (dense - want["mean"]) / want["std"]
    lam = want["eigenvalues"][:k] * (n - 1)
    R = Z.T @ (Z @ V) - V * lam
    rel = np.linalg.norm(R, axis=0) / lam
    print("eigen-residuals on the oracle matrix: max %.2e" % rel.max())
    assert rel.max() <= RTOL
    scores = sign_align(res["scores"], Z @ sign_align(V, want["components"]))
    good = well_separated(want["eigenvalues"], k)
    comps = sign_align(V, want["components"])
    for j in np.nonzero(good)[0]:
        assert np.max(np.abs(comps[:, j] - want["components"][:, j])) <= RTOL, j
    # scores = Z V for the RETURNED loadings (exact relation, no gap involved)
    ZV = Z @ V
    np.testing.assert_allclose(res["scores"], ZV, rtol=0, atol=10 * RTOL * np.abs(ZV).max())
    del scores
    mat.free()


def test_full_size_config_L_properties(ffi, ctx):
    """BASELINE.json configs[1]/[2] at FULL size (1M cells x 30k genes, ~1.5 G nnz) through size-independent properties:
    count conservation, the reference's normalisation invariant, and the PCA identities
    V^T V = I, scores^T scores = diag(lambda), lambda_j = evr_j * trace, trace = n*d for standardised columns."""
    from singlerust_b200 import synth
    n, m, d, k = 1_000_000, 30_000, 2000, 50
    thr, amp = synth.gene_tables(m, seed=0x5EED0002, mean_density=0.05)
    mat = ffi.DeviceMatrix.synth(ctx, 0x5EED0002, n, m, thr, amp)
    nnz = mat.info()["nnz"]
    assert 0.045 < nnz / (n * m) < 0.055
    per_cell, per_gene = mat.number(ffi.ROW), mat.number(ffi.COLUMN)
    assert int(per_cell.astype(np.uint64).sum()) == nnz == int(per_gene.astype(np.uint64).sum())   # bit-exact bookkeeping
    raw_cell, raw_gene = mat.sum(ffi.ROW), mat.sum(ffi.COLUMN)
    assert raw_cell.sum() == raw_gene.sum()                       # integer counts: both totals exact in f64
    work = mat.clone()
    work.normalize_total_inplace(1e4, ffi.ROW)
    s = work.sum(ffi.ROW)
    assert np.all(per_cell > 0)
    np.testing.assert_allclose(s, 1e4, rtol=3e-7)                 # processing/mod.rs:451-462 invariant (f32 storage)
    np.testing.assert_allclose(work.sum(ffi.COLUMN).sum(), 1e4 * n, rtol=1e-7)
    np.testing.assert_array_equal(work.number(ffi.COLUMN), per_gene)   # the pattern is untouched
    work.log1p_inplace()
    gv = work.variance(ffi.COLUMN)
    assert np.all(gv >= 0) and np.all(np.isfinite(gv))
    sel = work.select_hvg(d)
    assert len(set(sel.tolist())) == d
    order = np.argsort(-gv, kind="stable")[:d]
    np.testing.assert_array_equal(sel, order.astype(np.uint64))   # descending variance, ties by index
    res = work.pca(sel, k)
    V, S, evr = res["components"], res["scores"], res["explained_variance_ratio"]
    np.testing.assert_allclose(V.T @ V, np.eye(k), atol=1e-9)
    lam = evr * (n * d)                                           # trace(Z^T Z) = n * d when scale=True
    G = S.T @ S
    np.testing.assert_allclose(np.diag(G), lam, rtol=2e-5)
    off = G - np.diag(np.diag(G))
    assert np.max(np.abs(off)) <= 2e-5 * lam[0]
    mean_err = np.max(np.abs(S.mean(axis=0)) / np.sqrt(lam / n))   # centred scores: column means ~ 0 (vs the column rms)
    assert mean_err <= 2e-5, mean_err
    assert np.all(np.diff(evr) <= 1e-12) and 0 < evr.sum() < 1


def test_csc_densify_and_pca_match_csr_and_oracle(ffi, ctx_faithful):
    """CSC-stored X (csc.rs / convert_to_array_f64_csc_selected, shared/mod.rs:261-290): same dense block and PCA as CSR."""
    rng = np.random.default_rng(41)
    a = clustered_counts(rng, 1500, 260)
    ac = a.tocsc()
    ac.sort_indices()
    mr, mc = upload(ffi, ctx_faithful, a), upload(ffi, ctx_faithful, ac)
    oc = O.Compressed.from_scipy(ac)
    for mm in (mr, mc):
        mm.normalize_total_inplace(1e4, ffi.ROW)
        mm.log1p_inplace()
    ocl = O.log1p(O.normalize_total(oc, 1e4, O.ROW))
    sel_r, sel_c = mr.select_hvg(40), mc.select_hvg(40)
    # CSC/Column variance is the two-pass major form (csc.rs:161-173); same ranking on this data as the oracle's
    np.testing.assert_array_equal(sel_c, O.select_hvg(O.variance(ocl, O.COLUMN), 40))
    dc = mc.densify_selected(sel_c)
    np.testing.assert_allclose(dc, O.densify_selected(ocl, np.arange(1500), sel_c), rtol=1e-13)
    np.testing.assert_allclose(dc, mr.densify_selected(sel_c), rtol=1e-13)
    rc, rr = mc.pca(sel_c, 6, gram_mode=ffi.GRAM_FP64), mr.pca(sel_c, 6, gram_mode=ffi.GRAM_FP64)
    np.testing.assert_allclose(rc["explained_variance_ratio"], rr["explained_variance_ratio"], rtol=1e-10)
    np.testing.assert_allclose(sign_align(rc["scores"], rr["scores"]), rr["scores"], atol=1e-7)
    want = P.pca_pipeline(ocl, 40, 6, selection=sel_c)
    np.testing.assert_allclose(rc["explained_variance_ratio"], want["explained_variance_ratio"], rtol=RTOL)
