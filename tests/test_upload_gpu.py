"""HOST_PACK upload (srb_ctx_set_upload_mode): the index array is narrowed on the host before it crosses PCIe.
Whatever the mode, caller memory (pageable or pinned), index width or chunking, the device-resident matrix — and so
every statistic — must be bit-identical."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from tests._util import random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffi():
    from singlerust_b200 import _ffi
    return _ffi


@pytest.fixture(scope="module")
def ctxs(ffi):
    a, b, c = ffi.Context(0), ffi.Context(0), ffi.Context(0)
    a.set_upload_mode(ffi.UPLOAD_DEVICE_NARROW)
    b.set_upload_mode(ffi.UPLOAD_HOST_PACK)
    c.set_upload_mode(ffi.UPLOAD_HOST_PACK_VALUES)
    yield a, b, c
    a.close()
    b.close()
    c.close()


def same_matrix(ma, mb):
    for x, y in zip(ma.download(), mb.download()):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("index_dtype", [np.uint64, np.uint32])
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.uint16, np.int32])
def test_modes_agree_small(ffi, ctxs, fmt, index_dtype, dtype):
    rng = np.random.default_rng(11)
    a = random_csr(rng, 700, 90, 0.15, dtype=dtype, empty_rows=(3,), empty_cols=(7,))
    if fmt == "csc":
        a = sp.csc_matrix(a)
        a.sort_indices()
    ma = ffi.DeviceMatrix.from_scipy(ctxs[0], a, index_dtype=index_dtype)
    mb = ffi.DeviceMatrix.from_scipy(ctxs[1], a, index_dtype=index_dtype)
    same_matrix(ma, mb)
    off, idx, val = mb.download()
    np.testing.assert_array_equal(off, a.indptr)
    np.testing.assert_array_equal(idx, a.indices)
    np.testing.assert_array_equal(val, a.data.astype(np.float64))
    for d in (ffi.ROW, ffi.COLUMN):
        np.testing.assert_array_equal(ma.number(d), mb.number(d))
        np.testing.assert_array_equal(ma.sum(d), mb.sum(d))


def test_wide_minor_dimension_uses_4_byte_packing(ffi, ctxs):
    """nminor > 65 536: the packed width is 4 bytes and the copy lands in the index buffer directly."""
    rng = np.random.default_rng(12)
    n, m, per = 300, 200_000, 40
    cols = np.sort(np.stack([rng.choice(m, per, replace=False) for _ in range(n)]), axis=1)
    cols[0, -1] = m - 1  # the largest legal index
    a = sp.csr_matrix((rng.integers(1, 9, n * per).astype(np.float32), cols.ravel(), np.arange(0, n * per + 1, per)), shape=(n, m))
    ma, mb = ffi.DeviceMatrix.from_scipy(ctxs[0], a), ffi.DeviceMatrix.from_scipy(ctxs[1], a)
    same_matrix(ma, mb)
    np.testing.assert_array_equal(mb.sum(ffi.COLUMN), np.asarray(a.sum(axis=0)).ravel())


@pytest.mark.parametrize("bound_case", ["u16", "u32"])
def test_out_of_bounds_index_is_reported(ffi, ctxs, bound_case):
    ncols = 10 if bound_case == "u16" else 100_000
    a = random_csr(np.random.default_rng(1), 20, 10, 0.3)
    bad = a.indices.astype(np.uint64).copy()
    bad[3] = ncols
    for ctx in ctxs[:2]:
        with pytest.raises(ffi.SrbError) as e:
            ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, ncols, a.indptr.astype(np.uint64), bad, a.data)
        assert e.value.code == -3
    huge = a.indices.astype(np.uint64).copy()
    huge[0] = (1 << 63) + 2  # would alias a small index after narrowing
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctxs[1], ffi.CSR, 20, ncols, a.indptr.astype(np.uint64), huge, a.data)
    assert e.value.code == -3
    # the context stays usable after the error
    m = ffi.DeviceMatrix.from_scipy(ctxs[1], a)
    np.testing.assert_array_equal(m.number(ffi.ROW), np.diff(a.indptr))


@pytest.mark.parametrize("pinned", [False, True])
def test_many_chunks_ring_reuse(ffi, ctxs, pinned):
    """60 M entries = 15 chunks of the 4-slot staging ring; pageable (NumPy) and pinned (torch) caller memory."""
    from singlerust_b200 import synth
    n, m = 40_000, 30_000
    thr, amp = synth.gene_tables(m, seed=3, mean_density=0.05)
    src = ffi.DeviceMatrix.synth(ctxs[0], 0x5EED0011, n, m, thr, amp)
    off, idx, val = src.download(values="f32")
    nnz = int(off[-1])
    assert nnz > 4 * (1 << 22)
    off64, idx64 = off.astype(np.uint64), idx.astype(np.uint64)
    keep = None
    if pinned:
        import torch
        keep = [torch.from_numpy(x).pin_memory() for x in (off64.view(np.int64), idx64.view(np.int64), val)]
        args = (keep[0], keep[1], keep[2])
        mb = ffi.DeviceMatrix.upload(ctxs[1], ffi.CSR, n, m, *args, nnz=nnz, idx_width=8, dtype=ffi.F32)
    else:
        mb = ffi.DeviceMatrix.upload(ctxs[1], ffi.CSR, n, m, off64, idx64, val)
    o2, i2, v2 = mb.download(values="f32")
    np.testing.assert_array_equal(o2, off)
    np.testing.assert_array_equal(i2, idx)
    np.testing.assert_array_equal(v2, val)
    np.testing.assert_array_equal(mb.sum(ffi.COLUMN), src.sum(ffi.COLUMN))
    np.testing.assert_array_equal(mb.number(ffi.COLUMN), src.number(ffi.COLUMN))
    # a second upload on the same context reuses the ring
    mc = ffi.DeviceMatrix.upload(ctxs[1], ffi.CSR, n, m, off64, idx64, val)
    np.testing.assert_array_equal(mc.download(values="f32")[1], idx)


def test_chunk_stream_in_pack_mode(ffi, ctxs):
    rng = np.random.default_rng(9)
    a = random_csr(rng, 1500, 120, 0.1)
    cpu = O.Compressed.from_scipy(a)
    st = ffi.ChunkStream(ctxs[1], ffi.CSR, a.shape[0], a.shape[1])
    for r0 in range(0, a.shape[0], 400):
        c = a[r0:r0 + 400]
        st.push(c.indptr, c.indices, c.data)
    np.testing.assert_array_equal(st.number(ffi.COLUMN), O.number(cpu, O.COLUMN))
    np.testing.assert_array_equal(st.sum(ffi.ROW), O.sum_(cpu, O.ROW))
    st.free()


@pytest.mark.parametrize("kind", ["u8", "u16", "fraction_late", "negative_zero", "mixed_widths"])
def test_value_packing_is_lossless(ffi, ctxs, kind):
    """f32 values travel as u8 / u16 per chunk when every value of the chunk is such an integer; anything else goes raw.
    9 M entries = 3 chunks, so a late chunk can refuse after earlier ones were packed."""
    n, m, per = 30_000, 30_000, 300
    rng = np.random.default_rng(5)
    cols = np.sort(rng.integers(0, m // per, size=(n, per)) + np.arange(per) * (m // per), axis=1)
    nnz = n * per
    val = rng.integers(1, 200, nnz).astype(np.float32)
    if kind == "u16":
        val[nnz // 3] = 40_000
    elif kind == "fraction_late":
        val[nnz - 5] = 2.5
    elif kind == "negative_zero":
        val[17] = -0.0
    elif kind == "mixed_widths":
        val[nnz // 2] = 300      # second chunk needs u16
        val[nnz - 9] = 70_000    # last chunk goes raw
    off = np.arange(0, nnz + 1, per, dtype=np.uint64)
    idx = cols.ravel().astype(np.uint64)
    ma = ffi.DeviceMatrix.upload(ctxs[0], ffi.CSR, n, m, off, idx, val)
    mb = ffi.DeviceMatrix.upload(ctxs[2], ffi.CSR, n, m, off, idx, val)
    va, vb = ma.download(values="f32")[2], mb.download(values="f32")[2]
    np.testing.assert_array_equal(vb.view(np.uint32), val.view(np.uint32))   # bit for bit, including -0.0
    np.testing.assert_array_equal(va.view(np.uint32), vb.view(np.uint32))
    h2d, packed = ctxs[2].last_upload()
    assert packed
    raw = 8 * (n + 1) + 2 * nnz + 4 * nnz
    if kind == "u8":
        assert h2d == 8 * (n + 1) + 2 * nnz + nnz
    elif kind in ("fraction_late", "mixed_widths", "u16"):
        assert 8 * (n + 1) + 3 * nnz < h2d < raw
    else:
        assert h2d == raw
    np.testing.assert_array_equal(ma.sum(ffi.ROW), mb.sum(ffi.ROW))
