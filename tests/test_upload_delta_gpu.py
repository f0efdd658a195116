"""HOST_PACK_DELTA upload: the sorted indices of each line cross the link as one-byte gaps (+ an escape list) and are
rebuilt on the device. The device matrix must equal the caller's arrays exactly, and malformed input must still be
reported with the same status codes as in the other upload modes."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from tests._util import random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffi():
    from singlerust_b200 import _ffi
    return _ffi


@pytest.fixture(scope="module")
def ctx(ffi):
    c = ffi.Context(0)
    c.set_upload_mode(ffi.UPLOAD_HOST_PACK_DELTA)
    yield c
    c.close()


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("index_dtype", [np.uint64, np.uint32])
def test_small_matrices_round_trip(ffi, ctx, fmt, index_dtype):
    rng = np.random.default_rng(21)
    for shape, dens in (((700, 90), 0.15), ((40, 70_000), 0.002), ((5, 7), 0.0), ((1, 1), 1.0), ((64, 3000), 0.3)):
        a = random_csr(rng, shape[0], shape[1], dens, empty_rows=(0,) if shape[0] > 1 else ())
        if fmt == "csc":
            a = sp.csc_matrix(a)
            a.sort_indices()
        m = ffi.DeviceMatrix.from_scipy(ctx, a, index_dtype=index_dtype)
        off, idx, val = m.download()
        np.testing.assert_array_equal(off, a.indptr)
        np.testing.assert_array_equal(idx, a.indices)
        np.testing.assert_array_equal(val, a.data.astype(np.float64))
        assert ctx.last_upload()[1]
        np.testing.assert_array_equal(m.sum(ffi.ROW), np.asarray(a.astype(np.float64).sum(axis=1)).ravel())


def test_many_chunks_and_link_bytes(ffi, ctx):
    from singlerust_b200 import synth
    n, mg = 40_000, 30_000
    thr, amp = synth.gene_tables(mg, seed=3, mean_density=0.05)
    plain = ffi.Context(0)
    try:
        plain.set_upload_mode(ffi.UPLOAD_DEVICE_NARROW)
        src = ffi.DeviceMatrix.synth(plain, 0x5EED0011, n, mg, thr, amp)
        off, idx, val = src.download(values="f32")
        nnz = int(off[-1])
        m = ffi.DeviceMatrix.upload(ctx, ffi.CSR, n, mg, off, idx, val)
        o2, i2, v2 = m.download(values="f32")
        np.testing.assert_array_equal(o2, off)
        np.testing.assert_array_equal(i2, idx)
        np.testing.assert_array_equal(v2, val)
        h2d, packed = ctx.last_upload()
        assert packed and 8 * (n + 1) + 5 * nnz <= h2d < 8 * (n + 1) + 5 * nnz + 12 * (nnz // 100)   # 1 + 4 bytes per entry + few escapes
        np.testing.assert_array_equal(m.number(ffi.COLUMN), src.number(ffi.COLUMN))
    finally:
        plain.close()


def test_errors_are_the_same(ffi, ctx):
    a = random_csr(np.random.default_rng(1), 20, 10, 0.3)
    indptr = a.indptr.astype(np.uint64)
    bad = a.indices.astype(np.uint64).copy()
    bad[3] = 10
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, 10, indptr, bad, a.data)
    assert e.value.code == -3
    unsorted = a.indices.astype(np.uint64).copy()
    r0, r1 = int(a.indptr[0]), int(a.indptr[1])
    if r1 - r0 >= 2:
        unsorted[r0], unsorted[r0 + 1] = unsorted[r0 + 1], unsorted[r0]
        with pytest.raises(ffi.SrbError) as e:
            ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, 10, indptr, unsorted, a.data)
        assert e.value.code == -8
    wrong = indptr.copy()                      # non-monotone offsets: the delta coder declines, the device check reports
    j = int(np.nonzero(np.diff(a.indptr) > 0)[0][3])
    wrong[j], wrong[j + 1] = indptr[j + 1], indptr[j]
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, 20, 10, wrong, a.indices.astype(np.uint64), a.data)
    assert e.value.code == -8
    m = ffi.DeviceMatrix.from_scipy(ctx, a)   # the context stays usable
    np.testing.assert_array_equal(m.number(ffi.ROW), np.diff(a.indptr))
