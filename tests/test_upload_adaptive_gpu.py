"""HOST_PACK_ADAPTIVE upload: values are packed only for the chunks during which the host is ahead of the link. Whatever
mix of packed and raw chunks results, the device matrix must equal the host arrays bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_adaptive_value_packing_is_lossless():
    from singlerust_b200 import _ffi
    n, m, per = 60_000, 30_000, 300
    rng = np.random.default_rng(8)
    cols = np.sort(rng.integers(0, m // per, size=(n, per)) + np.arange(per) * (m // per), axis=1)
    nnz = n * per                                             # 18 M entries = 5 chunks
    val = rng.integers(1, 200, nnz).astype(np.float32)
    val[nnz - 3] = 0.25                                       # the last chunk can never be packed
    off = np.arange(0, nnz + 1, per, dtype=np.uint64)
    idx = cols.ravel().astype(np.uint64)
    ctx = _ffi.Context(0)
    try:
        ctx.set_upload_mode(_ffi.UPLOAD_HOST_PACK_ADAPTIVE)
        for _ in range(2):
            mt = _ffi.DeviceMatrix.upload(ctx, _ffi.CSR, n, m, off, idx, val)
            o2, i2, v2 = mt.download(values="f32")
            np.testing.assert_array_equal(o2, off)
            np.testing.assert_array_equal(i2, idx)
            np.testing.assert_array_equal(v2.view(np.uint32), val.view(np.uint32))
            h2d, packed = ctx.last_upload()
            assert packed and 8 * (n + 1) + 3 * nnz <= h2d <= 8 * (n + 1) + 6 * nnz
            np.testing.assert_array_equal(mt.sum(_ffi.ROW), val.reshape(n, per).astype(np.float64).sum(axis=1))
            mt.free()
    finally:
        ctx.close()
