"""Host-side index packing of the HOST_PACK upload (host_pack.cpp): pure host code, runs without a GPU."""
import numpy as np
import pytest

from singlerust_b200 import _ffi


@pytest.mark.parametrize("src_dtype", [np.uint64, np.uint32])
@pytest.mark.parametrize("dst_width", [2, 4])
@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 3 * 65536 + 17, 2_500_000])
@pytest.mark.parametrize("threads", [1, 0, 5])
def test_pack_is_a_plain_narrowing(src_dtype, dst_width, n, threads):
    rng = np.random.default_rng(n + dst_width)
    bound = 30_000 if dst_width == 2 else 3_000_000
    src = rng.integers(0, bound, size=n).astype(src_dtype)
    dst, oob = _ffi.host_pack_indices(src, dst_width, bound, threads)
    assert not oob
    np.testing.assert_array_equal(dst, src.astype(dst.dtype))


@pytest.mark.parametrize("bound", [1, 10, 30_000, 65_536, 70_000, (1 << 32) - 1])
def test_bounds_flag_is_exact(bound):
    n = 400_000
    rng = np.random.default_rng(bound % 97)
    base = rng.integers(0, bound, size=n, dtype=np.uint64)
    base[rng.integers(0, n)] = bound - 1  # the largest legal index
    dw = 2 if bound <= 65_536 else 4
    assert _ffi.host_pack_indices(base, dw, bound)[1] is False
    for bad in (bound, bound + 1, (1 << 32) + 3, (1 << 63) + 1, (1 << 64) - 1):
        for pos in (0, n // 2 + 1, n - 1):
            x = base.copy()
            x[pos] = bad
            assert _ffi.host_pack_indices(x, dw, bound)[1] is True, (bound, bad, pos)
    if bound < (1 << 32) - 1:
        x32 = base.astype(np.uint32)
        assert _ffi.host_pack_indices(x32, dw, bound)[1] is False
        x32[7] = bound
        assert _ffi.host_pack_indices(x32, dw, bound)[1] is True


def test_bad_arguments():
    lib = _ffi.lib()
    a = np.zeros(4, np.uint64)
    assert lib.srb_host_pack_indices(_ffi._ptr(a), 3, 4, _ffi._ptr(a), 2, 10, 1, None) == -1
    assert lib.srb_host_pack_indices(_ffi._ptr(a), 8, 4, _ffi._ptr(a), 8, 10, 1, None) == -1
    assert lib.srb_host_pack_indices(None, 8, 4, _ffi._ptr(a), 2, 10, 1, None) == -1


@pytest.mark.parametrize("dst_width", [1, 2])
@pytest.mark.parametrize("n", [0, 1, 65, 3 * 65536 + 17, 2_500_000])
def test_value_pack_round_trips_counts(dst_width, n):
    rng = np.random.default_rng(n + dst_width)
    hi = 256 if dst_width == 1 else 65536
    v = rng.integers(0, hi, size=n).astype(np.float32)
    if n:
        v[-1] = hi - 1
    d, ok = _ffi.host_pack_values_f32(v, dst_width)
    assert ok
    np.testing.assert_array_equal(d.astype(np.float32).view(np.uint32), v.view(np.uint32))  # bit for bit


@pytest.mark.parametrize("bad", [0.5, -1.0, -0.0, np.nan, np.inf, -np.inf, 65536.0, 65535.5, 1e30, -1e30, 3.0000002, 1e-45])
def test_value_pack_refuses_anything_lossy(bad):
    n = 300_000
    v = np.random.default_rng(3).integers(0, 200, size=n).astype(np.float32)
    for pos in (0, n // 2 + 3, n - 1):
        x = v.copy()
        x[pos] = bad
        assert _ffi.host_pack_values_f32(x, 1)[1] is False
        assert _ffi.host_pack_values_f32(x, 2)[1] is False
    x = v.copy()
    x[7] = 256.0
    assert _ffi.host_pack_values_f32(x, 1)[1] is False and _ffi.host_pack_values_f32(x, 2)[1] is True
