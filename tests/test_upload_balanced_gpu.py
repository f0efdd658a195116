"""BALANCED upload (what AUTO resolves to): every 4 M-entry chunk crosses PCIe either packed on the host (delta-coded
indices, u8 / u16 count values) or raw, decided from the measured packing time against the queued link work. Whatever
mix results — all packed (slow link), almost all raw (fast link), pageable or pinned caller memory, u64 or u32 indices,
delta or plain narrowing — the device matrix must equal the host arrays bit for bit and errors keep their codes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffi():
    from singlerust_b200 import _ffi
    return _ffi


@pytest.fixture(scope="module")
def ctx(ffi):
    c = ffi.Context(0)
    c.set_upload_mode(ffi.UPLOAD_BALANCED)
    yield c
    c.close()


@pytest.fixture(scope="module")
def data():
    n, m, per = 60_000, 30_000, 300
    rng = np.random.default_rng(9)
    cols = np.sort(rng.integers(0, m // per, size=(n, per)) + np.arange(per) * (m // per), axis=1)
    nnz = n * per                                             # 18 M entries = 5 chunks
    val = rng.integers(1, 200, nnz).astype(np.float32)
    val[nnz - 3] = 0.25                                       # the last chunk's values can never be packed
    off = np.arange(0, nnz + 1, per, dtype=np.uint64)
    return n, m, per, nnz, off, cols.ravel().astype(np.uint64), val


def pinned(a):
    import torch
    t = torch.from_numpy(a).pin_memory()
    return t, t.numpy()


def check(ffi, ctx, n, m, per, off, idx, val, mt):
    o2, i2, v2 = mt.download(values="f32")
    np.testing.assert_array_equal(o2, off)
    np.testing.assert_array_equal(i2, idx)
    np.testing.assert_array_equal(v2.view(np.uint32), val.view(np.uint32))
    np.testing.assert_array_equal(mt.sum(ffi.ROW), val.reshape(n, per).astype(np.float64).sum(axis=1))


@pytest.mark.parametrize("link_gbs", ["0.05", "50", "1000000"])
@pytest.mark.parametrize("memory", ["pageable", "pinned"])
def test_any_mix_of_packed_and_raw_chunks_is_lossless(ffi, ctx, data, link_gbs, memory):
    n, m, per, nnz, off, idx, val = data
    keep = []
    if memory == "pinned":
        (t0, off_), (t1, idx_), (t2, val_) = pinned(off), pinned(idx), pinned(val)
        keep = [t0, t1, t2]
    else:
        off_, idx_, val_ = off, idx, val
    os.environ["SRB_LINK_GBS"] = link_gbs
    try:
        for _ in range(2):
            mt = ffi.DeviceMatrix.upload(ctx, ffi.CSR, n, m, off_, idx_, val_)
            check(ffi, ctx, n, m, per, off, idx, val, mt)
            chunks, ip, vp = ctx.last_upload_chunks()
            h2d, _ = ctx.last_upload()
            assert chunks == 5 and 0 <= vp <= 4 and 1 <= ip <= 5          # chunk 0 is always a packing probe
            if link_gbs == "0.05" or memory == "pageable":
                assert ip == 5                                             # a slow link / pageable memory: every index chunk packed
            if link_gbs == "1000000" and memory == "pinned":
                assert ip <= 2 and vp <= 1                                 # an infinitely fast link: the probe chunk (+ rounding)
                assert h2d >= 8 * (n + 1) + 12 * (nnz - 2 * (1 << 22))
            if link_gbs == "0.05":
                assert vp == 4 and h2d <= 8 * (n + 1) + 2 * nnz + 4 * (nnz - 4 * (1 << 22)) + 64
            mt.free()
    finally:
        del os.environ["SRB_LINK_GBS"]
    del keep


def test_u32_indices_and_untrusted_offsets(ffi, ctx, data):
    """u32 host indices; and offsets that do not end at nnz: the delta coder is skipped (plain narrowing), the upload then
    reports the bad offsets exactly like the other modes."""
    n, m, per, nnz, off, idx, val = data
    mt = ffi.DeviceMatrix.upload(ctx, ffi.CSR, n, m, off.astype(np.uint32), idx.astype(np.uint32), val)
    check(ffi, ctx, n, m, per, off, idx, val, mt)
    mt.free()
    bad = off.copy()
    bad[-1] -= 1
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, n, m, bad, idx, val, nnz=nnz)
    assert e.value.code == -1
    oob = idx.copy()
    oob[5_000_000] = m
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, n, m, off, oob, val)
    assert e.value.code == -3
    unsorted = idx.copy()
    unsorted[[7, 8]] = unsorted[[8, 7]]
    with pytest.raises(ffi.SrbError) as e:
        ffi.DeviceMatrix.upload(ctx, ffi.CSR, n, m, off, unsorted, val)
    assert e.value.code == -8


def test_wide_minor_dimension_and_f64_values(ffi, ctx):
    """nminor > 65 536 (4-byte codes when delta is off, CSC of a tall matrix) and f64 values (never value-packed)."""
    rng = np.random.default_rng(10)
    nmaj, nmin, per = 20_000, 200_000, 250
    cols = np.sort(rng.choice(nmin // per, size=(nmaj, per)) + np.arange(per) * (nmin // per), axis=1)
    off = np.arange(0, nmaj * per + 1, per, dtype=np.uint64)
    idx = cols.ravel().astype(np.uint64)
    val = rng.uniform(0.5, 9.0, nmaj * per)
    mt = ffi.DeviceMatrix.upload(ctx, ffi.CSC, nmin, nmaj, off, idx, val)      # CSC: major = columns
    o2, i2, v2 = mt.download()
    np.testing.assert_array_equal(o2, off)
    np.testing.assert_array_equal(i2, idx)
    np.testing.assert_array_equal(v2, val)
    mt.free()
