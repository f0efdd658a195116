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


# ---- delta coding of the sorted minor indices (HOST_PACK_DELTA) -------------------------------------------------------
def delta_decode(offsets, codes, esc_pos, esc_val):
    """Reference decoder (what the device kernel computes): running sum within a line, restarted at every escape."""
    out = np.zeros(codes.shape[0], np.uint64)
    esc = dict(zip(esc_pos.tolist(), esc_val.tolist()))
    for r in range(offsets.shape[0] - 1):
        prev = 0
        for i in range(int(offsets[r]), int(offsets[r + 1])):
            prev = esc[i] if codes[i] == 255 else prev + int(codes[i])
            out[i] = prev
    return out


def ragged(rng, nrows, ncols, mean_len, empty=()):
    lens = rng.poisson(mean_len, nrows).clip(0, ncols)
    lens[list(empty)] = 0
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    indices = np.concatenate([np.sort(rng.choice(ncols, L, replace=False)) for L in lens] + [np.zeros(0, np.int64)]).astype(np.uint64)
    return offsets, indices


@pytest.mark.parametrize("dtype", [np.uint64, np.uint32])
@pytest.mark.parametrize("chunk", [1, 7, 64, 1000, 1 << 22])
@pytest.mark.parametrize("threads", [1, 0])
def test_delta_round_trip(dtype, chunk, threads):
    rng = np.random.default_rng(chunk)
    off, idx = ragged(rng, 300, 30_000, 40, empty=(0, 5, 6, 299))      # gaps of ~750: most entries are escapes
    off2, idx2 = ragged(rng, 300, 400, 60, empty=(1, 2))               # dense lines: gaps of a few units, no escapes inside
    for o, i, ncols in ((off, idx, 30_000), (off2, idx2, 400)):
        codes, pos, val, oob = _ffi.host_delta_encode(o.astype(dtype), i.astype(dtype), ncols, chunk, threads)
        assert not oob
        assert np.all(np.diff(pos.astype(np.int64)) > 0)                 # positions ascending
        assert np.array_equal(np.nonzero(codes == 255)[0], pos)          # one list entry per escape code
        np.testing.assert_array_equal(delta_decode(o, codes, pos, val), i)
    assert (codes == 255).sum() <= 300                                   # dense case: at most the line starts escape


def delta_decode_fast(offsets, codes, esc_pos, esc_val):
    """The same decoder, vectorised (segments restart at line starts and at escapes) for inputs of many parts."""
    d = codes.astype(np.int64)
    d[esc_pos.astype(np.int64)] = esc_val.astype(np.int64)
    start = np.zeros(codes.shape[0], bool)
    lens = np.diff(offsets.astype(np.int64))
    start[offsets[:-1].astype(np.int64)[lens > 0]] = True
    start[esc_pos.astype(np.int64)] = True
    cs = np.cumsum(d)
    first = np.nonzero(start)[0]
    base = cs[first] - d[first]
    return (cs - base[np.cumsum(start) - 1]).astype(np.uint64)


@pytest.mark.parametrize("dtype", [np.uint64, np.uint32])
@pytest.mark.parametrize("threads", [2, 3, 8])
def test_delta_round_trip_many_parts(dtype, threads):
    """Enough entries that one chunk is cut into several parts (>= 64 K entries each, up to 4 per thread): parts start in the
    middle of a line, the escape lists of the parts are merged in entry order, empty lines sit on part boundaries."""
    rng = np.random.default_rng(threads)
    nrows, ncols = 1500, 30_000
    lens = rng.poisson(450, nrows).clip(0, ncols)
    lens[[0, 1, 700, 701, 702, nrows - 1]] = 0
    lens[10] = 20_000                                                    # one line longer than a part
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    idx = np.concatenate([np.sort(rng.choice(ncols, L, replace=False)) for L in lens]).astype(np.uint64)
    assert idx.shape[0] > 5 * 65536
    small = delta_decode(off[:40], *(_ffi.host_delta_encode(off[:40].astype(dtype), idx[:int(off[39])].astype(dtype), ncols, 1 << 22, 1)[:3]))
    np.testing.assert_array_equal(small, idx[:int(off[39])])             # the fast decoder's reference agrees with the slow one below
    for chunk in (1 << 22, 200_000):
        codes, pos, val, oob = _ffi.host_delta_encode(off.astype(dtype), idx.astype(dtype), ncols, chunk, threads)
        assert not oob
        assert np.all(np.diff(pos.astype(np.int64)) > 0)
        assert np.array_equal(np.nonzero(codes == 255)[0], pos)
        np.testing.assert_array_equal(delta_decode_fast(off, codes, pos, val), idx)
    np.testing.assert_array_equal(delta_decode_fast(off[:40], *(_ffi.host_delta_encode(off[:40].astype(dtype), idx[:int(off[39])].astype(dtype), ncols, 1 << 22, 1)[:3])), small)
    # an out-of-bounds index in the middle of a late part is reported
    bad = idx.copy()
    bad[-70_000] = ncols
    assert _ffi.host_delta_encode(off.astype(dtype), bad.astype(dtype), ncols, 1 << 22, threads)[3] is True


def test_delta_keeps_non_canonical_input_exact():
    """Duplicates (gap 0) and unsorted pairs (negative gap -> escape) must decode to exactly what the caller passed, so
    the device's canonical-form check still reports them."""
    off = np.array([0, 4, 4, 9], np.uint64)
    idx = np.array([3, 3, 900, 2, 0, 254, 509, 510, 100], np.uint64)
    codes, pos, val, oob = _ffi.host_delta_encode(off, idx, 1000, 3)
    assert not oob
    assert codes.tolist() == [3, 0, 255, 255, 0, 254, 255, 1, 255]
    assert pos.tolist() == [2, 3, 6, 8] and val.tolist() == [900, 2, 509, 100]
    np.testing.assert_array_equal(delta_decode(off, codes, pos, val), idx)


def test_delta_bounds_and_bad_offsets():
    off, idx = ragged(np.random.default_rng(2), 50, 500, 20)
    assert _ffi.host_delta_encode(off, idx, 500)[3] is False
    bad = idx.copy()
    bad[17] = 500
    assert _ffi.host_delta_encode(off, bad, 500)[3] is True
    bad[17] = (1 << 63) + 4
    assert _ffi.host_delta_encode(off, bad, 500)[3] is True
    wrong = off.copy()
    wrong[10], wrong[11] = wrong[11], wrong[10] + 0 if wrong[11] != wrong[10] else wrong[10] + 1
    if np.any(np.diff(wrong.astype(np.int64)) < 0):
        with pytest.raises(ValueError, match="-3"):
            _ffi.host_delta_encode(wrong, idx, 500)
    short = off.copy()
    short[-1] -= 1
    with pytest.raises(ValueError, match="-3"):
        _ffi.host_delta_encode(short, idx, 500)


def test_delta_large_multithreaded_matches_single_thread():
    rng = np.random.default_rng(9)
    off, idx = ragged(rng, 4000, 30_000, 1500)                           # 6 M entries, the bench's line length
    a = _ffi.host_delta_encode(off, idx, 30_000, 1 << 22, 1)
    b = _ffi.host_delta_encode(off, idx, 30_000, 1 << 20, 0)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert (a[0] == 255).mean() < 1e-3                                   # escapes are rare at 5 % density
    # spot-check the decode on a slice of lines
    sub = slice(int(off[100]), int(off[140]))
    esc = dict(zip(a[1].tolist(), a[2].tolist()))
    for r in range(100, 140):
        prev = 0
        for i in range(int(off[r]), int(off[r + 1])):
            prev = esc[i] if a[0][i] == 255 else prev + int(a[0][i])
            assert prev == idx[i]
    assert sub.stop > sub.start


def warp_decode_emulation(offsets, codes, esc_pos, esc_val):
    """Lane-by-lane emulation of upload.cu: delta_decode_kernel (32-lane segmented inclusive scan with __shfl_up, an escape
    restarts the running sum, the last lane's result carries into the next 32 entries of the line)."""
    out = np.zeros(codes.shape[0], np.uint64)
    esc = dict(zip(esc_pos.tolist(), esc_val.tolist()))
    for r in range(offsets.shape[0] - 1):
        a, b, carry = int(offsets[r]), int(offsets[r + 1]), 0
        for base in range(a, b, 32):
            v, reset = [0] * 32, [0] * 32
            for lane in range(32):
                i = base + lane
                if i < b:
                    v[lane] = int(codes[i])
                    if v[lane] == 255:
                        v[lane], reset[lane] = esc[i], 1
            o = 1
            while o < 32:
                pv, pr = v[:], reset[:]                      # __shfl_up_sync reads the values before this step
                for lane in range(32):
                    if lane >= o and not reset[lane]:
                        v[lane] += pv[lane - o]
                        reset[lane] = pr[lane - o]
                o <<= 1
            col = [v[lane] if reset[lane] else v[lane] + carry for lane in range(32)]
            for lane in range(32):
                if base + lane < b:
                    out[base + lane] = col[lane]
            carry = col[31]
    return out


def test_device_decode_algorithm_by_emulation():
    rng = np.random.default_rng(4)
    for ncols, mean_len in ((30_000, 45), (400, 70), (30_000, 1)):
        off, idx = ragged(rng, 120, ncols, mean_len, empty=(0, 3, 119))
        codes, pos, val, _ = _ffi.host_delta_encode(off, idx, ncols, 50)
        np.testing.assert_array_equal(warp_decode_emulation(off, codes, pos, val), idx)
    off = np.array([0, 4, 4, 9, 75], np.uint64)                        # non-canonical pairs + a line longer than two warps
    idx = np.concatenate([[3, 3, 900, 2, 0, 254, 509, 510, 100], np.arange(66) * 7]).astype(np.uint64)
    codes, pos, val, _ = _ffi.host_delta_encode(off, idx, 1000, 5)
    np.testing.assert_array_equal(warp_decode_emulation(off, codes, pos, val), idx)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")   # forking a multi-threaded process is the scenario under test
def test_pool_survives_fork():
    """A forked child has the pool object but none of its worker threads: packing there must run serially, not hang."""
    import os
    src = np.random.default_rng(0).integers(0, 30_000, size=3_000_000).astype(np.uint64)
    _ffi.host_pack_indices(src, 2, 30_000)            # creates the pool in the parent
    pid = os.fork()
    if pid == 0:
        try:
            d, oob = _ffi.host_pack_indices(src, 2, 30_000)
            ok = (not oob) and np.array_equal(d, src.astype(np.uint16))
        except BaseException:
            ok = False
        os._exit(0 if ok else 1)
    for _ in range(600):
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            break
        import time
        time.sleep(0.05)
    else:
        os.kill(pid, 9)
        pytest.fail("the forked child hung in the packing pool")
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0


def test_delta_fuzz_against_both_decoders():
    """Random small matrices, including non-canonical ones, through encode -> reference decode and -> warp emulation."""
    from hypothesis import given, settings, strategies as st

    @st.composite
    def lines(draw):
        nlines = draw(st.integers(1, 6))
        ncols = draw(st.sampled_from([3, 40, 300, 70_000]))
        off, idx = [0], []
        for _ in range(nlines):
            n = draw(st.integers(0, 70))
            row = draw(st.lists(st.integers(0, ncols - 1), min_size=n, max_size=n))
            if draw(st.booleans()):
                row = sorted(set(row))                                  # canonical line; otherwise anything goes
            idx += row
            off.append(len(idx))
        return np.array(off, np.uint64), np.array(idx, np.uint64), ncols, draw(st.sampled_from([1, 3, 32, 33, 4096]))

    @settings(max_examples=150, deadline=None)
    @given(lines())
    def run(case):
        off, idx, ncols, chunk = case
        codes, pos, val, oob = _ffi.host_delta_encode(off, idx, ncols, chunk)
        assert not oob
        np.testing.assert_array_equal(delta_decode(off, codes, pos, val), idx)
        np.testing.assert_array_equal(warp_decode_emulation(off, codes, pos, val), idx)

    run()


def test_balanced_upload_rate_model():
    """srb_upload_mix: host and link must finish together. Chunk of 4 M entries, u64 indices -> 1-byte codes, f32 values."""
    from singlerust_b200 import _ffi
    n = 1 << 22
    link = 50.0                                   # GB/s -> bytes per ms = 5e7
    t_k = n * (1 + 4) / (link * 1e6)              # packed indices + raw values
    t_r = n * (8 + 4) / (link * 1e6)
    # a fast host (packing takes half of what the link needs for the packed chunk): nothing goes raw, values get packed
    g, f = _ffi.upload_mix(0.5 * t_k, 0.1, n, link_gbs=link)
    assert g == 0.0 and 0.0 < f <= 1.0
    host, linkt = 0.5 * t_k + f * 0.1, t_k - f * n * 3 / (link * 1e6)
    assert f == 1.0 or abs(host - linkt) < 1e-9   # equalised unless every chunk is already packed
    # a slow host (8 ranks sharing the cores): most chunks go raw, and both sides finish together
    t_idx = 8.0 * t_k
    g, f = _ffi.upload_mix(t_idx, 0.1, n, link_gbs=link)
    assert f == 0.0 and 0.5 < g < 1.0
    assert abs((1 - g) * t_idx - ((1 - g) * t_k + g * t_r)) < 1e-9
    # a slower link (8 GPUs uploading at once) moves the balance back towards packing
    g18, _ = _ffi.upload_mix(t_idx, 0.1, n, link_gbs=18.0)
    assert g18 < g
    # f64 values are never packed; nothing measured yet -> pack (the first chunk is the probe)
    assert _ffi.upload_mix(0.01, 0.1, n, value_bytes=8, link_gbs=link)[1] == 0.0
    assert _ffi.upload_mix(0.0, 0.0, n, link_gbs=link) == (0.0, 0.0)
