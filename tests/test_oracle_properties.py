"""Property tests of the C oracle (oracle/srb_oracle.c) against a second, pure-Python loop statement of the same
reference lines, on small random matrices that DO hold explicitly stored zeros, empty lines and singleton lines —
the cases where "stored entry" and "non-zero entry" differ (helper/csr.rs:21-36, 158-186, 200-220)."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as O


@st.composite
def compressed(draw):
    nmajor = draw(st.integers(1, 7))
    nminor = draw(st.integers(1, 6))
    fmt = draw(st.sampled_from(["csr", "csc"]))
    dtype = draw(st.sampled_from([np.float32, np.float64]))
    offsets, indices, values = [0], [], []
    for _ in range(nmajor):
        cols = sorted(draw(st.sets(st.integers(0, nminor - 1), max_size=nminor)))
        for c in cols:
            indices.append(c)
            values.append(draw(st.sampled_from([0.0, 0.0, 1.0, 2.0, 3.0, 7.0, 0.5, 100.0])))
        offsets.append(len(indices))
    nrows, ncols = (nmajor, nminor) if fmt == "csr" else (nminor, nmajor)
    return O.Compressed(fmt, nrows, ncols, np.array(offsets, np.uint64), np.array(indices, np.uint64), np.array(values, dtype))


def loops(m, direction):
    """number / sum / variance / min / max by plain loops, following csr.rs line by line (csc.rs is the mirror)."""
    major = bool(m.along_major(direction))
    n = m.out_len(direction)
    off, idx, val = m.offsets, m.indices, m.values.astype(np.float64)
    cnt, s, sq = [0] * n, [0.0] * n, [0.0] * n
    mn, mx = [math.inf] * n, [-math.inf] * n
    for i in range(len(off) - 1):
        for k in range(int(off[i]), int(off[i + 1])):
            t = i if major else int(idx[k])
            cnt[t] += 1
            s[t] += val[k]
            sq[t] += val[k] * val[k]
            mn[t], mx[t] = min(mn[t], val[k]), max(mx[t], val[k])
    var = []
    if major:  # two-pass, empty -> NaN (csr.rs:158-170)
        for i in range(n):
            if cnt[i] == 0:
                var.append(math.nan)
                continue
            mean, acc = s[i] / cnt[i], 0.0
            for k in range(int(off[i]), int(off[i + 1])):
                acc += (val[k] - mean) ** 2
            var.append(acc / cnt[i])
    else:      # one-pass, empty -> 0 (csr.rs:172-186)
        for j in range(n):
            var.append(sq[j] / cnt[j] - (s[j] / cnt[j]) ** 2 if cnt[j] > 0 else 0.0)
    return cnt, s, var, mn, mx


@settings(max_examples=300, deadline=None)
@given(compressed())
def test_oracle_equals_loop_statement(m):
    for direction in (O.ROW, O.COLUMN):
        cnt, s, var, mn, mx = loops(m, direction)
        assert O.number(m, direction).tolist() == cnt
        np.testing.assert_array_equal(O.sum_(m, direction), np.array(s))
        got = O.variance(m, direction)
        np.testing.assert_allclose(got, np.array(var), rtol=1e-12, atol=1e-12, equal_nan=True)
        sd = O.std_dev(m, direction)
        with np.errstate(invalid="ignore"):
            np.testing.assert_array_equal(sd, np.sqrt(got))
        omn, omx = O.min_max(m, direction)
        np.testing.assert_array_equal(omn, np.array(mn))
        np.testing.assert_array_equal(omx, np.array(mx))


@settings(max_examples=200, deadline=None)
@given(compressed(), st.sampled_from([1.0, 10.0, 1e4]))
def test_normalize_log1p_properties(m, target):
    for direction in (O.ROW, O.COLUMN):
        nm = O.normalize_total(m, target, direction)
        assert nm.values.dtype == np.float64                       # scale/mod.rs:82: always f64 afterwards
        sums0, sums = O.sum_(m, direction), O.sum_(nm, direction)
        for i in range(len(sums)):
            if sums0[i] == 0.0:
                assert sums[i] == 0.0                               # scale 0 for an empty or all-zero line (scale/mod.rs:9-15)
            else:
                assert abs(sums[i] - target) <= 1e-9 * target
        assert np.all(nm.values[m.values == 0] == 0)
        lg = O.log1p(nm)
        np.testing.assert_allclose(lg.values, np.log1p(nm.values), rtol=1e-15)
        # structure untouched
        np.testing.assert_array_equal(lg.offsets, m.offsets)
        np.testing.assert_array_equal(lg.indices, m.indices)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.sampled_from([0.0, 0.0, 1.0, 1.0, 2.5, 7.0, -0.0]), min_size=1, max_size=12), st.integers(0, 14))
def test_hvg_order_is_a_stable_descending_sort(vs, n_top):
    v = np.array(vs)
    got = O.select_hvg(v, n_top).tolist()
    want = sorted(range(len(vs)), key=lambda i: -vs[i])[:n_top]    # Python's sort is stable: ties keep ascending index
    assert got == want
