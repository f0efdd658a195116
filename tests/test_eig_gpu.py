"""K8: the Chebyshev-filtered subspace iteration (default for n_sel >= 1024, 8 k <= n_sel) against cuSOLVER's syevd on the
same correlation matrix — explained-variance ratios, loadings and scores must agree far inside the 1e-5 parity
tolerance, whichever way the spectrum looks (flat noise, or a few strong cell programmes on top of noise). If the
iteration declines (breakdown / no convergence) the library falls back to syevd by itself; the results must match
either way."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as O
from oracle import pca_oracle as P
from tests._util import sign_align

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ffi():
    from singlerust_b200 import _ffi
    return _ffi


@pytest.fixture(scope="module")
def ctx(ffi):
    c = ffi.Context(0, value_mode=ffi.VALUES_FAITHFUL)
    yield c
    c.close()


def both_solvers(ffi, ctx, m, sel, k):
    out = {}
    for name, mode in (("syevd", ffi.EIG_SYEVD), ("chfsi", ffi.EIG_CHFSI)):
        ctx.set_eig_mode(mode)
        out[name] = m.pca(sel, k, gram_mode=ffi.GRAM_FP64)
        out[name + "_info"] = ctx.last_eig()
    ctx.set_eig_mode(ffi.EIG_CHFSI)
    return out


def check_same(a, b, k):
    np.testing.assert_allclose(b["explained_variance_ratio"], a["explained_variance_ratio"], rtol=1e-9)
    # two converged fp64 solvers agree to ~1e-11 here; the asserted bound is the parity tolerance with a margin of 10
    np.testing.assert_allclose(sign_align(b["components"], a["components"]), a["components"], atol=1e-6)
    np.testing.assert_allclose(sign_align(b["scores"], a["scores"]), a["scores"], rtol=0, atol=1e-6 * np.abs(a["scores"]).max())
    V = b["components"]
    np.testing.assert_allclose(V.T @ V, np.eye(k), atol=1e-9)


def test_flat_spectrum_synthetic(ffi, ctx):
    from singlerust_b200 import synth
    n, mg, d, k = 30_000, 8_000, 1_200, 20
    thr, amp = synth.gene_tables(mg, seed=11, mean_density=0.05)
    m = ffi.DeviceMatrix.synth(ctx, 0x5EED0021, n, mg, thr, amp)
    m.normalize_total_inplace(1e4, ffi.ROW)
    m.log1p_inplace()
    sel = m.select_hvg(d)
    r = both_solvers(ffi, ctx, m, sel, k)
    assert r["syevd_info"]["solver"] == "syevd"
    assert r["chfsi_info"]["solver"] in ("chfsi", "chfsi->syevd")
    check_same(r["syevd"], r["chfsi"], k)


def test_structured_spectrum_against_the_oracle(ffi, ctx):
    """Five strong cell programmes over Poisson noise: a spiked spectrum, like real data."""
    rng = np.random.default_rng(12)
    n, d, k = 4_000, 1_500, 10
    load = rng.gamma(2.0, 1.0, size=(5, d)) * (rng.random((5, d)) < 0.15)
    act = rng.gamma(1.0, 1.0, size=(n, 5))
    lam = 0.15 + act @ load * 0.4
    a = sp.csr_matrix(rng.poisson(lam).astype(np.float32))
    a.sort_indices()
    keep = np.asarray((a != 0).sum(axis=0)).ravel() > 1          # drop never / once expressed genes (zero variance)
    a = sp.csr_matrix(a[:, keep])
    d = a.shape[1]
    assert d >= 1024
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    sel = np.arange(d, dtype=np.uint64)
    r = both_solvers(ffi, ctx, m, sel, k)
    check_same(r["syevd"], r["chfsi"], k)
    want = P.pca_fit_transform(O.densify_selected(O.Compressed.from_scipy(a), np.arange(n), sel), k, True, True)
    np.testing.assert_allclose(r["chfsi"]["explained_variance_ratio"], want["explained_variance_ratio"], rtol=1e-8)
    np.testing.assert_allclose(sign_align(r["chfsi"]["components"], want["components"]), want["components"], atol=1e-6)
