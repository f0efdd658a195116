"""The K8 algorithm (Chebyshev-filtered subspace iteration, csrc/eig.cu) checked through its NumPy twin
(tools/chfsi_twin.py: same flow, constants and update rules) against LAPACK on five kinds of spectrum. The CUDA code is
compared with cuSOLVER's syevd on the GPU (tests/test_eig_gpu.py); this test pins the numerical method itself."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import chfsi_twin as T  # noqa: E402


@pytest.mark.parametrize("kind", ["flat", "spiked", "lowrank", "clustered", "powerlaw"])
def test_twin_converges_to_lapack(kind):
    d, k = 1024, 24
    rng = np.random.default_rng(7)
    C = T.spectrum(kind, d, rng)
    w, V = np.linalg.eigh(C)
    log = []
    out = T.chfsi_topk(C, k, seed=1, log=log)
    assert out is not None, f"would fall back to syevd: {log}"
    lam, X, st = out
    assert st["outer"] <= 4 and st["max_residual"] <= 1e-11
    np.testing.assert_allclose(lam, w[-k:], rtol=0, atol=1e-11 * abs(w[-1]))
    np.testing.assert_allclose(X.T @ X, np.eye(k), atol=1e-10)
    # eigenvectors: compare the invariant subspace of every well-separated eigenvalue (a degenerate cluster has no
    # unique basis): the projector difference is what parity of the loadings needs
    gaps = np.minimum(np.abs(np.diff(w[-k - 1:]))[:-1], np.abs(np.diff(w[-k - 1:]))[1:]) if k > 1 else None
    sep = np.concatenate([gaps, [abs(w[-1] - w[-2])]]) > 1e-4 * abs(w[-1])
    s = np.sign(np.sum(X * V[:, -k:], axis=0))
    assert np.abs((X * s - V[:, -k:])[:, sep]).max() <= 1e-7 if sep.any() else True
    P1, P2 = X @ X.T, V[:, -k:] @ V[:, -k:].T
    if abs(w[-k] - w[-k - 1]) > 1e-4 * abs(w[-1]):
        assert np.abs(P1 - P2).max() <= 1e-7


def test_twin_declines_what_the_cuda_code_declines():
    rng = np.random.default_rng(0)
    assert T.chfsi_topk(np.eye(300), 10) is None                      # too small: syevd is used
    assert T.chfsi_topk(3.0 * np.eye(1024), 10) is None               # Krylov breakdown (span 0): syevd is used
    C = T.spectrum("flat", 1024, rng)
    C[5, 7] = C[7, 5] = np.nan
    assert T.chfsi_topk(C, 10) is None                                # non-finite: syevd reports the error


@pytest.mark.parametrize("name", ["spectrum_L_1M_f64.bin", "spectrum_2M_f64.bin", "spectrum_4M_f64.bin"])
def test_measured_bench_spectra_converge_in_one_outer_round(name):
    """The spectra of the correlation matrices the bench produces at 1M / 2M / 4M cells (2000 HVGs; dumped on a B200 with
    SRB_EIG_DUMP through the syevd path, tests/golden/): flat Marchenko-Pastur bulks whose top 50 values sit at the edge —
    the hard case for a subspace iteration. The solver must converge in ONE outer round on each of them (a second round
    costs another Rayleigh-Ritz step; with 24 Krylov steps instead of 16 it happened on lucky / unlucky draws), with a
    bounded number of block products, and return the exact leading pairs."""
    ev = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", name))
    assert ev.shape == (2000,) and np.all(np.diff(ev) >= 0)
    rng = np.random.default_rng(5)
    q, _ = np.linalg.qr(rng.standard_normal((2000, 2000)))
    c = (q * ev) @ q.T
    c = (c + c.T) / 2
    for seed in (0, 1):
        w, v, st = T.chfsi_topk(c, 50, seed=seed)
        assert st["outer"] == 1 and st["block_products"] <= 80, st
        np.testing.assert_allclose(np.sort(w), ev[-50:], rtol=1e-12)
        assert np.max(np.abs(np.abs(np.sum(q[:, -50:] * v[:, np.argsort(w)], axis=0)) - 1.0)) < 1e-9
