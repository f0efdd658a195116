"""The K8 algorithm (Chebyshev-filtered subspace iteration, csrc/eig.cu) checked through its NumPy twin
(tools/chfsi_twin.py: same flow, constants and update rules) against LAPACK on five kinds of spectrum. The CUDA code is
compared with cuSOLVER's syevd on the GPU (tests/test_zzz1_eig_gpu.py); this test pins the numerical method itself."""
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
