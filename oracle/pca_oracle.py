"""NumPy ORACLE of the PCA stage — TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).

The reference delegates PCA arithmetic to the un-vendored crate `single_algebra = "0.1.0-alpha.3"`
(/root/reference/Cargo.toml:42; call sites src/memory/processing/dim_red/mod.rs:53-69). Its published
behaviour is the classical centred/scaled SVD PCA, and the only in-tree statement of the intended
semantics is the dead-code module src/shared/processing/pca/mod.rs:74-185, restated here:

    mean = X.mean(axis=0); std = X.std(axis=0, ddof=0) if scale else 1           (:87-96)
    Z = X;  if center: Z -= mean;  if scale: Z /= std                            (:98-111)
    U,S,Vt = svd(Z)                                                              (:124)
    eigenvalues = S**2/(n-1); ratio = eigenvalues/eigenvalues.sum()              (:131-134)
    components = V[:, :k]; explained_variance_ratio = ratio[:k]                  (:145-151)
    transform(X) = ((X - mean)/std) @ components                                 (:156-185)
    loadings = components.T * std                                                (:204-215)

PARITY UNPINNED: no reference test asserts a PCA number (tests/test_basic_load.rs:61-235 only log).
Singular vectors are defined up to sign; compare after sign alignment (tests/_util.py).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


def pca_fit_transform(dense: np.ndarray, n_components: int, center: bool = True, scale: bool = True):
    """Returns dict(scores n×k, components d×k, explained_variance_ratio k, mean d, std d, eigenvalues)."""
    X = np.asarray(dense, dtype=np.float64)
    n, d = X.shape
    k = min(n_components, d)
    if center or scale:
        mean = X.mean(axis=0)
        std = X.std(axis=0, ddof=0) if scale else np.ones(d)
    else:
        mean, std = np.zeros(d), np.ones(d)
    Z = X.copy()
    if center:
        Z -= mean
    if scale:
        Z /= std
    _, S, Vt = np.linalg.svd(Z, full_matrices=False)
    eig = S * S / (n - 1)
    ratio = eig / eig.sum()
    comps = Vt.T[:, :k].copy()
    scores = Z @ comps
    return dict(scores=scores, components=comps, explained_variance_ratio=ratio[:k].copy(), mean=mean, std=std,
                eigenvalues=eig)


def pca_pipeline(m: O.Compressed, n_top: int, n_components: int = 2, center: bool = True, scale: bool = True,
                 selection=None):
    """pca_inplace (dim_red/mod.rs:24-94): select_features(HighlyVariable(n_top)) -> selected densify with
    columns in SELECTION order -> PCA fit/transform. `selection` overrides the HVG list (used to feed both
    sides the same list, SURVEY §7 hard part 3)."""
    if selection is None:
        selection = O.select_hvg(O.variance(m, O.COLUMN), n_top)
    rows = np.arange(m.nrows, dtype=np.uint64)
    dense = O.densify_selected(m, rows, selection)
    k = min(n_components, len(selection))  # dim_red/mod.rs:52
    res = pca_fit_transform(dense, k, center, scale)
    res["selection"] = np.asarray(selection, dtype=np.uint64)
    return res
