"""src/memory/processing/{mod.rs, scale, transform, dim_red} — normalise, log1p, feature selection, PCA."""
from __future__ import annotations

import numpy as np

from .. import _ffi
from ..anndata import IMAnnData
from ..shared import Direction, FeatureSelection


def normalize_total_inplace(adata: IMAnnData, target_sum: float, direction: Direction) -> None:
    """processing/mod.rs:303-312 -> scale::scale_row / scale_col (scale/mod.rs:7-173)."""
    adata.x().normalize_total_inplace(target_sum, int(direction))


def normalize_total(adata: IMAnnData, target_sum: float, direction: Direction) -> IMAnnData:
    """:314-322 — deep_clone + in-place."""
    new = adata.deep_clone()
    normalize_total_inplace(new, target_sum, direction)
    return new


def log1p_transform_inplace(adata: IMAnnData) -> None:
    """:324-326 -> transform::log1p_data (transform/mod.rs:8-62)."""
    adata.x().log1p_inplace()


def log1p_transform(adata: IMAnnData) -> IMAnnData:
    """:328-332."""
    new = adata.deep_clone()
    log1p_transform_inplace(new)
    return new


def select_features(adata: IMAnnData, feature_selection: FeatureSelection) -> np.ndarray:
    """dim_red/mod.rs:123-156."""
    fs = feature_selection
    if fs.kind == "HighlyVariableCol":
        if fs.value not in adata.var:
            raise KeyError(f"Error accessing column '{fs.value}'")
        col = np.asarray(adata.var[fs.value])
        if col.dtype != np.bool_:
            raise TypeError(f"Column '{fs.value}' is not boolean")
        return np.nonzero(col)[0].astype(np.uint64)
    if fs.kind == "HighlyVariable":
        return adata.x().select_hvg(fs.value)
    if fs.kind == "Randomized":
        idx = np.arange(adata.n_vars, dtype=np.uint64)
        np.random.default_rng().shuffle(idx)  # thread_rng in the reference: not reproducible there either
        return idx[:fs.value]
    if fs.kind == "VarianceThreshold":
        return adata.x().select_var_threshold(fs.value)
    if fs.kind == "None":
        return np.arange(adata.n_vars, dtype=np.uint64)
    raise ValueError(fs.kind)


def pca_inplace(adata: IMAnnData, n_components=None, center=None, scale=None, n_threads=None,
                feature_selection: FeatureSelection = FeatureSelection.None_(), svd_mode=None, gram_mode=_ffi.GRAM_TENSOR) -> None:
    """dim_red/mod.rs:24-94. Defaults as the reference: n_components 2 (capped at #features), center/scale True.
    n_threads (rayon pool) and svd_mode (FaerSVD / LapackSVD) are accepted and ignored: the SVD is replaced by the
    equivalent Gram + symmetric eigendecomposition on the device. Stores obsm["X_pca"]; the loadings and the
    explained-variance ratio, which the reference computes and drops (dim_red/mod.rs:77-88), are kept in
    varm["PCA_loadings"] (zero-filled to all genes, attach_pca_results :108-118) and uns-like attribute."""
    sel = select_features(adata, feature_selection)
    if sel.size < 2:
        # dense.column(1) panics in the reference when fewer than two features are selected (dim_red/mod.rs:38-39)
        raise ValueError("pca_inplace needs at least two selected features (the reference panics here)")
    k = min(2 if n_components is None else int(n_components), sel.size)
    res = adata.x().pca(sel, k, True if center is None else bool(center), True if scale is None else bool(scale),
                        gram_mode=gram_mode)
    adata.obsm["X_pca"] = res["scores"]
    full = np.zeros((adata.n_vars, k))
    full[sel.astype(np.int64)] = res["components"]
    adata.varm["PCA_loadings"] = full
    adata.explained_variance_ratio = res["explained_variance_ratio"]
