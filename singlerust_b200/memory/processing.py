"""src/memory/processing/{mod.rs, scale, transform, dim_red} — normalise, log1p, feature selection, PCA."""
from __future__ import annotations

import numpy as np

from .. import _ffi
from ..anndata import IMAnnData
from ..shared import Direction, FeatureSelection


# ---- filters: memory/processing/mod.rs:16-299 ---------------------------------------------------------
F64_MIN, F64_MAX = -np.finfo(np.float64).max, np.finfo(np.float64).max


def linear_quantile(values, q: float) -> float:
    """ndarray_stats::interpolate::Linear on the sorted values: index q (n - 1); lower + (higher - lower) * frac.
    (NumPy's 'linear' method evaluates the same point with a differently rounded lerp for frac >= 0.5.)"""
    v = np.sort(np.asarray(values, dtype=np.float64))
    if v.size == 0:
        raise ValueError("Error calculating percentile: empty input")
    if not 0.0 <= q <= 1.0:
        raise ValueError("Error calculating percentile: q outside [0, 1]")
    pos = q * (v.size - 1)
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    return float(v[lo] + (v[hi] - v[lo]) * (pos - lo))


def calculate_percentiles(values, lower_lim, upper_lim):
    """processing/mod.rs:148-174: ndarray-stats quantile with Linear interpolation; f64::MIN / f64::MAX when the limit
    is not Relative."""
    v = np.asarray(values, dtype=np.float64)
    if np.isnan(v).any():
        raise ValueError("NaN in the per-line sums (noisy_float n64 panics in the reference)")
    lo = linear_quantile(v, lower_lim.value) if lower_lim.is_relative() else F64_MIN
    hi = linear_quantile(v, upper_lim.value) if upper_lim.is_relative() else F64_MAX
    return lo, hi


def create_filter_mask(n, counts, sums, lower_lim, upper_lim, lower_percentile, upper_percentile):
    """processing/mod.rs:32-83 (cells) and :193-243 (genes): the nine (lower, upper) FlexValue combinations.
    Absolute limits compare the stored-entry COUNT, Relative limits compare the SUM against its percentile."""
    keep = np.ones(n, dtype=bool)
    if lower_lim.is_absolute():
        keep &= counts >= lower_lim.value
    elif lower_lim.is_relative():
        keep &= sums >= lower_percentile
    if upper_lim.is_absolute():
        keep &= counts <= upper_lim.value
    elif upper_lim.is_relative():
        keep &= sums <= upper_percentile
    return keep


def _filter(adata: IMAnnData, lower_lim, upper_lim, direction: Direction, inplace: bool):
    need_count = lower_lim.is_absolute() or upper_lim.is_absolute()
    counts = adata.x().number(int(direction)) if need_count else None        # calculate_{cell,gene}_stats :16-30, :176-191
    sums = adata.x().sum(int(direction))
    lo, hi = calculate_percentiles(sums, lower_lim, upper_lim)
    n = adata.n_obs if direction == Direction.Row else adata.n_vars
    mask = create_filter_mask(n, counts, sums, lower_lim, upper_lim, lo, hi)
    new_x = adata.x().subset(mask, None) if direction == Direction.Row else adata.x().subset(None, mask)
    side = "obs" if direction == Direction.Row else "var"
    new_cols = {k: np.asarray(v)[mask] for k, v in getattr(adata, side).items()}
    if inplace:
        adata._x = new_x
        setattr(adata, side, new_cols)
        if direction == Direction.Row:
            adata.obsm = {k: np.asarray(v)[mask] for k, v in adata.obsm.items()}
        else:
            adata.varm = {k: np.asarray(v)[mask] for k, v in adata.varm.items()}
        return None
    out = IMAnnData(new_x, new_cols if side == "obs" else dict(adata.obs), new_cols if side == "var" else dict(adata.var))
    return out


def filter_cells_inplace(adata: IMAnnData, lower_lim, upper_lim) -> None:
    """processing/mod.rs:86-121."""
    _filter(adata, lower_lim, upper_lim, Direction.Row, True)


def filter_cells(adata: IMAnnData, lower_lim, upper_lim) -> IMAnnData:
    """processing/mod.rs:123-146."""
    return _filter(adata, lower_lim, upper_lim, Direction.Row, False)


def filter_genes_inplace(adata: IMAnnData, lower_lim, upper_lim) -> None:
    """processing/mod.rs:245-271."""
    _filter(adata, lower_lim, upper_lim, Direction.Column, True)


def filter_genes(adata: IMAnnData, lower_lim, upper_lim) -> IMAnnData:
    """processing/mod.rs:273-299."""
    return _filter(adata, lower_lim, upper_lim, Direction.Column, False)


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
