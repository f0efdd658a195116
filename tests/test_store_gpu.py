"""Backed data from an on-disk chunk store (BackedAnnData.write_store / open_store): statistics and the full pipeline
streamed chunk by chunk from memory-mapped files equal the in-memory results."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests._util import random_csr, sign_align

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_store_streams_like_memory(tmp_path, fmt):
    from singlerust_b200 import _ffi, backed, memory
    from singlerust_b200.anndata import BackedAnnData, IMAnnData
    from singlerust_b200.shared import ComputationMode, Direction, FeatureSelection
    rng = np.random.default_rng(41)
    from tests.test_gpu_parity import clustered_counts
    a = clustered_counts(rng, 3000, 400)   # cell programmes over noise: well separated leading components
    if fmt == "csc":
        a = a.tocsc()
        a.sort_indices()
    BackedAnnData.write_store(str(tmp_path), a, np.uint32 if fmt == "csc" else np.uint64)
    disk = BackedAnnData.open_store(str(tmp_path))
    ctx = _ffi.Context(0)
    try:
        o = O.Compressed.from_scipy(a)
        for mode in (ComputationMode.Whole(), ComputationMode.Chunked(700), ComputationMode.Chunked(4096)):
            for d in (Direction.Row, Direction.Column):
                np.testing.assert_array_equal(backed.statistics.compute_number(ctx, disk, d, mode), O.number(o, int(d)))
                np.testing.assert_array_equal(backed.statistics.compute_sum(ctx, disk, d, mode), O.sum_(o, int(d)))
        if fmt == "csr":
            dev = backed.processing.normalize_hvg_pca(ctx, disk, ComputationMode.Chunked(700), 1e4, 64, 5)
            ref = IMAnnData.from_scipy(ctx, a)
            memory.processing.normalize_total_inplace(ref, 1e4, Direction.Row)
            memory.processing.log1p_transform_inplace(ref)
            memory.processing.pca_inplace(ref, 5, True, True, None, FeatureSelection.HighlyVariable(64))
            np.testing.assert_allclose(dev.explained_variance_ratio, ref.explained_variance_ratio, rtol=1e-9)
            np.testing.assert_allclose(sign_align(dev.obsm["X_pca"], ref.obsm["X_pca"]), ref.obsm["X_pca"], atol=1e-6)
    finally:
        ctx.close()
