"""Out-of-core pipeline (srb_gene_moments + srb_pca_stream_*; BASELINE.json config 5 in miniature): three passes over the row
chunks of an on-disk CSR store with one chunk resident at a time must reproduce the resident pipeline — the HVG list bit
for bit, explained-variance ratios and scores within the Gram path's accuracy — and the SVD oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pca_oracle as P
from tests._util import sign_align

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from singlerust_b200 import _ffi, backed, memory
    from singlerust_b200.anndata import BackedAnnData, IMAnnData
    from singlerust_b200.shared import ComputationMode, Direction, FeatureSelection
    ctx = _ffi.Context(0)
    yield dict(ffi=_ffi, ctx=ctx, backed=backed, memory=memory, BackedAnnData=BackedAnnData, IMAnnData=IMAnnData, CM=ComputationMode,
               D=Direction, FS=FeatureSelection)
    ctx.close()


@pytest.mark.parametrize("gram_mode", [1, 0])
def test_out_of_core_equals_resident(env, tmp_path, gram_mode):
    from tests.test_gpu_parity import clustered_counts
    ffi, ctx, mem = env["ffi"], env["ctx"], env["memory"]
    a = clustered_counts(np.random.default_rng(51), 4000, 600)
    env["BackedAnnData"].write_store(str(tmp_path), a)
    disk = env["BackedAnnData"].open_store(str(tmp_path))
    n_top, k = 120, 4
    ref = env["IMAnnData"].from_scipy(ctx, a)
    mem.processing.normalize_total_inplace(ref, 1e4, env["D"].Row)
    mem.processing.log1p_transform_inplace(ref)
    sel_ref = mem.processing.select_features(ref, env["FS"].HighlyVariable(n_top))
    mem.processing.pca_inplace(ref, k, True, True, None, env["FS"].HighlyVariable(n_top), gram_mode=gram_mode)
    tol = 1e-7 if gram_mode == 1 else 1e-4          # fp64 Gram: summation order + 1e-9 moments; tensor-core Gram: 1e-6-level entries
    for chunk in (900, 4000, 97):
        r = env["backed"].processing.normalize_hvg_pca_out_of_core(ctx, disk, env["CM"].Chunked(chunk), 1e4, n_top, k, gram_mode=gram_mode)
        np.testing.assert_array_equal(r["selection"], sel_ref)
        np.testing.assert_allclose(r["explained_variance_ratio"], ref.explained_variance_ratio, rtol=tol)
        want = ref.obsm["X_pca"]
        np.testing.assert_allclose(sign_align(r["scores"], want), want, rtol=0, atol=tol * np.abs(want).max())
    if gram_mode == 1:                               # and against the SVD oracle on the transformed matrix
        ol = O.log1p(O.normalize_total(O.Compressed.from_scipy(a), 1e4, O.ROW))
        w = P.pca_pipeline(ol, n_top, k, selection=r["selection"])
        np.testing.assert_allclose(r["explained_variance_ratio"], w["explained_variance_ratio"], rtol=1e-5)
        np.testing.assert_allclose(sign_align(r["scores"], w["scores"]), w["scores"], rtol=0, atol=1e-5 * np.abs(w["scores"]).max())


def test_gene_moments_of_a_chunk(env):
    ffi, ctx = env["ffi"], env["ctx"]
    from tests._util import random_csr
    a = random_csr(np.random.default_rng(52), 800, 150, 0.1, empty_cols=(4,))
    m = ffi.DeviceMatrix.from_scipy(ctx, a)
    cnt, s, q = m.gene_moments()
    d = a.toarray().astype(np.float64)
    np.testing.assert_array_equal(cnt, np.asarray((a != 0).sum(axis=0)).ravel())
    np.testing.assert_array_equal(s, d.sum(axis=0))                     # integer counts: exact
    np.testing.assert_array_equal(q, (d * d).sum(axis=0))
    m.normalize_total_inplace(1e4, ffi.ROW)
    m.log1p_inplace()
    cnt2, s2, q2 = m.gene_moments()                                      # pending transforms are applied first
    ol = O.log1p(O.normalize_total(O.Compressed.from_scipy(a), 1e4, O.ROW))
    np.testing.assert_array_equal(cnt2, cnt)
    np.testing.assert_allclose(s2, O.sum_(ol, O.COLUMN), rtol=1e-6)
    np.testing.assert_allclose(q2, O.sum_(O.Compressed("csr", 800, 150, ol.offsets, ol.indices, ol.values ** 2), O.COLUMN), rtol=1e-6)


def test_cpp_out_of_core():
    from tests.test_zz_cpp_host import build_exe
    r = subprocess.run([build_exe(), "--out-of-core"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "0 failed" in r.stdout
