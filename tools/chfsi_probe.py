"""K8 A/B at the bench size: cuSOLVER syevd vs the ChFSI top-k solver (srb_ctx_set_eig_mode). Same matrix, same pipeline;
compares explained-variance ratios, loadings and scores and reports the eig stage time of both."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("CELLS", "1000000"))
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, n, 30000, thr, amp)
out = {"cells": n}


def run(mode, reps=3):
    ctx.set_eig_mode(mode)
    w = mat.clone()
    r = w.pipeline_normalize_hvg_pca(1e4, 2000, 50)
    info = ctx.last_eig()
    w.free()
    ms, tot = [], []
    for _ in range(reps):
        w = mat.clone()
        w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_outputs=False)
        st = ctx.last_stage_ms()
        ms.append(st["eig"])
        tot.append(sum(st.values()))
        w.free()
    return r, info, ms, tot


ra, ia, ma, ta = run(_ffi.EIG_SYEVD)
rb, ib, mb, tb = run(_ffi.EIG_CHFSI)
out["syevd"] = {"eig_ms": [round(x, 2) for x in ma], "stages_total_ms": [round(x, 2) for x in ta], "info": ia}
out["chfsi"] = {"eig_ms": [round(x, 2) for x in mb], "stages_total_ms": [round(x, 2) for x in tb], "info": ib}
assert np.array_equal(ra["selection"], rb["selection"])
out["evr_max_rel_diff"] = float(np.max(np.abs(ra["explained_variance_ratio"] - rb["explained_variance_ratio"]) / ra["explained_variance_ratio"]))
s = np.sign(np.sum(ra["components"] * rb["components"], axis=0))
out["components_max_abs_diff"] = float(np.max(np.abs(ra["components"] - rb["components"] * s)))
out["components_sign_flips"] = int((s < 0).sum())
out["scores_max_rel_diff"] = float(np.max(np.abs(ra["scores"] - rb["scores"] * s)) / np.max(np.abs(ra["scores"])))
V = rb["components"]
out["chfsi_orthonormality"] = float(np.max(np.abs(V.T @ V - np.eye(V.shape[1]))))
print(json.dumps(out))
