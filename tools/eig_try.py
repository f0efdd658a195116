import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 200000, 30000, thr, amp)
w = mat.clone(); w.normalize_total_inplace(1e4, 0); w.log1p_inplace(); sel = w.select_hvg(2000)
e = []
for it in range(4):
    r = w.pca(sel, 50)
    e.append(ctx.last_stage_ms()["eig"])
tag = {k: v for k, v in os.environ.items() if k.startswith("SRB_")}
np.save("gpurun_out/evr_%s.npy" % os.environ.get("SRB_EIG_RANGE", "0"), r["explained_variance_ratio"])
np.save("gpurun_out/comp_%s.npy" % os.environ.get("SRB_EIG_RANGE", "0"), r["components"])
print(tag, "eig ms", [round(x, 2) for x in e], "evr", r["explained_variance_ratio"][[0, 24, 49]])
