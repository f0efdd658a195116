import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 200000, 30000, thr, amp)
w = mat.clone(); w.normalize_total_inplace(1e4, 0); w.log1p_inplace(); sel = w.select_hvg(2000)
e = []
for it in range(4):
    r = w.pca(sel, 50, want_scores=False)
    e.append(ctx.last_stage_ms()["eig"])
print("SRB_EIG_X", os.environ.get("SRB_EIG_X", "0"), "eig ms", [round(x, 2) for x in e], "evr0", r["explained_variance_ratio"][0])
