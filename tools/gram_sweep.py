"""Time the Gram stage for the chunk sizes given in SRB_GRAM_CHUNK (one process per value; env is read once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 1000000, 30000, thr, amp)
w = mat.clone(); w.normalize_total_inplace(1e4, 0); w.log1p_inplace(); sel = w.select_hvg(2000)
g = []
for it in range(4):
    r = w.pca(sel, 50, want_scores=False)
    g.append(ctx.last_stage_ms()["gram"])
print("chunk", os.environ.get("SRB_GRAM_CHUNK", "2"), "pair", os.environ.get("SRB_GRAM_PAIR", "1"), "gram ms", [round(x, 2) for x in g], "evr0", r["explained_variance_ratio"][0])
