"""Print the per-stage CUDA-event times of a few pipeline steps (environment variables select kernel variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 1000000, 30000, thr, amp)
w = mat.clone(); w.normalize_total_inplace(1e4, 0); w.log1p_inplace(); sel = w.select_hvg(2000)
out = []
for it in range(4):
    r = w.pca(sel, 50, want_scores=False)
    out.append(ctx.last_stage_ms())
tag = {k: v for k, v in os.environ.items() if k.startswith("SRB_")}
print(tag, {k: [round(o[k], 2) for o in out[1:]] for k in ("densify", "gram", "scores")}, "evr0", r["explained_variance_ratio"][0])
