"""Print the per-stage CUDA-event times of a few pipeline steps (environment variables select kernel variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 1000000, 30000, thr, amp)
out = []
for it in range(4):
    w = mat.clone()
    r = w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_scores=False)
    out.append(ctx.last_stage_ms())
    w.free()
tag = {k: v for k, v in os.environ.items() if k.startswith("SRB_")}
print(tag, {k: [round(o[k], 2) for o in out[1:]] for k in ("row_sums", "fused_norm_log1p_moments", "densify", "gram")}, "evr0", r["explained_variance_ratio"][0], "hvg0", r["selection"][:3])
