"""One pipeline step at bench size, for ncu captures (a number printed under ncu is never a bench value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth

cells = int(os.environ.get("CELLS", "1000000"))
steps = int(os.environ.get("STEPS", "1"))
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, cells, 30000, thr, amp)
for _ in range(steps):
    w = mat.clone()
    w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_outputs=False)
    w.free()
print("stages", ctx.last_stage_ms())
