import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
import subprocess
def mem():
    return subprocess.run(["nvidia-smi","--query-gpu=memory.used","--format=csv,noheader"],capture_output=True,text=True).stdout.strip()
cells = int(os.environ.get("CELLS", "1000000"))
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, cells, 30000, thr, amp)
for mode in ("pipeline_no_outputs", "pipeline_outputs_noscores", "separate"):
    for it in range(5):
        ctx.synchronize(); t0 = time.perf_counter()
        w = mat.clone()
        if mode == "pipeline_no_outputs":
            w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_outputs=False)
        elif mode == "pipeline_outputs_noscores":
            w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_scores=False)
        else:
            w.normalize_total_inplace(1e4, 0); w.log1p_inplace(); sel = w.select_hvg(2000); w.pca(sel, 50, want_scores=False)
        t1 = time.perf_counter()
        st = ctx.last_stage_ms(); t2 = time.perf_counter()
        w.free(); ctx.synchronize(); t3 = time.perf_counter()
        print(mode, it, f"call {1e3*(t1-t0):.1f} stage_query {1e3*(t2-t1):.1f} free {1e3*(t3-t2):.1f} ms  sum_stages {sum(st.values()):.1f}", mem(), flush=True)
