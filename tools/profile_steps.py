"""Wall-clock breakdown of one bench step (host view) — where does time go outside the kernels?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from singlerust_b200 import _ffi, synth

cells = int(os.environ.get("CELLS", "1000000"))
gram = int(os.environ.get("SRB_GRAM_MODE", "1"))
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
t = time.perf_counter(); mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, cells, 30000, thr, amp); ctx.synchronize()
print("synth s", time.perf_counter() - t, mat.info())
for it in range(4):
    ctx.synchronize(); t0 = time.perf_counter()
    w = mat.clone(); t1 = time.perf_counter()
    w.normalize_total_inplace(1e4, 0); ctx.synchronize(); t2 = time.perf_counter()
    w.log1p_inplace(); ctx.synchronize(); t3 = time.perf_counter()
    sel = w.select_hvg(2000); ctx.synchronize(); t4 = time.perf_counter()
    r = w.pca(sel, 50, gram_mode=gram, want_scores=False); ctx.synchronize(); t5 = time.perf_counter()
    st = ctx.last_stage_ms()
    w.free(); ctx.synchronize(); t6 = time.perf_counter()
    print(f"it{it}: clone {1e3*(t1-t0):.2f} norm {1e3*(t2-t1):.2f} log1p {1e3*(t3-t2):.2f} hvg {1e3*(t4-t3):.2f} pca {1e3*(t5-t4):.2f} free {1e3*(t6-t5):.2f} total {1e3*(t6-t0):.2f} ms | stages", {k: round(v, 2) for k, v in st.items()})
