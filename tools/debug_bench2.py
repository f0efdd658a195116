import os, sys, time
import torch
torch.cuda.set_device(0); torch.cuda.synchronize()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 1000000, 30000, thr, amp)
def steps(tag, n=8):
    out = []
    for it in range(n):
        ctx.synchronize(); t0 = time.perf_counter()
        w = mat.clone(); w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_outputs=False); st = ctx.last_stage_ms(); w.free()
        out.append("%.0f(eig %.0f)" % (1e3 * (time.perf_counter() - t0), st["eig"]))
    print(tag, out, flush=True)
steps("torch first")
os.system("grep -E 'cusolver|cublas|libnccl|cudart' /proc/%d/maps | awk '{print $6}' | sort -u" % os.getpid())
