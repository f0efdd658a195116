import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, 1000000, 30000, thr, amp)
def steps(tag, n=5):
    out = []
    for it in range(n):
        ctx.synchronize(); t0 = time.perf_counter()
        w = mat.clone(); w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_outputs=False); st = ctx.last_stage_ms(); w.free()
        out.append(1e3 * (time.perf_counter() - t0))
    print(tag, ["%.0f" % x for x in out], flush=True)
steps("plain")
import torch
steps("after import torch")
torch.cuda.set_device(0); torch.cuda.synchronize()
steps("after torch cuda init")
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
steps("after event record")
e1.record(stream); torch.cuda.synchronize(); print("elapsed", e0.elapsed_time(e1))
x = torch.zeros(10, device="cuda"); torch.cuda.synchronize()
steps("after torch alloc")
