"""K8 probe at the bench size: (1) full spectrum of the bench correlation matrix (syevd, SRB_EIG_DUMP) for offline tuning of
the ChFSI twin, (2) wall-clock per ChFSI phase (SRB_EIG_TRACE=1, synchronising), (3) untraced eig stage time."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("CELLS", "1000000"))
os.makedirs("gpurun_out", exist_ok=True)
ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(30000, seed=0x5EED0002, mean_density=0.05)
mat = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, n, 30000, thr, amp)


def run(mode, reps=3):
    ctx.set_eig_mode(mode)
    ms = []
    for _ in range(reps):
        w = mat.clone()
        w.pipeline_normalize_hvg_pca(1e4, 2000, 50, want_outputs=False)
        ms.append(ctx.last_stage_ms()["eig"])
        w.free()
    return ms


if os.environ.get("DUMP", "1") == "1":
    os.environ["SRB_EIG_DUMP"] = "gpurun_out/bench_spectrum_f64.bin"
    run(_ffi.EIG_SYEVD, 1)
    del os.environ["SRB_EIG_DUMP"]
    ev = np.fromfile("gpurun_out/bench_spectrum_f64.bin")
    print("spectrum: n =", ev.size, "min %.6g max %.6g" % (ev.min(), ev.max()), "top5", ev[-5:][::-1], file=sys.stderr)
out = {"cells": n, "chfsi_eig_ms": run(_ffi.EIG_CHFSI), "info": ctx.last_eig()}
os.environ["SRB_EIG_TRACE"] = "1"
run(_ffi.EIG_CHFSI, 2)
del os.environ["SRB_EIG_TRACE"]
print(json.dumps(out))
