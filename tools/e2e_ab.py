"""A/B of the two upload modes at the bench size: upload alone and the full e2e step (upload + pipeline + scores D2H)
from pinned host memory in the Rust usize layout. Appends one JSON line per measurement to gpurun_out/e2e_ab.jsonl."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from singlerust_b200 import _ffi, synth  # noqa: E402

n = int(os.environ.get("AB_CELLS", "1000000"))
m, hvg, k, reps = 30000, 2000, 50, int(os.environ.get("AB_REPS", "3"))
out = os.path.join("gpurun_out", "e2e_ab.jsonl")
os.makedirs("gpurun_out", exist_ok=True)


def emit(d):
    with open(out, "a") as f:
        f.write(json.dumps(d) + "\n")
    print(d, flush=True)


ctx = _ffi.Context(0)
thr, amp = synth.gene_tables(m, seed=0x5EED0002, mean_density=0.05)
src = _ffi.DeviceMatrix.synth(ctx, 0x5EED0002, n, m, thr, amp)
nnz = src.info()["nnz"]
t0 = time.time()
off = torch.empty(n + 1, dtype=torch.int64).pin_memory()
idx = torch.empty(nnz, dtype=torch.int64).pin_memory()
val = torch.empty(nnz, dtype=torch.float32).pin_memory()
scores = torch.empty((n, k), dtype=torch.float64).pin_memory()
_ffi.check(_ffi.lib().srb_mat_download(src._h, _ffi._ptr(off), _ffi._ptr(idx), None, _ffi._ptr(val)))
src.free()
emit({"what": "setup", "nnz": nnz, "pin_and_download_s": time.time() - t0, "threads": os.environ.get("SRB_UPLOAD_THREADS", "default"),
      "cpus": os.cpu_count()})
stream = torch.cuda.ExternalStream(ctx.stream)
ref_sum = None
ALL = ((_ffi.UPLOAD_DEVICE_NARROW, "device_narrow"), (_ffi.UPLOAD_HOST_PACK, "host_pack"),
       (_ffi.UPLOAD_HOST_PACK_VALUES, "host_pack_values"), (_ffi.UPLOAD_HOST_PACK_ADAPTIVE, "host_pack_adaptive"),
       (_ffi.UPLOAD_HOST_PACK_DELTA, "host_pack_delta"), (_ffi.UPLOAD_BALANCED, "balanced"))
want = os.environ.get("AB_MODES")
for mode, name in [x for x in ALL if (not want or x[1] in want.split(","))]:
    ctx.set_upload_mode(mode)
    for what in ("upload", "e2e"):
        ts = []
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.synchronize()
            w0 = time.perf_counter()
            e0.record(stream)
            mt = _ffi.DeviceMatrix.upload(ctx, _ffi.CSR, n, m, off, idx, val, nnz=nnz, idx_width=8, dtype=_ffi.F32)
            if what == "e2e":
                mt.pipeline_normalize_hvg_pca(1e4, hvg, k, scores_out=scores)
            e1.record(stream)
            ctx.synchronize()
            w1 = time.perf_counter()
            if r:  # first repetition is the warm-up (ring allocation, first touch)
                ts.append((e0.elapsed_time(e1), (w1 - w0) * 1e3))
            if what == "upload" and r == reps:
                s = mt.sum(_ffi.COLUMN)
                if ref_sum is None:
                    ref_sum = s
                assert np.array_equal(s, ref_sum), "upload modes disagree"
            mt.free()
        emit({"what": what, "mode": name, "event_ms": [round(t[0], 2) for t in ts], "wall_ms": [round(t[1], 2) for t in ts],
              "cells_per_s": n / (min(t[0] for t in ts) * 1e-3), "h2d_bytes": ctx.last_upload()[0],
              "chunks_idx_val_packed": ctx.last_upload_chunks()})
ctx.close()
