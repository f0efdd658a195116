#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): cells/s through
total-count normalise + log1p -> per-gene moments -> HVG(top 2000) -> 50-PC PCA on a synthetic
1M-cell x 30k-gene CSR (5 % nnz) per GPU.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port), rank 0 only

A "step" is one pass of the whole pipeline over one device-resident batch (the rank's row shard). Weak scaling:
every rank holds `--cells` rows of one global matrix of N * cells rows; the per-gene moments and the Gram matrix
are allreduced over NCCL inside the step. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells_per_sec_normalise_hvg_pca"
UNIT = "cells/s"
SEED = 0x5EED0002
TARGET_SUM = 1e4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000, help="cells per GPU")
    ap.add_argument("--genes", type=int, default=30_000)
    ap.add_argument("--hvg", type=int, default=2000)
    ap.add_argument("--pcs", type=int, default=50)
    ap.add_argument("--gram-mode", type=int, default=int(os.environ.get("SRB_GRAM_MODE", "0")), help="0 tcgen05, 1 fp64 CUDA cores")
    ap.add_argument("--inflight", type=int, default=1, help="batches in flight per GPU (one context + stream + host thread each): "
                    "the latency-bound eigensolver of batch i overlaps the bandwidth-bound kernels of batch i+1")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-inflight", type=int, default=2, help="e2e steps in flight on one GPU (one context + stream + host "
                    "thread each): the upload of step i+1 overlaps the kernels and the score download of step i. N = 1 only")
    ap.add_argument("--upload-mode", default="default", choices=["default", "device_narrow", "host_pack", "auto", "host_pack_values", "host_pack_adaptive", "host_pack_delta"],
                    help="how the e2e leg moves the u64 index array over PCIe (srb_ctx_set_upload_mode); default = library default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period-ms", type=int, default=100, help="nvidia-smi sampling period; 0 disables the sampler")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=0, help="0 = auto")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md): nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device, period_ms=200):
        self.device, self.proc, self.lines, self.period = device, None, [], period_ms

    def start(self):
        if self.period <= 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period), "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, mxc = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take the nearest samples
            sm = [float(l.split(",")[1]) for _, l in self.lines[-3:] if len(l.split(",")) >= 8]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (reference-faithful arithmetic, threaded where the loops allow) on a bounded sample
# ---------------------------------------------------------------------------------------------------------
def cpu_pipeline_once(sample, hvg, pcs):
    """One pass of normalise + log1p + per-gene variance + HVG + selected densify + PCA on the CPU. Returns seconds."""
    from oracle import oracle as O
    from oracle import pca_oracle as P
    t0 = time.perf_counter()
    lm, _, _, _, gv = O.norm_log1p_genevar_omp(sample, TARGET_SUM)           # scale/mod.rs + transform/mod.rs + csr.rs:172-186
    sel = O.select_hvg(gv, hvg)                                               # dim_red/mod.rs:135-140
    dense = O.densify_selected(lm, np.arange(sample.nrows, dtype=np.uint64), sel)  # shared/mod.rs:230-259
    P.pca_fit_transform(dense, min(pcs, len(sel)), True, True)                # PCABuilder.fit/transform (LAPACK, all cores)
    return time.perf_counter() - t0


def cpu_pipeline_reference_faithful(sample, hvg, pcs):
    """The same pipeline with the reference's own pass structure and threading: serial loops, f64 values
    after normalise, three passes for the per-gene variance, map-based selected densify; only the SVD uses all cores
    (the reference hands it to LAPACK / faer). Returns seconds."""
    from oracle import oracle as O
    from oracle import pca_oracle as P
    t0 = time.perf_counter()
    lm = O.log1p(O.normalize_total(sample, TARGET_SUM, O.ROW))        # scale/mod.rs:59-89 + transform/mod.rs:8-62
    gv = O.variance(lm, O.COLUMN)                                      # csr.rs:172-186 (sum, count and sum-of-squares passes)
    sel = O.select_hvg(gv, hvg)
    dense = O.densify_selected(lm, np.arange(sample.nrows, dtype=np.uint64), sel)
    P.pca_fit_transform(dense, min(pcs, len(sel)), True, True)
    return time.perf_counter() - t0


def cpu_baseline(args, sample_cells, repeats=1):
    from oracle import oracle as O
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(args.genes, seed=SEED, mean_density=0.05)
    sample = O.synth_csr(SEED, sample_cells, args.genes, thr, amp)
    times = [cpu_pipeline_once(sample, args.hvg, args.pcs) for _ in range(repeats)]
    t = min(times)
    t_faithful = cpu_pipeline_reference_faithful(sample, args.hvg, args.pcs)
    return {"value": sample_cells / t, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "reference_faithful": {"value": sample_cells / t_faithful, "unit": UNIT, "cores": 1,
                                   "note": "serial stats loops and pass structure as in the reference (its rayon pool is only "
                                           "handed to the SVD); same sample"},
            "sample": f"first {sample_cells} cells of the same synthetic matrix (seed 0x{SEED:X}), full pipeline, "
                      f"{t:.2f} s; stats loops threaded with OpenMP, PCA = LAPACK gesdd via NumPy; the reference's own "
                      f"stats loops are single-threaded (SURVEY F2)"}, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total = args.steps + args.warmup
    sample_cells = args.cpu_sample_cells or int(max(2048, min(16384, 120_000 // max(1, n_total))))
    from oracle import oracle as O
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(args.genes, seed=SEED, mean_density=0.05)
    sample = O.synth_csr(SEED, sample_cells, args.genes, thr, amp)
    for _ in range(args.warmup):
        cpu_pipeline_once(sample, args.hvg, args.pcs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline_once(sample, args.hvg, args.pcs)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = sample_cells / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample_cells=sample_cells),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                         "sample": f"{sample_cells} cells x {args.genes} genes per step (bounded sample of the workload)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust and cannot be built in this image (no cargo/rustc); this is the oracle port of its CPU path",
    }
    print(json.dumps(line))


def workload_config(args, sample_cells=None):
    return {"workload": "configs[2]: 1M cells x 30k genes CSR 5% nnz per GPU: normalise + log1p + HVG(top 2k) + 50-PC PCA",
            "cells_per_gpu": args.cells if sample_cells is None else sample_cells, "genes": args.genes, "hvg": args.hvg,
            "pcs": args.pcs, "target_sum": TARGET_SUM, "seed": f"0x{SEED:X}", "parallelism": f"row-shard x{args.gpus}", "batches_in_flight_per_gpu": getattr(args, "inflight", 1),
            "l2": "inputs (12 GB/GPU) are larger than L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from singlerust_b200 import _ffi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    L = max(1, args.inflight)
    ctxs = [_ffi.Context(local_rank) for _ in range(L)]
    if world > 1:
        for c in ctxs:  # one NCCL communicator per in-flight lane
            obj = [_ffi.Context.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            c.comm_init(obj[0], rank, world)
    ctx = ctxs[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for c in ctxs:
            c.synchronize()

    thr, amp = synth.gene_tables(args.genes, seed=SEED, mean_density=0.05)
    mats = []
    for c in ctxs:  # every lane owns a device-resident copy of the rank's shard (same seed => identical data)
        mt = _ffi.DeviceMatrix.synth(c, SEED, args.cells, args.genes, thr, amp, row0=rank * args.cells)
        mt.set_shard(rank * args.cells, world * args.cells)
        mats.append(mt)
    mat = mats[0]
    info = mat.info()
    nnz = info["nnz"]

    def step(lane=0):
        work = mats[lane].clone()  # copy-on-write: the fused kernel reads the raw counts and writes a fresh value buffer
        work.pipeline_normalize_hvg_pca(TARGET_SUM, args.hvg, args.pcs, gram_mode=args.gram_mode, want_outputs=False)
        st = ctxs[lane].last_stage_ms()
        work.free()
        return st

    def run_steps(n_steps):
        """n_steps pipeline passes, round-robin over the lanes, one host thread per lane. Returns the stage times."""
        out = [[] for _ in range(L)]

        def worker(lane):
            for _ in range(lane, n_steps, L):
                out[lane].append(step(lane))

        if L == 1:
            worker(0)
        else:
            ts = [threading.Thread(target=worker, args=(lane,)) for lane in range(L)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return [s for lane in out for s in lane]

    streams = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]
    run_steps(max(args.warmup, L))
    # untimed settling: lazily-loaded library modules (cuSOLVER) reach steady state at different speeds on a cold
    # box; keep stepping (at most 8 more rounds) until two consecutive rounds agree within 5 %
    # (the stop decision is taken collectively: every rank must run the same number of steps, each has 2 allreduces)
    prev = None
    for _ in range(8):
        t0 = time.perf_counter()
        run_steps(L)
        dt = time.perf_counter() - t0
        unsettled = torch.tensor([0.0 if (prev is not None and abs(dt - prev) <= 0.05 * prev) else 1.0], device=dev)
        if world > 1:
            dist.all_reduce(unsettled, op=dist.ReduceOp.MAX)
        if float(unsettled.item()) == 0.0:
            break
        prev = dt
    barrier()
    sampler = ClockSampler(local_rank, args.clock_period_ms)
    sampler.start()
    time.sleep(0.25)
    launches0 = _ffi.kernel_launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    barrier()
    t_wall0 = time.time()
    for e, st_ in zip(ev0, streams):
        e.record(st_)
    stages = run_steps(args.steps)
    for e, st_ in zip(ev1, streams):
        e.record(st_)
    barrier()
    t_wall1 = time.time()
    # device time of the region: earliest start event to latest end event (all lanes are on the same device)
    ms_total = max(a.elapsed_time(b) for a in ev0 for b in ev1)
    launches = _ffi.kernel_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * args.cells / (ms_step * 1e-3)

    # ---- rooflines from the library's per-stage CUDA-event times (ctx stream), averaged over the timed steps ----
    hbm, tf_burst, tf_sust, peak_src = measured_peaks()
    mean_stage = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
    n, d = args.cells, min(args.hvg, args.genes)
    dpad = (d + 255) // 256 * 256
    alg = {
        "row_sums": 4 * nnz + 8 * (n + 1),                       # K1: f32 values + offsets
        "fused_norm_log1p_moments": 12 * nnz + 16 * n,            # K4: idx + val read, val write, offsets + scale
        "densify": 8 * nnz + 8 * n + 4 * n * dpad,                # K6: idx + val read, split-fp16 panels written
        "scores": 4 * n * dpad + 8 * n * args.pcs,                # K9: panels read, f64 scores written
    }
    roof = {}
    for k, b in alg.items():
        ms = mean_stage.get(k, 0.0)
        if ms > 0:
            a = b / (ms * 1e-3) / 1e9
            roof[k] = {"bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm, "traffic": None,
                       "ms": ms, "algorithmic_bytes": b}
    gms = mean_stage.get("gram", 0.0)
    if gms > 0:
        a = n * d * d / (gms * 1e-3) / 1e12  # SYRK algorithmic flops n*d^2
        nt = dpad // 256
        executed = 6.0 * (nt * (nt + 1) // 2) * 65536 * n  # 3 split MMAs x 2 flop x upper-triangular 256x256 tiles x n cells
        ex = executed / (gms * 1e-3) / 1e12
        roof["gram"] = {"bound": "tensor", "achieved": a, "peak": tf_sust, "unit": "TFLOP/s", "frac": a / tf_sust,
                        "traffic": None, "ms": gms, "algorithmic_flops": n * d * d, "executed_flops": executed,
                        "executed_tflops": ex, "frac_executed": ex / tf_sust,
                        "note": "algorithmic = SYRK n*d^2; executed = split-fp16 x3 on upper-triangular 256x256 pair tiles "
                                "(tcgen05 cta_group::2); peak = sustained (power-capped) dense bf16"}
    # DRAM traffic per launch from the committed ncu --set full capture (profiles/), only for the exact config it was taken on
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_L.json")) as f:
            tr = json.load(f)
        if tr["config"] == {"cells_per_gpu": args.cells, "genes": args.genes, "hvg": args.hvg, "pcs": args.pcs}:
            for k, b in tr["traffic_bytes"].items():
                if k in roof:
                    roof[k]["traffic"] = b
    except (OSError, KeyError, ValueError):
        pass
    dominant = max(roof, key=lambda k: roof[k]["ms"]) if roof else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 values / f64 accumulate (fp16x2-split tensor Gram)" if args.gram_mode == 0 else "f32 values / f64 accumulate",
        "data": "synthetic", "config": workload_config(args), "gpu_launches": int(launches), "clocks": clocks,
        "nnz_per_gpu": int(nnz), "stage_ms": mean_stage, "eig_solver": ctx.last_eig(), "peaks": peak_src,
        "roofline": dict(roof[dominant], kernel=dominant) if dominant else None,
        "rooflines": roof,
    }

    # ---- e2e: through the C ABI with HOST buffers (pinned), H2D + pipeline + D2H of the scores, every step ----
    if not args.no_e2e:
        if rank == 0:
            print(json.dumps(dict(line, partial="main line before the e2e leg")), file=sys.stderr, flush=True)
        if world == 1:
            try:
                line["e2e"] = run_e2e(args, ctx, mat, rank, world, dev, barrier)
            except Exception as ex:  # never lose the main line
                line["e2e"] = {"value": None, "unit": UNIT, "error": str(ex)[:300]}
        else:
            line["e2e"] = run_e2e(args, ctx, mat, rank, world, dev, barrier)
    if rank == 0 and not args.no_cpu_baseline:
        try:
            sample = args.cpu_sample_cells or 16384
            cb, _ = cpu_baseline(args, sample)
            line["cpu_baseline"] = cb
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": str(ex)[:300]}
    if rank == 0:
        print(json.dumps(line))
    for mt in mats:
        mt.free()
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, ctx, mat, rank, world, dev, barrier):
    """End to end through the C ABI with HOST buffers: every step uploads the rank's CSR from pinned host memory in the
    reference's own layout (u64 offsets + u64 indices + f32 values), runs the pipeline and reads the scores back.
    The pinned host copy is capped at ~40 GB per node, so at N >= 4 each rank uses the first 40 GB / (N * 18 KB) of its
    cells (stated in the result). Every decision that could differ between ranks is agreed on with a collective first."""
    import psutil
    import torch
    import torch.distributed as dist
    from singlerust_b200 import _ffi

    def agree_min(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    info = mat.info()
    n_full, nnz_full, k = info["nrows"], info["nnz"], min(args.pcs, args.hvg)
    per_cell = 12.0 * nnz_full / max(1, n_full) + 8 * k + 8
    budget = min(40e9, 0.5 * psutil.virtual_memory().available) / max(1, world)
    n = int(agree_min(min(n_full, budget // per_cell)))
    if n < 1024:
        return {"value": None, "unit": UNIT, "skipped": "not enough host memory for a pinned copy of the input"}
    # the e2e input: the first n cells of this rank's shard, regenerated on the device and copied to pinned host memory
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(args.genes, seed=SEED, mean_density=0.05)
    src = mat if n == n_full else _ffi.DeviceMatrix.synth(ctx, SEED, n, args.genes, thr, amp, row0=rank * args.cells)
    nnz = src.info()["nnz"]
    ok = 1.0
    try:
        off = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        idx = torch.empty(nnz, dtype=torch.int64).pin_memory()
        val = torch.empty(nnz, dtype=torch.float32).pin_memory()
        scores = torch.empty((n, k), dtype=torch.float64).pin_memory()
        _ffi.check(_ffi.lib().srb_mat_download(src._h, _ffi._ptr(off), _ffi._ptr(idx), None, _ffi._ptr(val)))
    except Exception:
        ok = 0.0
    if agree_min(ok) < 1.0:
        return {"value": None, "unit": UNIT, "skipped": "pinned host allocation failed on at least one rank"}
    if src is not mat:
        src.free()

    modes = {"device_narrow": _ffi.UPLOAD_DEVICE_NARROW, "host_pack": _ffi.UPLOAD_HOST_PACK, "auto": _ffi.UPLOAD_AUTO,
             "host_pack_values": _ffi.UPLOAD_HOST_PACK_VALUES, "host_pack_adaptive": _ffi.UPLOAD_HOST_PACK_ADAPTIVE,
             "host_pack_delta": _ffi.UPLOAD_HOST_PACK_DELTA}
    # lanes: independent contexts on this GPU; with more than one rank every lane would need its own communicator and a
    # rank-consistent collective order, so pipelining is a single-GPU feature
    L = max(1, args.e2e_inflight) if world == 1 else 1
    lanes = [ctx] + [_ffi.Context(ctx.device) for _ in range(L - 1)]
    for c in lanes:
        if args.upload_mode != "default":
            c.set_upload_mode(modes[args.upload_mode])
    lane_scores = [scores] + [torch.empty((n, k), dtype=torch.float64).pin_memory() for _ in range(L - 1)]

    link = threading.Lock()  # one upload at a time: the PCIe link and the packing threads are the shared resource, and a
    # lane that holds them alone finishes sooner and starts computing while the next lane uploads

    def step(lane=0):
        c = lanes[lane]
        with link:
            m = _ffi.DeviceMatrix.upload(c, _ffi.CSR, n, args.genes, off, idx, val, nnz=nnz, idx_width=8, dtype=_ffi.F32)
        m.set_shard(rank * n, world * n)
        m.pipeline_normalize_hvg_pca(TARGET_SUM, args.hvg, args.pcs, gram_mode=args.gram_mode, scores_out=lane_scores[lane])
        m.free()

    def timed(n_steps, n_lanes):
        """n_steps e2e steps round-robin over n_lanes host threads; device time from the first start event to the last
        end event, max over ranks."""
        use = list(range(n_lanes))
        streams = [torch.cuda.ExternalStream(lanes[l].stream, device=dev) for l in use]
        e0 = [torch.cuda.Event(enable_timing=True) for _ in use]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in use]

        def worker(l):
            for _ in range(l, n_steps, n_lanes):
                step(l)

        barrier()
        for l in use:
            lanes[l].synchronize()
        for e, st_ in zip(e0, streams):
            e.record(st_)
        if n_lanes == 1:
            worker(0)
        else:
            ts = [threading.Thread(target=worker, args=(l,)) for l in use]
            for t_ in ts:
                t_.start()
            for t_ in ts:
                t_.join()
        for e, st_ in zip(e1, streams):
            e.record(st_)
        for l in use:
            lanes[l].synchronize()
        barrier()
        t = torch.tensor([max(a.elapsed_time(b) for a in e0 for b in e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_steps

    for l in range(L):
        step(l)  # warm-up (allocations, staging ring, first touch of the pinned pages)
    ms_seq = timed(args.e2e_steps, 1)
    ms = ms_seq
    if L > 1:
        ms_pipe = timed(max(args.e2e_steps, 5) * L, L)
        ms = min(ms_seq, ms_pipe)
    h2d, packed = ctx.last_upload()  # bytes that actually crossed PCIe (the library's own count)
    for c in lanes[1:]:
        c.close()
    d2h = 8 * n * k + 8 * min(args.hvg, args.genes) * (k + 1) + 8 * k
    return {"value": world * n / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": ms, "steps": args.e2e_steps, "cells_per_gpu": n,
            "steps_in_flight": (L if (L > 1 and ms == ms_pipe) else 1), "ms_per_step_sequential": ms_seq,
            "ms_per_step_pipelined": (ms_pipe if L > 1 else None),
            "host_input_bytes_per_step": int(8 * (n + 1) + 12 * nnz), "upload_mode": "host_pack" if packed else "device_narrow",
            "host_layout": "u64 offsets + u64 indices + f32 values in pinned memory (the Rust usize layout)"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
