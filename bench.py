#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): cells/s through
total-count normalise + log1p -> per-gene moments -> HVG(top 2000) -> 50-PC PCA on a synthetic
1M-cell x 30k-gene CSR (5 % nnz) per GPU.

  python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port), rank 0 only

A "step" is one pass of the whole pipeline over one device-resident batch (the rank's row shard). Weak scaling (the
default, and the headline `value`): every rank holds `--cells` rows of one global matrix of N * cells rows; the per-gene
moments and the Gram matrix are allreduced over NCCL inside the step. Prints ONE JSON line on rank 0.

The same line also carries, measured in the same run (each with its own CUDA-event timing, max over ranks):
  strong    BASELINE.md L-full "strong scaling": the 1M-cell matrix split over the N ranks
  xl        configs[3]: 4M cells x 30k genes (seed 0x5EED0004) row-sharded over the N ranks
  xxl       configs[4]: the backed chunk stream (131 072-row chunks from pinned host memory, 1.25M cells per GPU = 10M on 8)
            through the full pipeline: chunks uploaded once and kept resident, and the one-chunk-resident three-pass form
  faithful  the headline step in SRB_VALUES_FAITHFUL mode (f64 values and accumulation, the reference's arithmetic)
  pipelined the headline workload with two batches in flight on the GPU (N = 1): throughput when the eigensolver of one
            batch overlaps the streaming kernels of the next
`--scaling strong` / `--config xl|xxl` make one of them the main line instead.
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
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-host-memory variant of the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period-ms", type=int, default=100, help="nvidia-smi sampling period; 0 disables the sampler")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=0, help="0 = auto")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --cells per GPU; strong: --cells in total")
    ap.add_argument("--config", default="l", choices=["l", "xl", "xxl"], help="BASELINE.json configs[2] (default) / [3] / [4] as the main line")
    ap.add_argument("--no-legs", action="store_true", help="skip the strong / xl / xxl / faithful legs")
    ap.add_argument("--legs", default="strong,xl,xxl,faithful,pipelined", help="comma list of legs to run after the main line")
    ap.add_argument("--xxl-cells-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--xxl-chunk", type=int, default=131_072)
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


def effective_cpus() -> int:
    """CPUs this process may use: affinity mask and cgroup quota (os.cpu_count() sees neither)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def use_all_host_threads() -> int:
    """The CPU legs use every host core they may: torchrun exports OMP_NUM_THREADS=1 to its workers, which silently made
    the round-1 reference arm single-threaded at N > 1. Raises the OpenMP and BLAS pools at run time; returns the count."""
    n = effective_cpus()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


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


CPU_SAMPLE_CELLS = 16_384  # the one bounded sample both CPU legs (cpu_baseline and --impl reference) are timed on


def cpu_sample(args, sample_cells):
    from oracle import oracle as O
    from singlerust_b200 import synth
    thr, amp = synth.gene_tables(args.genes, seed=SEED, mean_density=0.05)
    return O.synth_csr(SEED, sample_cells, args.genes, thr, amp)


def cpu_baseline(args, sample_cells, repeats=1):
    from oracle import oracle as O
    cores = use_all_host_threads()
    sample = cpu_sample(args, sample_cells)
    times = [cpu_pipeline_once(sample, args.hvg, args.pcs) for _ in range(repeats)]
    t = min(times)
    t_faithful = cpu_pipeline_reference_faithful(sample, args.hvg, args.pcs)
    return {"value": sample_cells / t, "unit": UNIT, "cores": min(cores, O.num_threads()), "kind": "port",
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
    cores = use_all_host_threads()  # before the oracle library loads: torchrun sets OMP_NUM_THREADS=1 for its workers
    sample_cells = args.cpu_sample_cells or CPU_SAMPLE_CELLS
    from oracle import oracle as O
    sample = cpu_sample(args, sample_cells)
    for _ in range(args.warmup):
        cpu_pipeline_once(sample, args.hvg, args.pcs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline_once(sample, args.hvg, args.pcs)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = sample_cells / dt
    t_faithful = cpu_pipeline_reference_faithful(sample, args.hvg, args.pcs)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample_cells=sample_cells),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": min(cores, O.num_threads()), "kind": "port",
                         "reference_faithful": {"value": sample_cells / t_faithful, "unit": UNIT, "cores": 1,
                                                "note": "serial stats loops and pass structure as in the reference; same sample"},
                         "sample": f"first {sample_cells} cells x {args.genes} genes of the same synthetic matrix per step (the "
                                   f"sample bench.py's own cpu_baseline leg uses)"},
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
SEED_XL, SEED_XXL = 0x5EED0004, 0x5EED0005
TOTAL_XL = 4_000_000


class Env:
    """Rank bookkeeping + the torch.distributed plumbing (rendezvous, barriers, max-over-ranks) of one bench process."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world == args.gpus or self.world == 1, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctxs = []

    def new_ctx(self, value_mode=0):
        from singlerust_b200 import _ffi
        c = _ffi.Context(self.local, value_mode=value_mode)
        if self.world > 1:  # one NCCL communicator per context
            obj = [_ffi.Context.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(obj, src=0)
            c.comm_init(obj[0], self.rank, self.world)
        self.ctxs.append(c)
        return c

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        for c in self.ctxs:
            if c._h:
                c.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x: float) -> float:
        return -self.max_over_ranks(-x)

    def sum_over_ranks(self, arr):
        """f64 NumPy vector summed over ranks (identical result on every rank: NCCL allreduce)."""
        if self.world == 1:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr, np.float64)).to(self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


def timed_pipeline(env, args, ctxs, mats, steps, warmup, settle=True, sample_clocks=False):
    """`steps` timed passes of normalise + log1p + HVG + PCA over the device-resident shards `mats` (one per lane), after
    `warmup` untimed ones: CUDA events on the context streams, barrier + synchronize on both sides, max over ranks."""
    from singlerust_b200 import _ffi
    torch = env.torch
    L = len(ctxs)

    def step(lane=0):
        work = mats[lane].clone()  # copy-on-write: the fused kernel reads the raw counts and writes a fresh value buffer
        work.pipeline_normalize_hvg_pca(TARGET_SUM, args.hvg, args.pcs, gram_mode=args.gram_mode, want_outputs=False)
        st = ctxs[lane].last_stage_ms()
        work.free()
        return st

    def run_steps(n_steps):
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

    streams = [torch.cuda.ExternalStream(c.stream, device=env.dev) for c in ctxs]
    run_steps(max(warmup, L))
    if settle:
        # untimed settling: lazily-loaded library modules reach steady state at different speeds on a cold box; keep
        # stepping (at most 8 more rounds) until two consecutive rounds agree within 5 % (decided collectively: every
        # rank must run the same number of steps, each has two allreduces)
        prev = None
        for _ in range(8):
            t0 = time.perf_counter()
            run_steps(L)
            dt = time.perf_counter() - t0
            if env.max_over_ranks(0.0 if (prev is not None and abs(dt - prev) <= 0.05 * prev) else 1.0) == 0.0:
                break
            prev = dt
    env.barrier()
    sampler = ClockSampler(env.local, args.clock_period_ms if sample_clocks else 0)
    sampler.start()
    if sample_clocks:
        time.sleep(0.25)
    launches0 = _ffi.kernel_launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    env.barrier()
    t_wall0 = time.time()
    for e, st_ in zip(ev0, streams):
        e.record(st_)
    stages = run_steps(steps)
    for e, st_ in zip(ev1, streams):
        e.record(st_)
    env.barrier()
    t_wall1 = time.time()
    ms_total = max(a.elapsed_time(b) for a in ev0 for b in ev1)  # earliest start to latest end (lanes share the device)
    launches = _ffi.kernel_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sample_clocks else None
    ms_step = env.max_over_ranks(ms_total) / steps
    mean_stage = {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}
    return dict(ms_per_step=ms_step, stage_ms=mean_stage, launches=int(launches), clocks=clocks)


def make_shards(env, args, ctxs, seed, cells_total=None, cells_per_gpu=None):
    """One device-resident row shard per lane of the global matrix (seed): rank r owns rows [r n/G, (r+1) n/G)."""
    from singlerust_b200 import _ffi, synth
    from singlerust_b200.parallel import shard_rows
    thr, amp = synth.gene_tables(args.genes, seed=SEED, mean_density=0.05)
    if cells_total is None:
        cells_total = cells_per_gpu * env.world
    a, b = shard_rows(cells_total, env.world, env.rank)
    mats = []
    for c in ctxs:
        mt = _ffi.DeviceMatrix.synth(c, seed, b - a, args.genes, thr, amp, row0=a)
        mt.set_shard(a, cells_total)
        mats.append(mt)
    return mats, cells_total, b - a


def stage_rooflines(args, n, nnz, mean_stage):
    """Achieved fraction of the measured peaks per stage, from the library's per-stage CUDA-event times (context stream):
    algorithmic bytes / flops per launch (SURVEY §8d accounting: f32 values, u32 indices, i64 offsets) over the mean time."""
    hbm, tf_burst, tf_sust, peak_src = measured_peaks()
    d = min(args.hvg, args.genes)
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
        executed = gram_executed_flops(n, d)
        ex = executed / (gms * 1e-3) / 1e12
        roof["gram"] = {"bound": "tensor", "achieved": a, "peak": tf_sust, "unit": "TFLOP/s", "frac": a / tf_sust,
                        "traffic": None, "ms": gms, "algorithmic_flops": n * d * d, "executed_flops": executed,
                        "executed_tflops": ex, "frac_executed": ex / tf_sust,
                        "note": "algorithmic = SYRK n*d^2; executed = the split-fp16 tcgen05 MMAs the kernel issues "
                                "(cta_group::2 pair tiles); peak = sustained (power-capped) dense bf16"}
    # DRAM traffic per launch from a committed ncu --set full capture — only when that capture was taken on this very
    # source (hash of csrc/) and this very config; otherwise null
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_L.json")) as f:
            tr = json.load(f)
        if tr["config"] == {"cells_per_gpu": n, "genes": args.genes, "hvg": args.hvg, "pcs": args.pcs} and \
                tr.get("csrc_sha256") == csrc_hash():
            for k, b in tr["traffic_bytes"].items():
                if k in roof:
                    roof[k]["traffic"] = b
    except (OSError, KeyError, ValueError):
        pass
    return roof, peak_src


def gram_executed_flops(n, d):
    """Flops the default Gram kernel issues (csrc/gram_tc2.cu tile list): 3 split MMAs per 256 x 256 pair tile of the upper
    triangle; with SRB_GRAM_TRIM=1 the tiles of the last tile column run with N trimmed to the selected genes, rounded to 32."""
    return 2.0 * 256 * gram_tile_columns(d) * n


def gram_tile_columns(d):
    """Accumulator columns (x 3 split terms) summed over the scheduled tiles, kept in step with gram_tcgen05_pair()."""
    nt = (d + 255) // 256
    rem = d - (nt - 1) * 256
    last_n = max(32, min(256, (rem + 31) // 32 * 32)) if os.environ.get("SRB_GRAM_TRIM", "0")[:1] == "1" else 256
    return 3 * ((nt * (nt + 1) // 2 - nt) * 256 + nt * last_n)


def gram_tile_products(d):
    """256 x 256 tile-product equivalents per 1-cell slice of the contraction."""
    return gram_tile_columns(d) / 256.0


def csrc_hash():
    import hashlib
    h = hashlib.sha256()
    base = os.path.join(ROOT, "singlerust_b200", "csrc")
    for f in sorted(os.listdir(base)):
        if f.endswith((".cu", ".cuh")):  # device code only: host_pack.cpp cannot change a kernel's DRAM traffic
            with open(os.path.join(base, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def run_ours(args):
    env = Env(args)
    torch = env.torch
    from singlerust_b200 import _ffi
    rank, world = env.rank, env.world

    L = max(1, args.inflight)
    ctxs = [env.new_ctx() for _ in range(L)]
    ctx = ctxs[0]

    if args.config == "xxl":
        res = leg_xxl(env, args, ctx)
        if rank == 0:
            rs = res["resident_stream"]
            print(json.dumps({
                "metric": METRIC, "value": rs["value"], "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": rs["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 values / f64 accumulate (fp16x2-split tensor Gram)", "data": "synthetic",
                "config": {"workload": "configs[4]: backed chunk stream through normalise + HVG + PCA", **res["config"]},
                "gpu_launches": rs["launches"], "xxl": res}))
        finish(env, [], ctxs)
        return

    # ---- main line: configs[2] per GPU (weak), or --scaling strong / --config xl ----
    if args.config == "xl":
        seed, scaling = SEED_XL, "strong"
        mats, cells_total, n = make_shards(env, args, ctxs, seed, cells_total=TOTAL_XL)
        workload = "configs[3]: 4M cells x 30k genes CSR 5% nnz row-sharded over the GPUs: normalise + log1p + HVG(top 2k) + 50-PC PCA"
    elif args.scaling == "strong":
        seed, scaling = SEED, "strong"
        mats, cells_total, n = make_shards(env, args, ctxs, seed, cells_total=args.cells)
        workload = "configs[2], strong scaling: 1M cells x 30k genes in total, row-sharded over the GPUs"
    else:
        seed, scaling = SEED, "weak"
        mats, cells_total, n = make_shards(env, args, ctxs, seed, cells_per_gpu=args.cells)
        workload = None
    mat = mats[0]
    nnz = mat.info()["nnz"]
    main = timed_pipeline(env, args, ctxs, mats, args.steps, args.warmup, settle=True, sample_clocks=True)
    ms_step = main["ms_per_step"]
    value = cells_total / (ms_step * 1e-3)
    roof, peak_src = stage_rooflines(args, n, nnz, main["stage_ms"])
    dominant = max(roof, key=lambda k: roof[k]["ms"]) if roof else None
    cfg = workload_config(args)
    if workload:
        cfg["workload"], cfg["cells_total"], cfg["cells_per_gpu"] = workload, cells_total, n
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32 values / f64 accumulate (fp16x2-split tensor Gram)" if args.gram_mode == 0 else "f32 values / f64 accumulate",
        "data": "synthetic", "config": cfg, "gpu_launches": main["launches"], "clocks": main["clocks"],
        "nnz_per_gpu": int(nnz), "stage_ms": main["stage_ms"], "eig_solver": ctx.last_eig(), "peaks": peak_src,
        "roofline": dict(roof[dominant], kernel=dominant) if dominant else None,
        "rooflines": roof,
    }
    if rank == 0:
        print(json.dumps(dict(line, partial="main line before the extra legs")), file=sys.stderr, flush=True)

    # ---- extra legs (each guarded: the main line must survive a failing leg at N = 1; at N > 1 a failure on one rank
    #      would leave the others in a collective, so errors there are fatal on purpose) ----
    legs = [] if (args.no_legs or args.config != "l" or args.scaling != "weak") else [x for x in args.legs.split(",") if x]

    def guarded(name, fn):
        if world > 1:
            line[name] = fn()
            return
        try:
            line[name] = fn()
        except Exception as ex:
            line[name] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    if "faithful" in legs:
        guarded("faithful", lambda: leg_faithful(env, args, n, cells_total))
        if isinstance(line.get("faithful"), dict) and "value" in line["faithful"]:
            line["value_faithful"], line["ms_per_step_faithful"] = line["faithful"]["value"], line["faithful"]["ms_per_step"]
    if "pipelined" in legs and L == 1 and world == 1:
        guarded("pipelined", lambda: leg_pipelined(env, args, n, cells_total))
    if "strong" in legs:
        if world == 1:
            line["strong"] = {"cells_total": args.cells, "n_gpus": 1, "value": value, "ms_per_step": ms_step,
                              "note": "at one GPU the strong-scaled step is the main line"}
        else:
            guarded("strong", lambda: leg_sharded(env, args, ctx, SEED, args.cells, "configs[2] strong scaling: 1M cells in total"))
    if "xl" in legs:
        guarded("xl", lambda: leg_sharded(env, args, ctx, SEED_XL, TOTAL_XL, "configs[3]: 4M cells x 30k genes row-sharded"))

    # ---- e2e: through the C ABI with HOST buffers (pinned), H2D + pipeline + D2H of the scores, every step ----
    if not args.no_e2e and args.config == "l" and args.scaling == "weak":
        if world == 1:
            try:
                line["e2e"] = run_e2e(args, ctx, mat, rank, world, env.dev, env.barrier)
            except Exception as ex:  # never lose the main line
                line["e2e"] = {"value": None, "unit": UNIT, "error": str(ex)[:300]}
        else:
            line["e2e"] = run_e2e(args, ctx, mat, rank, world, env.dev, env.barrier)
    for mt in mats:
        mt.free()
    mats = []
    if "xxl" in legs:
        guarded("xxl", lambda: leg_xxl(env, args, ctx))
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cb, _ = cpu_baseline(args, args.cpu_sample_cells or CPU_SAMPLE_CELLS)
            line["cpu_baseline"] = cb
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": str(ex)[:300]}
    if rank == 0:
        print(json.dumps(line))
    finish(env, mats, ctxs)


def finish(env, mats, ctxs):
    for mt in mats:
        mt.free()
    for c in env.ctxs:
        c.close()
    if env.world > 1:
        env.dist.destroy_process_group()


def leg_faithful(env, args, n, cells_total):
    """The headline step with SRB_VALUES_FAITHFUL: values promoted to f64 by normalise / log1p exactly like the reference
    (scale/mod.rs:74-83), f64 accumulation everywhere (the fixed-point moments are replaced by fp64 accumulation)."""
    from singlerust_b200 import _ffi
    c = env.new_ctx(value_mode=_ffi.VALUES_FAITHFUL)
    mats, _, _ = make_shards(env, args, [c], SEED, cells_total=cells_total)
    r = timed_pipeline(env, args, [c], mats, max(2, min(args.steps, 3)), 2, settle=False)
    mats[0].free()
    c.close()
    return {"value": cells_total / (r["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms_per_step"], "stage_ms": r["stage_ms"],
            "dtype": "f64 values / f64 accumulate (SRB_VALUES_FAITHFUL; Gram still on tcgen05 from split-fp16 panels)"}


def leg_pipelined(env, args, n, cells_total):
    """Throughput with two batches in flight on the GPU (two contexts, streams and host threads): the latency-bound
    eigensolver of one batch overlaps the bandwidth-bound kernels of the other. Same work per batch, nothing skipped; the
    main line stays single-lane so that its per-stage times are uncontended."""
    cs = [env.new_ctx() for _ in range(2)]
    mats, _, _ = make_shards(env, args, cs, SEED, cells_total=cells_total)
    r = timed_pipeline(env, args, cs, mats, max(4, min(args.steps, 10)) // 2 * 2, 4, settle=False)
    for mt in mats:
        mt.free()
    for c in cs:
        c.close()
    return {"value": cells_total / (r["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms_per_step"], "batches_in_flight": 2}


def leg_sharded(env, args, ctx, seed, cells_total, what):
    """A matrix of `cells_total` cells row-sharded over the ranks (strong-scaled): same pipeline, same timing rules."""
    mats, cells_total, n = make_shards(env, args, [ctx], seed, cells_total=cells_total)
    nnz = mats[0].info()["nnz"]
    r = timed_pipeline(env, args, [ctx], mats, max(3, min(args.steps, 10)), 3, settle=False)
    mats[0].free()
    roof, _ = stage_rooflines(args, n, nnz, r["stage_ms"])
    return {"workload": what, "cells_total": cells_total, "cells_per_gpu": n, "n_gpus": env.world,
            "value": cells_total / (r["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms_per_step"], "stage_ms": r["stage_ms"],
            "frac": {k: round(v["frac"], 4) for k, v in roof.items()}, "seed": f"0x{seed:X}"}


def leg_xxl(env, args, ctx):
    """configs[4]: the backed chunk iterator (131 072-row CSR chunks in HOST memory, the Rust usize layout) through the full
    pipeline. Per GPU 1.25 M cells (= 10 M on 8 GPUs). Two forms, both timed with the H2D copies inside:
      resident_stream  chunks are uploaded once, appended to one device-resident matrix (srb_stream_set_retain) and the
                       in-memory kernels run on it — the B200-first form: 15 GB per GPU fits HBM many times over;
      three_pass       one chunk resident at a time (data beyond HBM): moments -> Gram -> scores, every pass re-uploads.
    Chunk source: a pool of 2 distinct pinned chunks of the rank's rows of the 0x5EED0005 matrix, cycled (a 22 GB pinned
    copy per rank is not needed to time the path); reported against the H2D rate measured in the same run."""
    import torch
    from singlerust_b200 import _ffi, synth
    from singlerust_b200.backed.processing import select_from_moments
    from singlerust_b200.shared import FeatureSelection
    rank, world, dev = env.rank, env.world, env.dev
    cells, chunk, m, k = args.xxl_cells_per_gpu, args.xxl_chunk, args.genes, min(args.pcs, args.hvg)
    cells_total = cells * world
    thr, amp = synth.gene_tables(m, seed=SEED, mean_density=0.05)
    pool = []
    for i in range(2):
        src = _ffi.DeviceMatrix.synth(ctx, SEED_XXL, chunk, m, thr, amp, row0=rank * cells + i * chunk)
        nnz = src.info()["nnz"]
        off = torch.empty(chunk + 1, dtype=torch.int64).pin_memory()
        idx = torch.empty(nnz, dtype=torch.int64).pin_memory()
        val = torch.empty(nnz, dtype=torch.float32).pin_memory()
        _ffi.check(_ffi.lib().srb_mat_download(src._h, _ffi._ptr(off), _ffi._ptr(idx), None, _ffi._ptr(val)))
        src.free()
        pool.append((off.numpy().view(np.uint64), idx.numpy().view(np.uint64), val.numpy()))
    nchunks = (cells + chunk - 1) // chunk

    def chunks():
        """(offsets, indices, values, rows, nnz) of chunk c: pool entry c % 2, truncated for the last (short) chunk"""
        for c in range(nchunks):
            off, idx, val = pool[c % 2]
            rows = min(chunk, cells - c * chunk)
            nz = int(off[rows])
            yield off[:rows + 1], idx[:nz], val[:nz], rows, nz

    total_nnz = sum(nz for _, _, _, _, nz in chunks())
    # the H2D rate of this box (pinned -> device, 1 GiB, best of 3)
    a = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    b = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    link_gbs = (1 << 30) / (best * 1e-3) / 1e9
    del a, b
    scores = torch.empty((cells, k), dtype=torch.float64).pin_memory()
    scores_np = scores.numpy()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def timed(fn):
        env.barrier()
        l0 = _ffi.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = fn()
        e1.record(stream)
        env.barrier()
        return env.max_over_ranks(e0.elapsed_time(e1)), out, _ffi.kernel_launch_count() - l0

    def resident_stream():
        st = _ffi.ChunkStream(ctx, _ffi.CSR, cells, m)
        st.set_retain(total_nnz, keep_statistics=False)
        h2d = 0
        for off, idx, val, rows, nz in chunks():
            st.push(off, idx, val)
            h2d += ctx.last_upload()[0]
        mt = st.finish_matrix()
        st.free()
        mt.set_shard(rank * cells, cells_total)
        r = mt.pipeline_normalize_hvg_pca(TARGET_SUM, args.hvg, args.pcs, gram_mode=args.gram_mode, scores_out=scores_np)
        mt.free()
        return h2d, r

    def three_pass():
        h2d = 0
        cnt, tot, sq = np.zeros(m), np.zeros(m), np.zeros(m)

        def upload(off, idx, val, rows, nz):
            nonlocal h2d
            dm = _ffi.DeviceMatrix.upload(ctx, _ffi.CSR, rows, m, off, idx, val, nnz=nz)
            h2d += ctx.last_upload()[0]
            dm.normalize_total_inplace(TARGET_SUM, _ffi.ROW)
            dm.log1p_inplace()
            return dm

        for ch in chunks():                                   # pass 1: per-gene moments of the transformed values
            dm = upload(*ch)
            c_, s_, q_ = dm.gene_moments()
            cnt += c_
            tot += s_
            sq += q_
            dm.free()
        red = env.sum_over_ranks(np.concatenate([cnt, tot, sq]))
        cnt, tot, sq = red[:m], red[m:2 * m], red[2 * m:]
        sel, _ = select_from_moments(cnt, tot, sq, FeatureSelection.HighlyVariable(args.hvg))
        ps = _ffi.PcaStream(ctx, m, cells_total, tot, sq, sel, k, True, True, args.gram_mode)
        for ch in chunks():                                   # pass 2: Gram matrix (allreduced inside fit)
            dm = upload(*ch)
            ps.push_gram(dm)
            dm.free()
        comps, evr = ps.fit()
        r0 = 0
        for ch in chunks():                                   # pass 3: scores
            dm = upload(*ch)
            ps.transform(dm, scores_np[r0:r0 + ch[3]])
            r0 += ch[3]
            dm.free()
        ps.free()
        return h2d, dict(selection=sel, explained_variance_ratio=evr, components=comps)

    resident_stream()  # warm-up (staging ring, block cache, lazy modules)
    ms_a, (h2d_a, ra), la = timed(resident_stream)
    ms_b, (h2d_b, rb), lb = timed(three_pass)
    same_sel = bool(np.array_equal(ra["selection"], rb["selection"]))
    evr_diff = float(np.max(np.abs(rb["explained_variance_ratio"] / ra["explained_variance_ratio"] - 1.0)))
    host_bytes = sum(8 * (rows + 1) + 12 * nz for _, _, _, rows, nz in chunks())
    out = {
        "config": {"cells_total": cells_total, "cells_per_gpu": cells, "chunk_rows": chunk, "chunks_per_gpu": nchunks, "genes": m,
                   "hvg": args.hvg, "pcs": k, "seed": f"0x{SEED_XXL:X}", "n_gpus": world,
                   "source": "2 distinct pinned host chunks per rank, cycled; u64 offsets + u64 indices + f32 values"},
        "h2d_link_gbs_measured": link_gbs,
        "resident_stream": {"value": cells_total / (ms_a * 1e-3), "unit": UNIT, "ms": ms_a, "h2d_bytes_per_gpu": int(h2d_a),
                            "h2d_gbs_achieved": h2d_a / (ms_a * 1e-3) / 1e9, "host_input_bytes_per_gpu": int(host_bytes), "launches": int(la)},
        "three_pass": {"value": cells_total / (ms_b * 1e-3), "unit": UNIT, "ms": ms_b, "h2d_bytes_per_gpu": int(h2d_b),
                       "h2d_gbs_achieved": h2d_b / (ms_b * 1e-3) / 1e9, "host_input_bytes_per_gpu": int(3 * host_bytes), "launches": int(lb)},
        "check": {"same_hvg_list": same_sel, "explained_variance_ratio_max_rel_diff": evr_diff,
                  "ok": bool(same_sel and evr_diff < 1e-6)},
    }
    return out


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
    chunks, idx_packed, val_packed = ctx.last_upload_chunks()
    # the same step from PAGEABLE host memory (what a Rust Vec is): the library stages it through its pinned ring
    pageable = None
    if world == 1 and not args.no_pageable:
        try:
            hold = (off, idx, val)
            off, idx, val = off.numpy().copy(), idx.numpy().copy(), val.numpy().copy()  # plain malloc'ed copies

            def step_pageable(lane=0):
                c = lanes[lane]
                with link:
                    m = _ffi.DeviceMatrix.upload(c, _ffi.CSR, n, args.genes, off.view(np.uint64), idx.view(np.uint64), val, nnz=nnz)
                m.set_shard(rank * n, world * n)
                m.pipeline_normalize_hvg_pca(TARGET_SUM, args.hvg, args.pcs, gram_mode=args.gram_mode, scores_out=lane_scores[lane])
                m.free()

            step = step_pageable
            step(0)
            ms_pg = timed(max(args.e2e_steps, 4) if L > 1 else args.e2e_steps, L)
            pageable = {"ms_per_step": ms_pg, "value": world * n / (ms_pg * 1e-3), "steps_in_flight": L,
                        "h2d_bytes_per_step": int(ctx.last_upload()[0])}
            off, idx, val = hold
        except Exception as ex:
            pageable = {"error": str(ex)[:200]}
    # the same step from a pinned host copy with 32-bit indices (the .h5ad on-disk width, which srb_mat_upload accepts as it
    # is): 12 instead of 18 GB of host memory to read per step — what a host shim that does not widen to usize would see
    narrow = None
    if world == 1 and not args.no_pageable:
        try:
            hold = (off, idx, val)
            if nnz >= 2 ** 31:
                raise ValueError("more than 2^31 stored entries: 32-bit offsets do not hold them")
            idx32 = torch.empty(nnz, dtype=torch.int32).pin_memory()
            idx32.copy_(idx)
            off32 = torch.empty(n + 1, dtype=torch.int32).pin_memory()
            off32.copy_(off)

            def step_u32(lane=0):
                c = lanes[lane]
                with link:
                    m = _ffi.DeviceMatrix.upload(c, _ffi.CSR, n, args.genes, off32, idx32, val, nnz=nnz, idx_width=4, dtype=_ffi.F32)
                m.set_shard(rank * n, world * n)
                m.pipeline_normalize_hvg_pca(TARGET_SUM, args.hvg, args.pcs, gram_mode=args.gram_mode, scores_out=lane_scores[lane])
                m.free()

            step = step_u32
            step(0)
            ms_n = timed(max(args.e2e_steps, 4) if L > 1 else args.e2e_steps, L)
            narrow = {"ms_per_step": ms_n, "value": world * n / (ms_n * 1e-3), "steps_in_flight": L,
                      "h2d_bytes_per_step": int(ctx.last_upload()[0]), "host_input_bytes_per_step": int(4 * (n + 1) + 8 * nnz)}
            del idx32, off32
            off, idx, val = hold
        except Exception as ex:
            narrow = {"error": str(ex)[:200]}
    for c in lanes[1:]:
        c.close()
    d2h = 8 * n * k + 8 * min(args.hvg, args.genes) * (k + 1) + 8 * k
    return {"value": world * n / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": ms, "steps": args.e2e_steps, "cells_per_gpu": n,
            "steps_in_flight": (L if (L > 1 and ms == ms_pipe) else 1), "ms_per_step_sequential": ms_seq,
            "ms_per_step_pipelined": (ms_pipe if L > 1 else None),
            "host_input_bytes_per_step": int(8 * (n + 1) + 12 * nnz), "upload_mode": "host_pack" if packed else "device_narrow",
            "upload_chunks": {"chunks": chunks, "index_chunks_host_packed": idx_packed, "value_chunks_host_packed": val_packed},
            "pageable_input": pageable, "u32_index_input": narrow,
            "host_layout": "u64 offsets + u64 indices + f32 values in pinned memory (the Rust usize layout)"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
