"""torchrun script: the row-sharded pipeline on N GPUs must reproduce the 1-GPU result — HVG list and per-gene integer
moments bit for bit, scores / loadings of every well-separated component within 1e-5 — and the CPU oracle's (exact SVD)
within the same tolerance. Launched by tests/test_api_mirror_gpu.py or by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from singlerust_b200 import _ffi, synth  # noqa: E402
from singlerust_b200.parallel import comm_init_from_torch, shard_rows  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, n_top, k = 60000, 4000, 300, 10
    thr, amp = synth.gene_tables(m, seed=21, mean_density=0.05)
    ctx = _ffi.Context(local)
    comm_init_from_torch(ctx)
    a, b = shard_rows(n, world, rank)
    shard = _ffi.DeviceMatrix.synth(ctx, 0x5EED0004, b - a, m, thr, amp, row0=a, skew=True)
    shard.set_shard(a, n)
    res = shard.pipeline_normalize_hvg_pca(1e4, n_top, k)
    gene_var = shard.variance(_ffi.COLUMN)
    gene_cnt = shard.number(_ffi.COLUMN)
    # gather the sharded scores on rank 0
    sc = torch.from_numpy(res["scores"]).cuda()
    sizes = [shard_rows(n, world, r)[1] - shard_rows(n, world, r)[0] for r in range(world)]
    parts = [torch.empty((s, k), dtype=torch.float64, device="cuda") for s in sizes]
    dist.all_gather(parts, sc)
    ok = True
    if rank == 0:
      try:
        from oracle import oracle as O
        from oracle import pca_oracle as P
        from tests._util import sign_align
        ctx1 = _ffi.Context(local)
        whole = _ffi.DeviceMatrix.synth(ctx1, 0x5EED0004, n, m, thr, amp, skew=True)
        ref = whole.pipeline_normalize_hvg_pca(1e4, n_top, k)
        assert np.array_equal(ref["selection"], res["selection"]), "HVG list differs between 1 and N GPUs"
        assert np.array_equal(whole.number(_ffi.COLUMN), gene_cnt)
        assert np.array_equal(whole.variance(_ffi.COLUMN), gene_var), "integer moments must be bit-identical across shardings"
        # tcgen05 Gram: the fp32 chunk sums fall on different cell boundaries in every sharding (measured 1e-8)
        np.testing.assert_allclose(res["explained_variance_ratio"], ref["explained_variance_ratio"], rtol=1e-6)
        got = torch.cat(parts).cpu().numpy()

        def well_separated(ev, k, rel_gap=1e-3):
            return np.array([min([ev[j - 1] - ev[j]] * (j > 0) + [ev[j] - ev[j + 1]]) > rel_gap * ev[j] for j in range(k)])

        # (a) N GPUs vs 1 GPU: fp64 partial Gram sums re-associate, nothing else differs
        rms = np.linalg.norm(ref["scores"], axis=0) / np.sqrt(n)
        err = np.max(np.abs(sign_align(got, ref["scores"]) - ref["scores"]), axis=0) / rms
        cerr = np.max(np.abs(sign_align(res["components"], ref["components"]) - ref["components"]), axis=0)
        # (b) N GPUs vs the CPU oracle (exact SVD PCA on the device's stored values)
        off, idx, val = whole.download()
        ol = O.Compressed("csr", n, m, off, idx, val)
        want = P.pca_pipeline(ol, n_top, k, selection=res["selection"])
        good = well_separated(want["eigenvalues"], k)
        oerr = np.max(np.abs(sign_align(res["components"], want["components"]) - want["components"]), axis=0)
        oserr = np.max(np.abs(sign_align(got, want["scores"]) - want["scores"]), axis=0) / (np.linalg.norm(want["scores"], axis=0) / np.sqrt(n))
        print("N vs 1 GPU: score err / rms", ["%.1e" % e for e in err], "loadings", ["%.1e" % e for e in cerr])
        print("N GPUs vs oracle: loadings", ["%.1e" % e for e in oerr], "scores / rms", ["%.1e" % e for e in oserr], "well separated:", good.tolist())
        np.testing.assert_allclose(res["explained_variance_ratio"], want["explained_variance_ratio"], rtol=1e-5)
        np.testing.assert_array_equal(res["selection"], O.select_hvg(O.variance(ol, O.COLUMN), n_top))
        assert good.sum() >= 1 and good[0]
        for j in np.nonzero(good)[0]:
            # loadings: the north_star tolerance (1e-5). Scores: the worst cell over the component's rms, 10 x that, as in
            # tests/test_gpu_parity.py (the scores kernel accumulates 2048 genes in fp32 TMEM chunks)
            assert cerr[j] < 1e-5 and err[j] < 1e-4, ("N vs 1 GPU", j, err[j], cerr[j])
            assert oerr[j] < 1e-5 and oserr[j] < 1e-4, ("N GPUs vs oracle", j, oerr[j], oserr[j])
        print("MULTIGPU OK world =", world, flush=True)
      except Exception:  # the other ranks must not be left waiting in the barrier below
        import traceback
        traceback.print_exc()
        ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)


if __name__ == "__main__":
    main()
