"""torchrun script: the row-sharded pipeline on N GPUs must reproduce the 1-GPU result — HVG list and per-gene integer
moments bit for bit, scores / loadings within 1e-5. Launched by tests/test_api_mirror_gpu.py or by hand:
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
    if rank == 0:
        ctx1 = _ffi.Context(local)
        whole = _ffi.DeviceMatrix.synth(ctx1, 0x5EED0004, n, m, thr, amp, skew=True)
        ref = whole.pipeline_normalize_hvg_pca(1e4, n_top, k)
        assert np.array_equal(ref["selection"], res["selection"]), "HVG list differs between 1 and N GPUs"
        assert np.array_equal(whole.number(_ffi.COLUMN), gene_cnt)
        assert np.array_equal(whole.variance(_ffi.COLUMN), gene_var), "integer moments must be bit-identical across shardings"
        np.testing.assert_allclose(res["explained_variance_ratio"], ref["explained_variance_ratio"], rtol=1e-6)
        got = torch.cat(parts).cpu().numpy()
        s = np.sign(np.sum(got * ref["scores"], axis=0))
        rms = np.linalg.norm(ref["scores"], axis=0) / np.sqrt(n)
        err = np.max(np.abs(got * s - ref["scores"]), axis=0) / rms
        print("score err / rms per component:", err)
        assert np.all(err[:3] < 1e-4)
        print("MULTIGPU OK world =", world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
