#!/bin/bash
# round 2, call 2 (2 GPUs): un-gated suite incl. the 2-GPU parity check, bench with the new legs at N = 1 and N = 2
mkdir -p gpurun_out
S=gpurun_out/c2_summary.txt
: > $S
nvidia-smi -L >> $S; nproc >> $S; free -g | head -2 >> $S
timeout 900 python -m pytest tests -m gpu -q --durations=15 --timeout=300 > gpurun_out/c2_tests.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/c2_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c2_tests.log | head -20 >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tests/multigpu_check.py > gpurun_out/c2_multigpu2.log 2>&1; echo "multigpu_check(2) rc=$? $(grep 'MULTIGPU OK' gpurun_out/c2_multigpu2.log)" >> $S
timeout 900 python bench.py > gpurun_out/c2_bench_n1.json 2> gpurun_out/c2_bench_n1.err; echo "bench n1 rc=$?" >> $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 > gpurun_out/c2_bench_n2.json 2> gpurun_out/c2_bench_n2.err; echo "bench n2 rc=$?" >> $S
tail -c 1500 gpurun_out/c2_bench_n1.err >> $S
cat $S
