#!/bin/bash
# round 2, call 3 (1 GPU): new tests, K8 phase probe + spectrum dump, upload A/B with the balanced mode, bench
mkdir -p gpurun_out
S=gpurun_out/c3_summary.txt
: > $S
nproc >> $S
timeout 300 python -m pytest tests/test_upload_balanced_gpu.py tests/test_edge_gpu.py tests/test_eig_gpu.py tests/test_upload_delta_gpu.py tests/test_store_gpu.py tests/test_out_of_core_gpu.py -m gpu -q --durations=8 --timeout=120 > gpurun_out/c3_tests.log 2>&1; echo "pytest new rc=$? $(tail -1 gpurun_out/c3_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c3_tests.log | head -20 >> $S
timeout 200 python tools/eig_phases.py > gpurun_out/c3_eig.json 2> gpurun_out/c3_eig.err; echo "eig_phases rc=$?" >> $S
grep "chfsi phases\|spectrum" gpurun_out/c3_eig.err >> $S; cat gpurun_out/c3_eig.json >> $S
rm -f gpurun_out/e2e_ab.jsonl
AB_REPS=2 AB_MODES=host_pack,host_pack_delta,host_pack_adaptive,balanced timeout 300 python tools/e2e_ab.py > gpurun_out/c3_e2e_ab.log 2>&1; echo "e2e_ab rc=$?" >> $S
grep -E "'what': '(upload|e2e)'" gpurun_out/c3_e2e_ab.log | cut -c1-260 >> $S
timeout 600 python bench.py --legs faithful > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?" >> $S
cat $S
