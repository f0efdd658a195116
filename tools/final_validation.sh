#!/bin/bash
# One GPU call: full GPU test-suite, smoke, the default bench (both arms) and the ncu captures kept under profiles/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 400 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 600 gpurun_out/bench_r1_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err; tail -c 400 gpurun_out/bench_r1_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv python tools/ncu_step.py > gpurun_out/ncu_list_final.log 2>&1; tail -1 gpurun_out/ncu_list_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_exact|gram_tc2_kernel|major_reduce|densify_panels|scores_tc" -c 5 -o gpurun_out/prof_r1_final python tools/ncu_step.py > gpurun_out/ncu_full_final.log 2>&1; tail -2 gpurun_out/ncu_full_final.log
