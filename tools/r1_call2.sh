#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary2.txt
: > $S
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_all_auto.log 2>&1; echo "all tests (default AUTO) rc=$? $(tail -1 gpurun_out/t_all_auto.log)" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke2.log | cut -c1-200)" >> $S
timeout 400 python bench.py > gpurun_out/bench_r1_lanes.json 2> gpurun_out/bench_r1_lanes.err; echo "bench rc=$?" >> $S
cat $S
