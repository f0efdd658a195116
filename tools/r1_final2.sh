#!/bin/bash
# Final round-1 validation at defaults: full GPU suite, smoke, both bench arms.
mkdir -p gpurun_out
S=gpurun_out/summary_final2.txt
: > $S
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_final2.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/t_final2.log)" >> $S
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final2.log | cut -c1-120)" >> $S
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final2_reference.json 2> gpurun_out/bench_final2_reference.err; echo "bench reference rc=$?" >> $S
timeout 400 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?" >> $S
cat $S
