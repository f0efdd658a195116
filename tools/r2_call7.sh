#!/bin/bash
# round 2, call 7 (1 GPU): K4 block layout, K1 consumer counts, K6 prefetch / batch 8
mkdir -p gpurun_out
S=gpurun_out/c7_summary.txt
: > $S
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/c7_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c7_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c7_tests.log | head -20 >> $S
for cfg in "" "SRB_K1_CONSUMERS=6" "SRB_K1_CONSUMERS=8 SRB_DENSIFY_BATCH=8" "SRB_DENSIFY_PIPE=0"; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 10 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c7_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/c7_bench.err').read()[-1500:])
PY
done
cat $S
