#!/bin/bash
# round 2, call 20 (1 GPU): K7 with N trimmed in the last tile column
mkdir -p gpurun_out
S=gpurun_out/c20_summary.txt
: > $S
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_edge_gpu.py tests/test_eig_gpu.py tests/test_api_mirror_gpu.py tests/test_out_of_core_gpu.py tests/test_store_gpu.py -m gpu -q -x > gpurun_out/c20_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c20_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c20_tests.log | head >> $S
for cfg in "" "SRB_GRAM_TRIM=0"; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/c20_b.json 2> gpurun_out/c20_b.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c20_b.json').read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()})
PY
done
cat $S
