#!/bin/bash
# round 2, call 18 (1 GPU): K6 lazy value loads A/B, pageable e2e with always-packed values
mkdir -p gpurun_out
S=gpurun_out/c18_summary.txt
: > $S
SRB_DENSIFY_LAZY=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_edge_gpu.py tests/test_upload_balanced_gpu.py -m gpu -q -x > gpurun_out/c18_tests.log 2>&1; echo "pytest (lazy) rc=$? $(tail -1 gpurun_out/c18_tests.log)" >> $S
for cfg in "SRB_DENSIFY_LAZY=1" ""; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 10 > gpurun_out/c18_b.json 2> gpurun_out/c18_b.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c18_b.json').read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()})
PY
done
timeout 400 python bench.py --no-legs --no-cpu-baseline --steps 5 > gpurun_out/c18_bench_n1.json 2> gpurun_out/c18_bench_n1.err; echo "bench e2e rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c18_bench_n1.json').read().strip().splitlines()[-1])
print('n1', round(d['ms_per_step'],2), 'e2e', {k:v for k,v in d['e2e'].items() if k in ('ms_per_step','ms_per_step_sequential','h2d_bytes_per_step','upload_chunks','pageable_input')})
PY
cat $S
