#!/bin/bash
# round 2, call 9 (1 GPU): sparse panel shift (zero-preserving panels + rank-one correction) vs centre-first; clocks during K7
mkdir -p gpurun_out
S=gpurun_out/c9_summary.txt
: > $S
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/c9_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c9_tests.log)" >> $S
grep -E "FAILED|ERROR|Error|assert" gpurun_out/c9_tests.log | head -20 >> $S
for cfg in "" "SRB_PCA_SHIFT=center"; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c9_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()}, d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/c9_bench.err').read()[-1500:])
PY
done
cat $S
