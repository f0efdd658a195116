#!/bin/bash
# round 2, call 8 (1 GPU): K6 bulk v2 (row-owning consumers), ncu on the reworked K1 / K4 / K6
mkdir -p gpurun_out
S=gpurun_out/c8_summary.txt
: > $S
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/c8_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c8_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c8_tests.log | head -20 >> $S
for cfg in "" "SRB_K6_BULK=0 SRB_DENSIFY_BATCH=8"; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 10 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c8_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/c8_bench.err').read()[-1500:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"major_sum_bulk|fused_exact|densify_panels" -c 3 -o gpurun_out/r02_csr python tools/ncu_step.py > gpurun_out/c8_ncu.log 2>&1; echo "ncu full rc=$?" >> $S
ncu -i gpurun_out/r02_csr.ncu-rep --page raw --csv > gpurun_out/r02_csr_raw.csv 2>> gpurun_out/c8_ncu.log
ls -la gpurun_out/r02_csr* >> $S
cat $S
