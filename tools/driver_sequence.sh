#!/bin/bash
# What the driver runs at round end on one GPU, in its order: smoke(), python bench.py (defaults), python bench.py --impl reference
mkdir -p gpurun_out
S=gpurun_out/final_sanity.txt
: > $S
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$? $(grep 'smoke OK' gpurun_out/final_smoke.log | cut -c1-200)" >> $S
grep real gpurun_out/final_smoke.log >> $S
( time timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err ) 2>> $S; echo "bench default rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_default.json').read().strip().splitlines()[-1])
print(d['metric'], '%.4g'%d['value'], d['unit'], 'ms', round(d['ms_per_step'],2), 'steps', d['steps'], d['warmup'], 'launches', d['gpu_launches'], d['clocks'])
print('roofline', d['roofline']); print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step','d2h_bytes_per_step')}); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print({k:round(v['frac'],3) for k,v in d['rooflines'].items()})
PY
( time timeout 600 python bench.py --impl reference ) > gpurun_out/final_bench_reference.json 2>> $S; echo "reference rc=$?" >> $S
cut -c1-250 gpurun_out/final_bench_reference.json >> $S
cat $S
