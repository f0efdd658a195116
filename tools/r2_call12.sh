#!/bin/bash
# round 2, call 12 (1 GPU): K9 4-chunk accumulation, K8 top-up, full bench with every leg; Gram chunk length A/B
mkdir -p gpurun_out
S=gpurun_out/c12_summary.txt
: > $S
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/c12_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c12_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c12_tests.log | head -20 >> $S
SRB_GRAM_CHUNK=8 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_eig_gpu.py tests/test_edge_gpu.py tests/test_api_mirror_gpu.py -m gpu -q -x -s > gpurun_out/c12_tests_chunk8.log 2>&1; echo "pytest chunk8 rc=$? $(tail -1 gpurun_out/c12_tests_chunk8.log)" >> $S
grep -E "component max-abs|eigen-residuals" gpurun_out/c12_tests_chunk8.log | cut -c1-300 | head -6 >> $S
for cfg in "SRB_GRAM_CHUNK=8" "SRB_GRAM_CHUNK=4"; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 10 > gpurun_out/c12_b.json 2> gpurun_out/c12_b.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c12_b.json').read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()})
PY
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench full rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c12_bench.json').read().strip().splitlines()[-1])
print('FULL', round(d['ms_per_step'],2), '%.4g'%d['value'], {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()}, d['eig_solver'], d['clocks'])
for k in ('faithful','pipelined','strong','xl'):
    v=d.get(k,{}); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error')})
x=d.get('xxl',{}); print('xxl', x.get('error'), x.get('h2d_link_gbs_measured'), x.get('resident_stream'), x.get('three_pass'), x.get('check'))
print('e2e', d.get('e2e')); print('cpu', d.get('cpu_baseline',{}).get('value'))
PY
cat $S
