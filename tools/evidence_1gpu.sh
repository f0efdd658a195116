#!/bin/bash
# 1-GPU evidence (profiles/r02_gputests.log, r02_bench_n1.json, r02_launches.csv, r02_full_raw.csv): `gpurun -- 'bash tools/evidence_1gpu.sh'`
# runs the GPU test suite, the full bench line, one e2e run with 3 steps in flight, the ncu launch list of one step and the
# ncu --set full capture of the five hot kernels (raw page exported as CSV; profiles/ncu_traffic_L.json is made from it).
mkdir -p gpurun_out
S=gpurun_out/ev1_summary.txt
: > $S
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gputests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_gputests.log)" >> $S
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench full rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('FULL', round(d['ms_per_step'],2), '%.4g'%d['value'], {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()}, d['eig_solver'], d['clocks'], 'launches', d['gpu_launches'])
for k in ('faithful','pipelined','strong','xl'):
    v=d.get(k,{}); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error')})
x=d.get('xxl',{}); print('xxl', x.get('error'), x.get('h2d_link_gbs_measured'), x.get('resident_stream'), x.get('three_pass'), x.get('check'))
print('e2e', d.get('e2e')); print('cpu', d.get('cpu_baseline')); print('roofline', d.get('roofline'))
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline --no-pageable --e2e-inflight 3 --e2e-steps 3 > gpurun_out/ev1_e2e3.json 2> gpurun_out/ev1_e2e3.err; echo "e2e inflight 3 rc=$?" >> $S
python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/ev1_e2e3.json').read().strip().splitlines()[-1]); print('e2e x3', d.get('e2e'))
except Exception as e: print('e2e3 parse failed', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/ncu_step.py > gpurun_out/ev1_ncu1.log 2>&1; echo "ncu launches rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"major_sum_bulk|fused_exact|densify_panels|gram_tc2_kernel|scores_tc_kernel" -c 5 -o gpurun_out/r02_full python tools/ncu_step.py > gpurun_out/ev1_ncu2.log 2>&1; echo "ncu full rc=$?" >> $S
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>> gpurun_out/ev1_ncu2.log
rm -f gpurun_out/r02_full.ncu-rep
cat $S
