#!/bin/bash
# round 2, call 5 (1 GPU): bulk-staged K1 / K6 (cp.async.bulk), K8 at b = 160, A/B against the previous kernels, ncu evidence
mkdir -p gpurun_out
S=gpurun_out/c5_summary.txt
: > $S
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/c5_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c5_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c5_tests.log | head -20 >> $S
for cfg in "" "SRB_K1_BULK=0 SRB_K6_BULK=0"; do
  echo "== $cfg" >> $S
  env $cfg timeout 300 python bench.py --no-legs --no-e2e --no-cpu-baseline --steps 10 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?" >> $S
  python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c5_bench.json').read().strip().splitlines()[-1])
    print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, {k:round(v['frac'],3) for k,v in d['rooflines'].items()}, d['eig_solver'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/c5_bench.err').read()[-1500:])
PY
done
cp gpurun_out/c5_bench.json gpurun_out/c5_bench_oldkernels.json
# ncu: every launch of one step (shares), then --set full on the hot kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/ncu_step.py > gpurun_out/c5_ncu1.log 2>&1; echo "ncu launches rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"major_sum_bulk|fused_exact|densify_panels_bulk|gram_tc2_kernel|scores_tc_kernel" -c 5 -o gpurun_out/r02_full python tools/ncu_step.py > gpurun_out/c5_ncu2.log 2>&1; echo "ncu full rc=$?" >> $S
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>> gpurun_out/c5_ncu2.log
ls -la gpurun_out/r02_full* >> $S
cat $S
