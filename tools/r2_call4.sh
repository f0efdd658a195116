#!/bin/bash
# round 2, call 4 (1 GPU): K8 rework (own Lanczos / Cholesky, trimmed degree), balanced upload v2, bench
mkdir -p gpurun_out
S=gpurun_out/c4_summary.txt
: > $S
timeout 400 python -m pytest tests/test_eig_gpu.py tests/test_upload_balanced_gpu.py tests/test_gpu_parity.py tests/test_out_of_core_gpu.py -m gpu -q --durations=5 --timeout=120 > gpurun_out/c4_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c4_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/c4_tests.log | head -20 >> $S
for cfg in "" "SRB_CHFSI_RR=syevj" "SRB_CHFSI_BLOCK=128" "SRB_CHFSI_BLOCK=160"; do
  echo "== $cfg" >> $S
  env $cfg DUMP=0 timeout 200 python tools/eig_phases.py > gpurun_out/c4_eig.json 2> gpurun_out/c4_eig.err; echo "eig_phases rc=$?" >> $S
  grep "chfsi phases" gpurun_out/c4_eig.err | tail -1 >> $S; cat gpurun_out/c4_eig.json >> $S
done
rm -f gpurun_out/e2e_ab.jsonl
AB_REPS=2 AB_MODES=host_pack_delta,balanced timeout 300 python tools/e2e_ab.py > gpurun_out/c4_e2e_ab.log 2>&1; echo "e2e_ab rc=$?" >> $S
grep -E "'what': '(upload|e2e)'" gpurun_out/c4_e2e_ab.log | cut -c1-260 >> $S
timeout 600 python bench.py --no-legs > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c4_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['eig_solver'])
print('e2e', d['e2e'])
PY
cat $S
