#!/bin/bash
# round 2, call 24 (1 GPU): e2e after the finer pack partition (4 parts per thread): upload alone, sequential, 2 lanes
mkdir -p gpurun_out
S=gpurun_out/c24_summary.txt
: > $S
rm -f gpurun_out/e2e_ab.jsonl
AB_MODES=balanced AB_REPS=3 timeout 300 python tools/e2e_ab.py > gpurun_out/c24_ab.log 2>&1; echo "ab rc=$?" >> $S
grep -h "balanced" gpurun_out/e2e_ab.jsonl | cut -c1-330 >> $S
for i in 1 2; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline --no-pageable --e2e-steps 3 > gpurun_out/c24_b$i.json 2> gpurun_out/c24_b$i.err; echo "bench rc=$?" >> $S
python - $i >> $S <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c24_b%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']; print('e2e', round(e['ms_per_step_sequential'],1), round(e['ms_per_step_pipelined'],1), e['h2d_bytes_per_step'], e['upload_chunks'])
PY
done
SRB_UPLOAD_THREADS=15 timeout 300 python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline --no-pageable --e2e-steps 3 > gpurun_out/c24_b3.json 2> gpurun_out/c24_b3.err; echo "bench threads=15 rc=$?" >> $S
python - 3 >> $S <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c24_b%s.json'%sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']; print('e2e', round(e['ms_per_step_sequential'],1), round(e['ms_per_step_pipelined'],1), e['h2d_bytes_per_step'], e['upload_chunks'])
PY
cat $S
