#!/bin/bash
# round 2, call 25 (1 GPU): parts per thread of the host pack regions: 1 / 2 / 4 in the same call (upload alone, e2e sequential, 2 lanes)
mkdir -p gpurun_out
S=gpurun_out/c25_summary.txt
: > $S
for ppt in 1 4 2 1; do
echo "== parts per thread $ppt" >> $S
rm -f gpurun_out/e2e_ab.jsonl
SRB_PACK_PARTS_PER_THREAD=$ppt AB_MODES=balanced AB_REPS=2 timeout 300 python tools/e2e_ab.py > gpurun_out/c25_ab.log 2>&1; echo "ab rc=$?" >> $S
grep -h "balanced" gpurun_out/e2e_ab.jsonl | cut -c1-130 >> $S
SRB_PACK_PARTS_PER_THREAD=$ppt timeout 300 python bench.py --steps 3 --warmup 3 --no-legs --no-cpu-baseline --no-pageable --e2e-steps 3 > gpurun_out/c25_b.json 2> gpurun_out/c25_b.err; echo "bench rc=$?" >> $S
python - >> $S <<'PY'
import json,sys
d=json.loads(open('gpurun_out/c25_b.json').read().strip().splitlines()[-1]); e=d['e2e']; print('e2e seq', round(e['ms_per_step_sequential'],1), 'lanes', round(e['ms_per_step_pipelined'],1), e['h2d_bytes_per_step'], e['upload_chunks'])
PY
done
cat $S
