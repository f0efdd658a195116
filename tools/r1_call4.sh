#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/summary4.txt
: > $S
timeout 300 python -m pytest tests/test_upload_gpu.py tests/test_zz_cpp_host.py -m gpu -x -q > gpurun_out/t_new4.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/t_new4.log)" >> $S
rm -f gpurun_out/e2e_ab.jsonl
AB_REPS=2 timeout 240 python tools/e2e_ab.py > gpurun_out/e2e_ab4.log 2>&1; echo "e2e_ab rc=$?" >> $S
grep -E "'what': '(upload|e2e)'" gpurun_out/e2e_ab4.log | cut -c1-160 >> $S
timeout 400 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_r1_vpack.json 2> gpurun_out/bench_r1_vpack.err; echo "bench rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_vpack.json').read().strip().splitlines()[-1])
print(d['e2e'])
PY
cat $S
