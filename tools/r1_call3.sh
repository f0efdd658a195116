#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_r1_lanes2.json 2> gpurun_out/bench_r1_lanes2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_lanes2.json').read().strip().splitlines()[-1])
print(d['e2e'])
PY
