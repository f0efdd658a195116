#!/bin/bash
# round 2, call 13 (2 GPUs): 2-GPU parity (pytest's torchrun test + the check itself, log kept), smoke(), bench N = 2 with legs
mkdir -p gpurun_out
S=gpurun_out/c13_summary.txt
: > $S
timeout 300 python __graft_entry__.py --smoke > gpurun_out/c13_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/c13_smoke.log | cut -c1-200)" >> $S
timeout 600 python -m pytest tests/test_api_mirror_gpu.py -m gpu -q -x --durations=5 > gpurun_out/c13_tests.log 2>&1; echo "pytest api_mirror rc=$? $(tail -1 gpurun_out/c13_tests.log)" >> $S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29691 tests/multigpu_check.py > gpurun_out/r02_multigpu2.log 2>&1; echo "multigpu_check(2) rc=$? $(grep 'MULTIGPU OK' gpurun_out/r02_multigpu2.log)" >> $S
grep -E "N vs 1 GPU|N GPUs vs oracle" gpurun_out/r02_multigpu2.log | cut -c1-420 >> $S
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29692 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?" >> $S
python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
    print('n2 value %.4g ms %.2f'%(d['value'],d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['eig_solver'])
    for k in ('faithful','strong','xl'):
        v=d.get(k,{}); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error','cells_total')})
    x=d.get('xxl',{}); print('xxl', x.get('error'), x.get('resident_stream'), x.get('three_pass'), x.get('check'))
    print('e2e', d.get('e2e'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02_bench_n2.err').read()[-2500:])
PY
cat $S
