#!/bin/bash
# round 2, call 17 (2 GPUs): K8 estimate fixes on the 2M-cell matrix, measured link rate in the balanced upload
mkdir -p gpurun_out
S=gpurun_out/c17_summary.txt
: > $S
timeout 400 python -m pytest tests/test_upload_balanced_gpu.py tests/test_upload_gpu.py tests/test_upload_delta_gpu.py tests/test_eig_gpu.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/c17_tests.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c17_tests.log)" >> $S
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/multigpu_check.py > gpurun_out/c17_multigpu2.log 2>&1; echo "multigpu_check(2) rc=$? $(grep 'MULTIGPU OK' gpurun_out/c17_multigpu2.log)" >> $S
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --legs strong,xl --no-cpu-baseline > gpurun_out/c17_bench_n2.json 2> gpurun_out/c17_bench_n2.err; echo "bench n2 rc=$?" >> $S
python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c17_bench_n2.json').read().strip().splitlines()[-1])
    print('n2 value %.4g ms %.2f'%(d['value'],d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['eig_solver'])
    for k in ('strong','xl'):
        v=d.get(k,{}); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error')}, {a:round(b,2) for a,b in v.get('stage_ms',{}).items()})
    print('e2e', d.get('e2e'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/c17_bench_n2.err').read()[-2500:])
PY
timeout 300 python bench.py --no-legs --no-cpu-baseline --steps 5 > gpurun_out/c17_bench_n1.json 2> gpurun_out/c17_bench_n1.err; echo "bench n1 rc=$?" >> $S
python - >> $S <<'PY'
import json
d=json.loads(open('gpurun_out/c17_bench_n1.json').read().strip().splitlines()[-1])
print('n1', round(d['ms_per_step'],2), d['eig_solver'], 'e2e', {k:v for k,v in d['e2e'].items() if k in ('ms_per_step','ms_per_step_sequential','h2d_bytes_per_step','upload_chunks','pageable_input')})
PY
cat $S
