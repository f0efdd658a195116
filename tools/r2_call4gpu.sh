#!/bin/bash
# round 2, call 15 (8 GPUs, final state): N-GPU == 1-GPU == oracle parity check at world 8 (log kept), bench at N = 8 with every leg
mkdir -p gpurun_out
S=gpurun_out/c26_summary.txt
: > $S
nvidia-smi -L | wc -l >> $S; nproc >> $S; free -g | head -2 >> $S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29681 tests/multigpu_check.py > gpurun_out/r02_multigpu4.log 2>&1; echo "multigpu_check(4) rc=$? $(grep 'MULTIGPU OK' gpurun_out/r02_multigpu4.log)" >> $S
grep -E "N vs 1 GPU|N GPUs vs oracle" gpurun_out/r02_multigpu4.log | cut -c1-400 >> $S
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 4 --steps 10 --warmup 3 --legs strong,xl --no-pageable > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench n4 rc=$?" >> $S
python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n4.json').read().strip().splitlines()[-1])
    print('n4 value %.4g ms %.2f'%(d['value'],d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms'].items()})
    for k in ('faithful','strong','xl'):
        v=d.get(k,{}); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error','cells_total')}, {a:round(b,2) for a,b in v.get('stage_ms',{}).items()})
    x=d.get('xxl',{}); print('xxl', x.get('error'), x.get('h2d_link_gbs_measured'), x.get('resident_stream'), x.get('three_pass'), x.get('check'))
    print('e2e', d.get('e2e'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02_bench_n4.err').read()[-2500:])
PY
cat $S
