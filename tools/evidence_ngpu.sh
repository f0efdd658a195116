#!/bin/bash
# N-GPU evidence (profiles/r02_multigpu{2,4,8}.log, r02_bench_n{2,4,8}.json): `gpurun --gpus N -- 'bash tools/evidence_ngpu.sh N'`
# runs tests/multigpu_check.py (N-GPU == 1-GPU == CPU oracle, log kept) and the bench at N GPUs (weak + strong + xl legs + e2e).
N=${1:-8}
mkdir -p gpurun_out
S=gpurun_out/ngpu_summary.txt
: > $S
nvidia-smi -L | wc -l >> $S; nproc >> $S; free -g | head -2 >> $S
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29681 tests/multigpu_check.py > gpurun_out/r02_multigpu$N.log 2>&1; echo "multigpu_check($N) rc=$? $(grep 'MULTIGPU OK' gpurun_out/r02_multigpu$N.log)" >> $S
grep -E "N vs 1 GPU|N GPUs vs oracle" gpurun_out/r02_multigpu$N.log | cut -c1-400 >> $S
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus $N --steps 10 --warmup 3 --legs strong,xl --no-pageable > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc=$?" >> $S
python - $N >> $S <<'PY'
import json, sys
try:
    d=json.loads(open('gpurun_out/r02_bench_n'+sys.argv[1]+'.json').read().strip().splitlines()[-1])
    print('value %.4g ms %.2f'%(d['value'],d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms'].items()})
    for k in ('faithful','strong','xl'):
        v=d.get(k,{}); print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error','cells_total')}, {a:round(b,2) for a,b in v.get('stage_ms',{}).items()})
    x=d.get('xxl',{}); print('xxl', x.get('error'), x.get('h2d_link_gbs_measured'), x.get('resident_stream'), x.get('three_pass'), x.get('check'))
    print('e2e', d.get('e2e'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r02_bench_n'+sys.argv[1]+'.err').read()[-2500:])
PY
cat $S
