#!/bin/bash
# round 2, call 16 (2 GPUs): ChFSI filter columns sharded over the ranks
mkdir -p gpurun_out
S=gpurun_out/c16_summary.txt
: > $S
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tests/multigpu_check.py > gpurun_out/c16_multigpu2.log 2>&1; echo "multigpu_check(2) rc=$? $(grep 'MULTIGPU OK' gpurun_out/c16_multigpu2.log)" >> $S
grep -E "N vs 1 GPU|N GPUs vs oracle|Error|error" gpurun_out/c16_multigpu2.log | cut -c1-300 >> $S
for cfg in "" "SRB_CHFSI_SHARD=0"; do
echo "== $cfg" >> $S
env $cfg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 10 --warmup 3 --legs strong --no-e2e --no-cpu-baseline > gpurun_out/c16_bench_n2.json 2> gpurun_out/c16_bench_n2.err; echo "bench n2 rc=$?" >> $S
python - >> $S <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c16_bench_n2.json').read().strip().splitlines()[-1])
    print('n2 value %.4g ms %.2f'%(d['value'],d['ms_per_step']), {k:round(v,2) for k,v in d['stage_ms'].items()}, d['eig_solver'])
    v=d.get('strong',{}); print('strong', {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','error')}, {a:round(b,2) for a,b in v.get('stage_ms',{}).items()})
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/c16_bench_n2.err').read()[-2500:])
PY
done
cat $S
