#!/bin/bash
# First GPU call of round 2 (~4 min of box time): verify everything round 1 could only build, then the tunable sweeps.
#   1. full GPU suite at defaults (includes the tests added after the last GPU run of round 1: test_zzz1..5; SRB_TEST_PENDING=1 un-gates them)
#   2. upload modes A/B at the bench size, incl. the unmeasured HOST_PACK_ADAPTIVE and HOST_PACK_DELTA
#   3. K8 tunables (each setting needs its own process: they are read once)
#   4. bench at defaults
mkdir -p gpurun_out
S=gpurun_out/r2_summary.txt
: > $S
SRB_TEST_PENDING=1 timeout 600 python -m pytest tests -m gpu -q --durations=30 --timeout=180 > gpurun_out/r2_tests.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r2_tests.log)" >> $S
grep -E "FAILED|ERROR" gpurun_out/r2_tests.log | head -20 >> $S
rm -f gpurun_out/e2e_ab.jsonl
AB_REPS=2 timeout 300 python tools/e2e_ab.py > gpurun_out/r2_e2e_ab.log 2>&1; echo "e2e_ab rc=$?" >> $S
grep -E "'what': '(upload|e2e)'" gpurun_out/r2_e2e_ab.log | cut -c1-170 >> $S
for cfg in "" "SRB_CHFSI_KRYLOV=16" "SRB_CHFSI_KRYLOV=24" "SRB_CHFSI_TARGET=10" "SRB_CHFSI_KRYLOV=16 SRB_CHFSI_TARGET=10" \
           "SRB_CHFSI_BLOCK=128" "SRB_CHFSI_BLOCK=256" "SRB_CHFSI_KRYLOV=16 SRB_CHFSI_TARGET=10 SRB_CHFSI_BLOCK=256"; do
  echo "== $cfg" >> $S
  env $cfg timeout 120 python tools/chfsi_probe.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chfsi eig ms', d['chfsi']['eig_ms'], d['chfsi']['info'], 'comp diff %.1e' % d['components_max_abs_diff'])" >> $S 2>&1
done
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> $S
cat $S
