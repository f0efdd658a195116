#!/bin/bash
# SASS mnemonics per hot kernel of the in-tree library (B200_PROFILING.md: UTCHMMA / UTMALDG / LDTM prove tcgen05 + TMA,
# UBLKCP proves cp.async.bulk, ATOMS the native shared-memory integer atomics). Usage: tools/sass_evidence.sh > profiles/r02_sass.txt
SO=singlerust_b200/libsrb200.so
echo "# cuobjdump -sass $SO  ($(date -u +%F), $(sha256sum $SO | cut -c1-16))"
cuobjdump -sass $SO | awk '
/Function :/ { fn=$3 }
{ if (match($0, /UTCHMMA(\.2CTA)?|UTMALDG(\.2D)?(\.2CTA)?|UBLKCP(\.S\.G)?|LDTM|UTCBAR|SYNCS\.ARRIVE\.TRANS64|ATOMS\.ADD|REDS|RED\.E\.ADD\.(64|F64)|ATOMG|DMMA|VIMNMX3/)) { k=substr($0, RSTART, RLENGTH); c[fn" "k]++ } }
END { for (x in c) print c[x], x }' | sort -k2,2 -k3,3 | awk '{printf "%6d  %-28s %s\n", $1, $3, $2}' | grep -E "gram_tc|scores_tc|major_sum_bulk|fused_exact_kernelIffLb1|densify_panels_pipe_kernelIfLi8|symv_cols|lanczos_orth|delta_decode|minor_minmax" 
