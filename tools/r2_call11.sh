#!/bin/bash
# round 2, call 11 (1 GPU): spectra + ChFSI behaviour at 2M / 4M cells (why does K8 need a second outer round at 8M cells?)
mkdir -p gpurun_out
S=gpurun_out/c11_summary.txt
: > $S
for cells in 2000000 4000000; do
  echo "== cells $cells" >> $S
  CELLS=$cells timeout 300 python tools/eig_phases.py > gpurun_out/c11_eig_$cells.json 2> gpurun_out/c11_eig_$cells.err; echo "rc=$?" >> $S
  mv gpurun_out/bench_spectrum_f64.bin gpurun_out/spectrum_${cells}_f64.bin
  grep "chfsi phases\|spectrum" gpurun_out/c11_eig_$cells.err | tail -3 >> $S; cat gpurun_out/c11_eig_$cells.json >> $S
done
cat $S
