#!/bin/bash
# One short GPU call: the new upload / C++ host tests, the upload-mode A/B at the bench size, then (time permitting)
# the full GPU suite with the packed upload as the process default and a bench line.
mkdir -p gpurun_out
S=gpurun_out/summary.txt
: > $S
(nvidia-smi -L; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread|Core") > gpurun_out/box.txt 2>&1
timeout 300 python -m pytest tests/test_upload_gpu.py tests/test_zz_cpp_host.py -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "new tests rc=$? $(tail -1 gpurun_out/t_new.log)" >> $S
AB_REPS=2 timeout 240 python tools/e2e_ab.py > gpurun_out/e2e_ab.log 2>&1; echo "e2e_ab rc=$?" >> $S
SRB_UPLOAD_PACK=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all_pack.log 2>&1; echo "all tests (SRB_UPLOAD_PACK=1) rc=$? $(tail -1 gpurun_out/t_all_pack.log)" >> $S
timeout 400 python bench.py --upload-mode auto > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err; echo "bench rc=$?" >> $S
cat $S
