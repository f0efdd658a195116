#!/bin/bash
mkdir -p gpurun_out
timeout 70 python tools/chfsi_probe.py > gpurun_out/chfsi_probe.json 2> gpurun_out/chfsi_probe.err; echo "probe rc=$?"; tail -c 1500 gpurun_out/chfsi_probe.json; tail -3 gpurun_out/chfsi_probe.err
SRB_EIG_MODE=chfsi timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "full_size or pipeline_call" 2>&1 | tail -3
