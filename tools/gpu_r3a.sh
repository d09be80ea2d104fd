#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3a}
XRD_CHECK_TIMEOUT=120 timeout 300 python tools/gpu_diag.py conv3s_gn_fp16 conv3s_gn_cat_fp16 > gpurun_out/${T}_diag.log 2>&1
if grep -q "timed out\|rror\|FAIL" gpurun_out/${T}_diag.log; then cut -c1-400 gpurun_out/${T}_diag.log | tail -20; exit 1; fi
timeout 600 python tools/conv3s_time.py > gpurun_out/${T}_time.log 2>&1
cut -c1-300 gpurun_out/${T}_diag.log; cut -c1-420 gpurun_out/${T}_time.log
