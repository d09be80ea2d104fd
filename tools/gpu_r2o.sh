#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2o}
XRD_CHECK_TIMEOUT=120 timeout 400 python tools/gpu_diag.py attention_tc_fp16 attention_tc_bf16 > gpurun_out/${T}_diag.log 2>&1
cut -c1-700 gpurun_out/${T}_diag.log
if grep -q "timed out\|rror\|FAIL" gpurun_out/${T}_diag.log; then exit 1; fi
timeout 300 python tools/attn_time.py > gpurun_out/${T}_attn_time.log 2>&1
cat gpurun_out/${T}_attn_time.log
