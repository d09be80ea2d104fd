#!/bin/bash
set -u
mkdir -p gpurun_out
XRD_CHECK_TIMEOUT=200 timeout 900 python tools/gpu_diag.py conv3s_fp16 conv3s_cat_fp16 conv3s_stats_fp16 conv3s_gn_fp16 conv3s_gn_cat_fp16 conv3s_bf16 > gpurun_out/r2d_diag.log 2>&1
timeout 600 python tools/conv3s_time.py > gpurun_out/r2d_time.log 2>&1
cut -c1-400 gpurun_out/r2d_diag.log; cat gpurun_out/r2d_time.log
