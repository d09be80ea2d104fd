#!/bin/bash
set -u
mkdir -p gpurun_out
XRD_CHECK_TIMEOUT=200 timeout 600 python tools/gpu_diag.py conv3s_fp16 conv3s_cat_fp16 conv3s_stats_fp16 conv3s_cat_stats_fp16 conv3s_gn_fp16 conv3s_gn_cat_fp16 > gpurun_out/r2g_diag.log 2>&1
timeout 600 python tools/conv3s_time.py > gpurun_out/r2g_time.log 2>&1
ncu --set full --clock-control none -k regex:"k_conv3s" --launch-count 4 -o /tmp/c3s_rep -f python tools/ncu_targets_c3s.py > gpurun_out/r2g_ncu.log 2>&1
python tools/ncu_export.py /tmp/c3s_rep.ncu-rep gpurun_out/r2g_c3s_raw.csv >> gpurun_out/r2g_ncu.log 2>&1
cut -c1-500 gpurun_out/r2g_diag.log; cut -c1-420 gpurun_out/r2g_time.log
