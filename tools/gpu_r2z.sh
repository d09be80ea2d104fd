#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_conv3s|k_attn_tc" -o /tmp/r2_rep -f python tools/ncu_targets_r2.py > gpurun_out/r2z_ncu.log 2>&1
python tools/ncu_export.py /tmp/r2_rep.ncu-rep gpurun_out/r2z_hot_raw.csv >> gpurun_out/r2z_ncu.log 2>&1
ncu -i /tmp/r2_rep.ncu-rep --page source --csv > gpurun_out/r2z_src.csv 2>> gpurun_out/r2z_ncu.log
tail -3 gpurun_out/r2z_ncu.log; ls -la gpurun_out/r2z_*
