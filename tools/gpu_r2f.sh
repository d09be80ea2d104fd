#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_conv3s" --launch-count 4 -o /tmp/c3s_rep -f python tools/ncu_targets_c3s.py > gpurun_out/r2f_ncu.log 2>&1
python tools/ncu_export.py /tmp/c3s_rep.ncu-rep gpurun_out/r2f_c3s_raw.csv >> gpurun_out/r2f_ncu.log 2>&1
ncu -i /tmp/c3s_rep.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/r2f_src1.csv 2>> gpurun_out/r2f_ncu.log
tail -2 gpurun_out/r2f_ncu.log
