#!/bin/bash
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_conv3s" --launch-count 1 -o /tmp/c3s_rep -f python tools/ncu_targets_c3s.py > gpurun_out/r2h_ncu.log 2>&1
ncu -i /tmp/c3s_rep.ncu-rep --page source --csv > gpurun_out/r2h_src1.csv 2>> gpurun_out/r2h_ncu.log
ncu -i /tmp/c3s_rep.ncu-rep --page details --csv > gpurun_out/r2h_details.csv 2>> gpurun_out/r2h_ncu.log
tail -2 gpurun_out/r2h_ncu.log
