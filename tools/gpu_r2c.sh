#!/bin/bash
# ncu --set full of the conv3s launches (raw metrics for all, source-level stall sites for the first two)
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_conv3s" --launch-count 4 -o /tmp/c3s_rep -f python tools/ncu_targets_c3s.py > gpurun_out/r2c_ncu.log 2>&1
python tools/ncu_export.py /tmp/c3s_rep.ncu-rep gpurun_out/r2c_c3s_raw.csv >> gpurun_out/r2c_ncu.log 2>&1
ncu -i /tmp/c3s_rep.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/r2c_src1.csv 2>> gpurun_out/r2c_ncu.log
ncu -i /tmp/c3s_rep.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/r2c_src2.csv 2>> gpurun_out/r2c_ncu.log
ncu -i /tmp/c3s_rep.ncu-rep --page details --csv > gpurun_out/r2c_details.csv 2>> gpurun_out/r2c_ncu.log
python tools/ncu_source_top.py gpurun_out/r2c_src1.csv 40 > gpurun_out/r2c_src1_top.txt 2>&1
python tools/ncu_source_top.py gpurun_out/r2c_src2.csv 40 > gpurun_out/r2c_src2_top.txt 2>&1
python tools/gpu_diag.py ref_tf32_noise modes_512_b16 > gpurun_out/r2c_diag.log 2>&1
tail -3 gpurun_out/r2c_ncu.log; cat gpurun_out/r2c_c3s_raw.csv | cut -c1-1500; cat gpurun_out/r2c_src1_top.txt; cat gpurun_out/r2c_diag.log
