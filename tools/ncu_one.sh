#!/bin/bash
# ncu --set full of ONE kernel (regex $1) launched by script $2..., exported on the box to small CSVs (raw + source page).
# usage: bash tools/ncu_one.sh k_attn_tc tools/attn_time.py
set -u
K=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-count 1 -o /tmp/one_rep -f python "$@" > gpurun_out/ncu_one.log 2>&1
python tools/ncu_export.py /tmp/one_rep.ncu-rep gpurun_out/one_raw.csv >> gpurun_out/ncu_one.log 2>&1
ncu -i /tmp/one_rep.ncu-rep --page source --csv > gpurun_out/one_source.csv 2>> gpurun_out/ncu_one.log
ls -la gpurun_out/one_source.csv gpurun_out/one_raw.csv
