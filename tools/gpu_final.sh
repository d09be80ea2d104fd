#!/bin/bash
# Round-end evidence on one GPU box: parity suite, bench line, ncu launch lists of the sampler loop and of the whole hybrid,
# and one --set full capture of the hot kernels exported to CSV on the box (gpurun_out/ is limited to 64 MiB).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
for w in naf router hybrid ddim; do python tools/profile_step.py 16 512 2 $w; done > gpurun_out/final_step_times.log 2>&1
for w in naf hybrid ddim; do python tools/profile_step.py 1 512 8 $w; done >> gpurun_out/final_step_times.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_ddim.csv python tools/profile_step.py 16 512 2 ddim > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_hybrid.csv python tools/profile_step.py 16 512 2 hybrid > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_conv3|k_conv1|k_attn_tc|k_gn_act2" -o /tmp/hot_rep -f python tools/ncu_targets.py > gpurun_out/final_ncu_hot.log 2>&1
python tools/ncu_export.py /tmp/hot_rep.ncu-rep gpurun_out/final_ncu_hot_kernels.csv >> gpurun_out/final_ncu_hot.log 2>&1
cat gpurun_out/final_pytest.log gpurun_out/final_step_times.log; cut -c1-300 gpurun_out/final_bench.json; tail -2 gpurun_out/final_ncu_hot.log
