#!/bin/bash
# Final round-2 evidence on one GPU box (everything lands in gpurun_out/r02f_*): GPU parity suite, bench line, step times, measured
# parity table, ncu launch lists of the sampler loop and of bench.py's timed region.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02f_gpu_pytest.log
python bench.py --steps 4 --warmup 3 > $O/r02f_bench.json 2> $O/r02f_bench.err
for w in naf router hybrid ddim; do python tools/profile_step.py 16 512 2 $w; done > $O/r02f_step_times.txt 2>&1
for w in naf hybrid ddim; do python tools/profile_step.py 1 512 8 $w; done >> $O/r02f_step_times.txt 2>&1
XRD_OVERLAP=0 python tools/profile_step.py 16 512 50 hybrid >> $O/r02f_step_times.txt 2>&1
python tools/profile_step.py 16 512 50 hybrid >> $O/r02f_step_times.txt 2>&1
python tools/gpu_diag.py hybrid512_fp32 hybrid512_fp16 hybrid512_bf16 unet256_fp32 unet256_fp16 modes_512_b16 unet_teacher_fp32 unet_teacher_fp16 unet_teacher_bf16 hybrid_fp32 hybrid_fp16 hybrid_bf16 attention_tc_fp16 expert_fp32 expert_fp16 f16_range ref_tf32_noise first_conv_fp16 first_conv_stats_fp16 > $O/r02f_diag.log 2>&1
cp $O/diag.json $O/r02f_parity_measured.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02f_launches_unet_eval_b16.csv python tools/profile_step.py 16 512 2 ddim > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 24150 -c 1600 --csv --log-file $O/r02f_launches_bench_timed_region.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > $O/r02f_launches_bench.log 2>&1
cat $O/r02f_gpu_pytest.log $O/r02f_step_times.txt; cut -c1-600 $O/r02f_bench.json; tail -3 $O/r02f_diag.log
