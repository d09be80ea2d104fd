#!/bin/bash
# Round-2 evidence on one GPU box (everything lands in gpurun_out/r02_*): bench line, step times, tiled-1024 line (configs[4], one
# GPU's share), measured parity table, ncu launch lists, one --set full capture of the round-2 hot kernels exported on the box.
set -u
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 4 --warmup 3 > $O/r02_bench.json 2> $O/r02_bench.err
for w in naf router hybrid ddim; do python tools/profile_step.py 16 512 2 $w; done > $O/r02_step_times.txt 2>&1
for w in naf hybrid ddim; do python tools/profile_step.py 1 512 8 $w; done >> $O/r02_step_times.txt 2>&1
python bench.py --workload tiled1024 --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > $O/r02_bench_tiled1024_n1.json 2> $O/r02_bench_tiled1024_n1.err
python tools/gpu_diag.py hybrid512_fp32 hybrid512_fp16 hybrid512_bf16 unet256_fp32 unet256_fp16 modes_512_b16 unet_teacher_fp32 unet_teacher_fp16 unet_teacher_bf16 hybrid_fp32 hybrid_fp16 hybrid_bf16 attention_tc_fp16 expert_fp32 expert_fp16 f16_range ref_tf32_noise > $O/r02_diag.log 2>&1
cp $O/diag.json $O/r02_parity_measured.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_unet_eval_b16.csv python tools/profile_step.py 16 512 2 ddim > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 24300 -c 1600 --csv --log-file $O/r02_launches_bench_timed_region.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > $O/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_conv3s|k_attn_tc" -o /tmp/r02_rep -f python tools/ncu_targets_r2.py > $O/r02_ncu.log 2>&1
python tools/ncu_export.py /tmp/r02_rep.ncu-rep $O/r02_ncu_hot_kernels.csv >> $O/r02_ncu.log 2>&1
python tools/ncu_traffic.py $O/r02_ncu_hot_kernels.csv 'k_conv3s<__half, 3, 0, 8, 1, 1>' 16 512 $O/r02_top_kernel_traffic.json >> $O/r02_ncu.log 2>&1
cat $O/r02_step_times.txt; cut -c1-400 $O/r02_bench.json; cut -c1-400 $O/r02_bench_tiled1024_n1.json; tail -3 $O/r02_bench_tiled1024_n1.err; tail -3 $O/r02_ncu.log; tail -3 $O/r02_diag.log
