#!/bin/bash
# round-2 first GPU visit: the full parity suite (with per-test durations), the smoke, the bench line, the reference arm with the
# one-off full-image validation of its extrapolation
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --impl reference --steps 1 --warmup 0 --full-image > gpurun_out/r2a_ref_full.json 2> gpurun_out/r2a_ref_full.err
python tools/gpu_diag.py f16_range hybrid512_fp16 hybrid512_fp32 expert_fp16 modes_512_b16 > gpurun_out/r2a_diag.log 2>&1
tail -30 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_smoke.log | tail -5; cut -c1-400 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err; cut -c1-300 gpurun_out/r2a_ref_full.json
