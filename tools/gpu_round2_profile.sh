#!/bin/bash
# Round-2 evidence run (1 GPU): full -m gpu suite, the default bench line, ncu launch list of one bench step,
# ncu --set full of the PPO update kernels and of one encoder forward. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
T=${1:-q}
python -m pytest tests -q -m gpu -x > gpurun_out/r2_tests_$T.log 2>&1; tail -3 gpurun_out/r2_tests_$T.log
( time python bench.py ) > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err; tail -c 600 gpurun_out/r2_bench_$T.err
( time python bench.py --impl reference ) > gpurun_out/r2_bench_ref_$T.json 2> gpurun_out/r2_bench_ref_$T.err; tail -c 300 gpurun_out/r2_bench_ref_$T.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_bench_launches_$T.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-full-windows > gpurun_out/r2_bench_under_ncu_$T.log 2>&1
timeout 900 ncu --set full --clock-control none -o gpurun_out/r2_ppo_full_$T -f python tools/ncu_ppo.py 4 > gpurun_out/r2_ncu_ppo_$T.log 2>&1
timeout 900 ncu --set full --clock-control none -o gpurun_out/r2_enc_full_$T -f python tools/ncu_encoder_u8.py 640 > gpurun_out/r2_ncu_enc_$T.log 2>&1
ls -la gpurun_out/*.ncu-rep
