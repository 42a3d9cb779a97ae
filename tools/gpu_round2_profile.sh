#!/bin/bash
# Round-2 evidence run (1 GPU): the default bench line + reference arm, the ncu launch list of ONE bench step and an
# ncu --set full capture of ONE PPO update step. Every stage has its own time limit; outputs under gpurun_out/ (small).
set -u
mkdir -p gpurun_out
T=${1:-q}
( time timeout 200 python bench.py ) > gpurun_out/r2_bench_$T.json 2> gpurun_out/r2_bench_$T.err; tail -c 300 gpurun_out/r2_bench_$T.err
( time timeout 200 python bench.py --impl reference ) > gpurun_out/r2_bench_ref_$T.json 2> gpurun_out/r2_bench_ref_$T.err; tail -c 200 gpurun_out/r2_bench_ref_$T.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_bench_launches_$T.csv \
  python bench.py --profile-step --warmup 2 --no-cpu-baseline --no-eager-baseline --no-full-windows > gpurun_out/r2_bench_under_ncu_$T.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_bench_launches_$T.csv)"
timeout 300 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2_ppo_full_$T -f \
  python tools/ncu_ppo.py 4 > gpurun_out/r2_ncu_ppo_$T.log 2>&1
echo "ppo full rc=$?"; ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
