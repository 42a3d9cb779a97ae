#!/bin/bash
# Round-2 data-parallel run (N GPUs of one box): in-switch all-reduce probe vs NCCL, the two-rank parity test, and the
# bench line with both exchanges. Every stage has its own time limit.
set -u
N=${1:-2}; T=${2:-a}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29611 tools/gpu_probe_switch_allreduce.py > gpurun_out/r2_switch_probe_${N}gpu_$T.json 2> gpurun_out/r2_switch_probe_${N}gpu_$T.err
echo "probe rc=$?"; tail -c 1500 gpurun_out/r2_switch_probe_${N}gpu_$T.json; tail -c 600 gpurun_out/r2_switch_probe_${N}gpu_$T.err
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_dp_gpu.py -q -x > gpurun_out/r2_tests_dp_$T.log 2>&1; echo "dp tests rc=$?"; tail -5 gpurun_out/r2_tests_dp_$T.log
fi
for X in switch nccl; do
  CADRE_ALLREDUCE=$X timeout 240 $TR --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --no-full-windows \
    > gpurun_out/r2_bench_${N}gpu_${X}_$T.json 2> gpurun_out/r2_bench_${N}gpu_${X}_$T.err
  echo "bench $X rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu_${X}_$T.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${N}gpu_${X}_$T.json").read().strip().splitlines()[-1])
    print("$X", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["phases"])
except Exception as e:
    print("no line", e)
P
done
