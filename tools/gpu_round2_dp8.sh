#!/bin/bash
# Round-2 8-GPU run: in-switch all-reduce probe vs NCCL, then the bench line with the switch exchange (pipelined ranges and
# single range). Every stage has its own time limit.
set -u
N=${1:-8}; T=${2:-a}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PROBE_P2P=0 PROBE_BLOCKS=16,32,64,128 timeout 200 $TR --master-port 29611 tools/gpu_probe_switch_allreduce.py > gpurun_out/r2_switch_probe_${N}gpu_$T.json 2> gpurun_out/r2_switch_probe_${N}gpu_$T.err
echo "probe rc=$?"; tail -c 2500 gpurun_out/r2_switch_probe_${N}gpu_$T.json; tail -c 400 gpurun_out/r2_switch_probe_${N}gpu_$T.err
for X in switch switch1; do
  if [ "$X" = "switch1" ]; then export CADRE_NO_ALLREDUCE_OVERLAP=1; fi
  CADRE_ALLREDUCE=switch timeout 200 $TR --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --no-full-windows \
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
