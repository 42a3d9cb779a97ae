#!/bin/bash
# Round-2 4-GPU run: default bench line (exchange auto -> in-switch kernel, prefetch on) and the same without prefetch.
set -u
N=${1:-4}; T=${2:-a}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for X in default noprefetch; do
  if [ "$X" = "noprefetch" ]; then export CADRE_PREFETCH=0; fi
  timeout 200 $TR --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --no-full-windows \
    > gpurun_out/r2_bench_${N}gpu_${X}_$T.json 2> gpurun_out/r2_bench_${N}gpu_${X}_$T.err
  echo "bench $X rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu_${X}_$T.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_${N}gpu_${X}_$T.json").read().strip().splitlines()[-1])
    print("$X", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["phases"])
except Exception as e:
    print("no line", e)
P
done
