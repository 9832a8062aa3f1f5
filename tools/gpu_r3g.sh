#!/bin/bash
# 2 GPUs: where does the flow's +0.28 ms at N=2 come from?  exchange none / p2p back to back (extras off)
mkdir -p gpurun_out
for ex in none p2p; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --extras 0 --exchange $ex 2> gpurun_out/r3g_${ex}_err.log | tail -1 > gpurun_out/r3g_${ex}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3g_${ex}.json"))
    print("${ex}: value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["stage_ms"])
except Exception as e:
    print("${ex} failed", e)
PY
done
