#!/bin/bash
# round-2: 2-GPU sanity of the scaling bench (the driver's launch line): peer-slab exchange, config4_sharded extra
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r3e_bench2_err.log | tail -1 > gpurun_out/r3e_bench2.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3e_bench2.json"))
    print("N=2 value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "exchange", d.get("exchange"), d["stage_ms"])
    print("config4", json.dumps(d.get("config4_sharded"))[:900])
except Exception as ex:
    print("bench failed", ex)
PY
tail -5 gpurun_out/r3e_bench2_err.log | cut -c1-300
