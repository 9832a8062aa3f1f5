#!/bin/bash
# round-2 call P (session 3 status): GPU tests, smoke, fp16 bench (headline only), launch list of the fp16 build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_tests.log 2>&1; tail -4 gpurun_out/r2p_tests.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -1 gpurun_out/r2p_smoke.log
for prec in fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2p_bench_${prec}_err.log | tail -1 > gpurun_out/r2p_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2p_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2p_bench_${prec}_err.log
done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2p_launches_fp16.csv python tools/profile_step.py --steps 3 --precision fp16 > gpurun_out/r2p_ncu.log 2>&1
tail -3 gpurun_out/r2p_ncu.log
