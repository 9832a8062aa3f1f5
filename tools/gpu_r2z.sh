#!/bin/bash
# round-2 call Z: FFN hidden tensor as f16 operand image (conv_2 without prologue), bias-only accumulator init before the PDL wait, trigger after the wait: timelines, subset of GPU tests, bench
mkdir -p gpurun_out
PROBE_FLOW=1 timeout 200 tests/cuda/tc_probe > gpurun_out/r2z_flow_timeline.log 2>&1; echo "flow probe exit $?"; grep -A4 "^FLOW" gpurun_out/r2z_flow_timeline.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x -k "flow or full_infer or config2 or wn_flow or config3" > gpurun_out/r2z_tests.log 2>&1; tail -4 gpurun_out/r2z_tests.log | cut -c1-300
for prec in fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2z_bench_${prec}_err.log | tail -1 > gpurun_out/r2z_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2z_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2z_bench_${prec}_err.log
done
