#!/bin/bash
# round-2 call X: key-split fused attention (cluster + DSMEM merge): probe (KS = 1/2/4, ragged lengths), all GPU tests, bench, launch list
mkdir -p gpurun_out
PROBE_ATTN=1 timeout 300 tests/cuda/tc_probe perf > gpurun_out/r2x_attn_probe.log 2>&1; echo "attn probe exit $?"; grep "ATTN" gpurun_out/r2x_attn_probe.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2x_tests.log 2>&1; tail -4 gpurun_out/r2x_tests.log | cut -c1-300
for prec in fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2x_bench_${prec}_err.log | tail -1 > gpurun_out/r2x_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2x_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2x_bench_${prec}_err.log
done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2x_launches_fp16.csv python tools/profile_step.py --steps 3 --precision fp16 > gpurun_out/r2x_ncu.log 2>&1
tail -2 gpurun_out/r2x_ncu.log
