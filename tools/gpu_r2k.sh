#!/bin/bash
# round-2 call K: unrolled MMA issue: G2 probe timings + timeline, GPU tests, bench A/B
mkdir -p gpurun_out
timeout 300 tests/cuda/g2_probe perf > gpurun_out/r2k_g2_probe.log 2>&1; echo "probe exit $?"; grep "ms \|PROBE" gpurun_out/r2k_g2_probe.log | cut -c1-60,125-250
C="tests/cuda/g2_probe case"
G2_PROF=1 $C 128 128 7 1 65472 4 20 0  16 16 7 1 523776 0 20 0  64 64 7 1 130944 7 20 1 2>&1 | tee gpurun_out/r2k_prof.log | cut -c1-230
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1; tail -8 gpurun_out/r2k_tests.log | cut -c1-300
for prec in fp16 tf32; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2k_bench_${prec}_err.log | tail -1 > gpurun_out/r2k_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2k_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2k_bench_${prec}_err.log
done
