#!/bin/bash
# round-2 call W: compile-time (NK, MG) issuer in k_g2_conv (straight-line MMA issue per stage): probe, phase accounting, subset of GPU tests, bench
mkdir -p gpurun_out
timeout 300 tests/cuda/g2_probe perf > gpurun_out/r2w_g2_probe.log 2>&1; echo "probe exit $?"; grep -c PASS gpurun_out/r2w_g2_probe.log; grep "FAIL\|PROBE\|error" gpurun_out/r2w_g2_probe.log | head
C="tests/cuda/g2_probe case"
G2_PROF=1 timeout 200 $C 256 256 11 1 8184 0 20 1  256 256 3 1 8184 0 20 1  128 128 11 1 32736 0 20 1  128 128 3 1 32736 0 20 1  64 64 11 1 130944 0 20 1  32 32 11 1 261888 0 20 1  16 16 11 1 523776 0 20 1  16 16 3 1 523776 0 20 1 > gpurun_out/r2w_prof.log 2>&1
grep "PASS\|FAIL\|cta    0" gpurun_out/r2w_prof.log | cut -c1-60,100-330
timeout 600 python -m pytest tests -m gpu -q -x -k "generator or full_infer or config2 or flow_stage" > gpurun_out/r2w_tests.log 2>&1; tail -4 gpurun_out/r2w_tests.log | cut -c1-300
for prec in fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2w_bench_${prec}_err.log | tail -1 > gpurun_out/r2w_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2w_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2w_bench_${prec}_err.log
done
