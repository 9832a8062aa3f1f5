#!/bin/bash
# round-2 call A: hardware probe of the FP16-operand conv family + existing GPU tests + A/B bench (tf32 vs fp16 engine)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -1
timeout 420 tests/cuda/tc_probe perf > gpurun_out/r2a_probe.log 2>&1; echo "probe exit $?"; grep -c PASS gpurun_out/r2a_probe.log; grep -E "FAIL|error|TIMEOUT|PROBE|MN-major" gpurun_out/r2a_probe.log | head -40
PROBE_ATTN=1 timeout 120 tests/cuda/tc_probe perf > gpurun_out/r2a_attn1.log 2>&1; echo "attn(mn=1) exit $?"; grep -E "ATTN|MN-major|error|TIMEOUT" gpurun_out/r2a_attn1.log | head -20
PROBE_ATTN=0 timeout 120 tests/cuda/tc_probe > gpurun_out/r2a_attn0.log 2>&1; echo "attn(mn=0) exit $?"; grep -E "ATTN|error|TIMEOUT" gpurun_out/r2a_attn0.log | head -12
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2a_tests.log 2>&1; tail -25 gpurun_out/r2a_tests.log | cut -c1-220
for knob in "BV2_DDS_FUSED=0" "BV2_TOK_GEMM=0" "BV2_DDS_FUSED=0 BV2_TOK_GEMM=0"; do echo "== $knob"; env $knob timeout 300 python -m pytest tests -m gpu -q -k "duration_stage or text_encoder" 2>&1 | tail -3 | cut -c1-200; done
for prec in tf32 fp16g fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2a_bench_${prec}_err.log | tail -1 > gpurun_out/r2a_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2a_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"])
except Exception as ex:
    print("${prec} bench failed", ex)
PY
done
tail -3 gpurun_out/r2a_bench_fp16_err.log
