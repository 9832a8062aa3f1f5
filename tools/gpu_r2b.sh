#!/bin/bash
# round-2 call B: timings after the barrier-poll fix, locate the launch failure, full GPU tests, A/B bench
mkdir -p gpurun_out
timeout 300 tests/cuda/tc_probe perf > gpurun_out/r2b_probe.log 2>&1; echo "probe exit $?"; grep -c PASS gpurun_out/r2b_probe.log; grep -E "FAIL|error|TIMEOUT|PROBE" gpurun_out/r2b_probe.log | head -20
PROBE_ATTN=1 timeout 120 tests/cuda/tc_probe perf > gpurun_out/r2b_attn.log 2>&1; grep -E "ATTN|error|TIMEOUT" gpurun_out/r2b_attn.log | head -12
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests -m gpu -x -q -k "peer_slab" > gpurun_out/r2b_t_peer.log 2>&1; tail -15 gpurun_out/r2b_t_peer.log | cut -c1-300
if grep -q "failed\|Aborted\|error" gpurun_out/r2b_t_peer.log; then
  timeout 400 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests -m gpu -x -q -k "peer_slab" > gpurun_out/r2b_sanitizer.log 2>&1; grep -E "Invalid|at |by thread|Address|=========     in" gpurun_out/r2b_sanitizer.log | head -40
fi
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_peer_slab_output_pointer_world1 > gpurun_out/r2b_tests.log 2>&1; tail -30 gpurun_out/r2b_tests.log | cut -c1-300
for prec in tf32 fp16g fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2b_bench_${prec}_err.log | tail -1 > gpurun_out/r2b_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"]["frames_per_utterance"])
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2b_bench_${prec}_err.log
done
