#!/bin/bash
# round-2 (session 4) call A: LayerNorm tails with the TMA-staged residual tile, halo zeroing inside the producers, 4-output conv_post:
# flow timelines, a parity subset, bench; A/B of FFN conv_2 on one full-width N tile with the LayerNorm tail (tuning build, BV2_F2_NT=192)
mkdir -p gpurun_out
PROBE_FLOW=1 timeout 200 tests/cuda/tc_probe > gpurun_out/r3a_flow_timeline.log 2>&1; echo "flow probe exit $?"; grep -A4 "^FLOW" gpurun_out/r3a_flow_timeline.log | cut -c1-200
timeout 500 python -m pytest tests -m gpu -q -x -k "flow or full_infer or generator or config5 or pcm16 or config2 or config3" > gpurun_out/r3a_tests.log 2>&1; tail -4 gpurun_out/r3a_tests.log | cut -c1-300
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --precision fp16 --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r3a_bench_${tag}_err.log | tail -1 > gpurun_out/r3a_bench_${tag}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3a_bench_${tag}.json"))
    print("${tag}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${tag} bench failed", ex)
PY
  tail -2 gpurun_out/r3a_bench_${tag}_err.log
}
run_bench product BV2_DUMMY=1
run_bench tuning_f2nt192 BV2_LIB=$PWD/bert_vits2_b200/libbv2_tuning.so BV2_F2_NT=192
run_bench tuning_res_acc BV2_LIB=$PWD/bert_vits2_b200/libbv2_tuning.so BV2_LN_RES_SMEM=0
BV2_LIB=$PWD/bert_vits2_b200/libbv2_tuning.so BV2_F2_NT=192 timeout 300 python -m pytest tests -m gpu -q -x -k "flow_stage_tensor_core or full_infer" > gpurun_out/r3a_tests_f2nt192.log 2>&1; tail -3 gpurun_out/r3a_tests_f2nt192.log | cut -c1-300
