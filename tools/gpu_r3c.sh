#!/bin/bash
# round-2 (session 4) call C: fused attention with whole-tile TMEM loads + all DSMEM merge loads in flight; A/B of the register-resident LN tail and
# the conv-deep weight ring (tuning build knobs) on one box
mkdir -p gpurun_out
PROBE_ATTN=1 PROBE_ATTN_TIMELINE=1 timeout 200 tests/cuda/tc_probe perf > gpurun_out/r3c_attn_timeline.log 2>&1; echo "attn probe exit $?"; grep "FAIL\|PROBE\| ms" gpurun_out/r3c_attn_timeline.log | cut -c1-200; grep -B6 "T=1023 B=1.*key-split=4 " gpurun_out/r3c_attn_timeline.log | cut -c1-200
timeout 500 python -m pytest tests -m gpu -q -x -k "flow or full_infer or config2 or config3 or wn_flow" > gpurun_out/r3c_tests.log 2>&1; tail -4 gpurun_out/r3c_tests.log | cut -c1-300
run_bench() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 200 python bench.py --precision fp16 --steps 20 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r3c_bench_${tag}_err.log | tail -1 > gpurun_out/r3c_bench_${tag}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3c_bench_${tag}.json"))
    print("${tag}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"])
except Exception as ex:
    print("${tag} bench failed", ex)
PY
  tail -2 gpurun_out/r3c_bench_${tag}_err.log
}
T=$PWD/bert_vits2_b200/libbv2_tuning.so
run_bench product BV2_DUMMY=1
run_bench t_default BV2_LIB=$T
run_bench t_lnregs0 BV2_LIB=$T BV2_LN_REGS=0
run_bench t_nws8 BV2_LIB=$T BV2_TC_NWS_MAX=8
run_bench t_both_off BV2_LIB=$T BV2_LN_REGS=0 BV2_TC_NWS_MAX=8
run_bench product_again BV2_DUMMY=1
