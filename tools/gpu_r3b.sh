#!/bin/bash
# round-2 (session 4) call B: register-resident LayerNorm tail (k_tc_conv1d<1,1,1>), weight ring as deep as the conv on small grids, conv_post staged
# through shared memory with constant-bank weights; attention phase timeline; parity subset; bench; ncu launch list of one step
mkdir -p gpurun_out
PROBE_FLOW=1 timeout 200 tests/cuda/tc_probe > gpurun_out/r3b_flow_timeline.log 2>&1; echo "flow probe exit $?"; grep -A4 "^FLOW" gpurun_out/r3b_flow_timeline.log | grep -A4 "conv_o+LN  \|conv_2 f16i  \|conv_2+LN f16i\|^FLOW post" | cut -c1-200
PROBE_ATTN=1 PROBE_ATTN_TIMELINE=1 timeout 200 tests/cuda/tc_probe perf > gpurun_out/r3b_attn_timeline.log 2>&1; echo "attn probe exit $?"; grep -A6 "T=1023 B=1.*key-split=[04]\|T=1024 B=1" gpurun_out/r3b_attn_timeline.log | cut -c1-220; tail -1 gpurun_out/r3b_attn_timeline.log
timeout 500 python -m pytest tests -m gpu -q -x -k "flow or full_infer or generator or config5 or pcm16 or config2 or config3 or packed" > gpurun_out/r3b_tests.log 2>&1; tail -4 gpurun_out/r3b_tests.log | cut -c1-300
timeout 200 python bench.py --precision fp16 --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r3b_bench_err.log | tail -1 > gpurun_out/r3b_bench.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3b_bench.json"))
    print("bench value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("bench failed", ex)
PY
tail -2 gpurun_out/r3b_bench_err.log
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r3b_launches_raw.csv python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r3b_profile_step.log 2>&1
tail -2 gpurun_out/r3b_profile_step.log
python tools/ncu_summary.py launches gpurun_out/r3b_launches_raw.csv --skip-first 278 2>/dev/null | head -12
