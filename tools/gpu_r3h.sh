#!/bin/bash
# last call of round 2: attention with chunk-level band predicate + vector Ek loads: probe (all split cases), flow/full-infer parity, bench, launch list
mkdir -p gpurun_out
PROBE_ATTN=1 PROBE_ATTN_TIMELINE=1 timeout 60 tests/cuda/tc_probe perf > gpurun_out/r3h_attn_timeline.log 2>&1; echo "attn probe exit $?"; grep -c PASS gpurun_out/r3h_attn_timeline.log; grep "FAIL\|PROBE" gpurun_out/r3h_attn_timeline.log; grep -B3 "T=1023 B=1.*key-split=4 " gpurun_out/r3h_attn_timeline.log | cut -c1-200
timeout 150 python -m pytest tests -m gpu -q -x -k "flow_stage or full_infer or config2_full or config3" > gpurun_out/r3h_tests.log 2>&1; tail -2 gpurun_out/r3h_tests.log | cut -c1-200
timeout 100 python bench.py --steps 20 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r3h_bench_err.log | tail -1 > gpurun_out/r3h_bench.json
python -c "
import json
d = json.load(open('gpurun_out/r3h_bench.json'))
print('bench value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), d['stage_ms'])"
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r3h_launches_raw.csv python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r3h_profile_step.log 2>&1
tail -1 gpurun_out/r3h_profile_step.log
