#!/bin/bash
# round-2 final evidence run (one B200): full GPU test suite, smoke, default bench line (extras on), ncu launch list with DRAM bytes of one step,
# ncu --set full of the flow kernels and of k_g2_conv, timelines of the final build, halo-zeroing cost probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r3d_tests.log 2>&1; tail -3 gpurun_out/r3d_tests.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3d_smoke.log 2>&1; tail -1 gpurun_out/r3d_smoke.log
timeout 400 python bench.py 2> gpurun_out/r3d_bench_err.log | tail -1 > gpurun_out/r3d_bench.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r3d_bench.json"))
    print("bench value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches/step", d["launches_per_step"])
    for k in ("config2_length_scale_1", "config3_batched", "config5_generator", "flow_wn", "cpu_baseline"):
        v = d.get(k, {})
        print("  ", k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("value", "error", "stage_ms", "frac", "cores", "ms_per_step")})
except Exception as ex:
    print("bench failed", ex)
PY
tail -2 gpurun_out/r3d_bench_err.log
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r3d_launches_raw.csv python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r3d_profile_step.log 2>&1
tail -1 gpurun_out/r3d_profile_step.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv1d|k_flow_attn" --launch-skip 130 --launch-count 7 -f \
    -o gpurun_out/r3d_flow_kernels python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r3d_ncu_flow.log 2>&1; tail -1 gpurun_out/r3d_ncu_flow.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_g2_conv --launch-skip 122 --launch-count 4 -f \
    -o gpurun_out/r3d_g2_conv python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r3d_ncu_g2.log 2>&1; tail -1 gpurun_out/r3d_ncu_g2.log
PROBE_FLOW=1 timeout 100 tests/cuda/tc_probe > gpurun_out/r3d_flow_timeline.log 2>&1; echo "flow probe exit $?"
PROBE_ATTN=1 PROBE_ATTN_TIMELINE=1 timeout 100 tests/cuda/tc_probe perf > gpurun_out/r3d_attn_timeline.log 2>&1; echo "attn probe exit $?"; grep -B6 "T=1023 B=1.*key-split=4 " gpurun_out/r3d_attn_timeline.log | cut -c1-200
CASES="256 256 3 1 8184 0 20 1  128 128 7 1 65472 0 20 1  64 64 11 1 130944 0 20 1  32 32 3 1 261888 0 20 1  16 16 7 1 523776 0 20 1"
timeout 100 tests/cuda/g2_probe case $CASES > gpurun_out/r3d_g2_halo_on.log 2>&1
G2_DBG=16 timeout 100 tests/cuda/g2_probe case $CASES > gpurun_out/r3d_g2_halo_off.log 2>&1
paste -d'\n' gpurun_out/r3d_g2_halo_on.log gpurun_out/r3d_g2_halo_off.log | cut -c1-170
ls -la gpurun_out | head -40
