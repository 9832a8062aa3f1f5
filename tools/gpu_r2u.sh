#!/bin/bash
# round-2 call U: incremental ring state in issuers/producers (no runtime division per stage), FFN conv_2 + LayerNorm fused: probe, tests, bench, launch list
mkdir -p gpurun_out
timeout 300 tests/cuda/g2_probe perf > gpurun_out/r2u_g2_probe.log 2>&1; echo "probe exit $?"; grep -c PASS gpurun_out/r2u_g2_probe.log; grep "FAIL\|PROBE\|error" gpurun_out/r2u_g2_probe.log | head
C="tests/cuda/g2_probe case"
G2_PROF=1 timeout 200 $C 256 256 11 1 8184 0 20 1  256 256 3 1 8184 0 20 1  128 128 11 1 32736 0 20 1  128 128 3 1 32736 0 20 1  64 64 11 1 130944 0 20 1  32 32 11 1 261888 0 20 1  16 16 11 1 523776 0 20 1  16 16 3 1 523776 0 20 1 > gpurun_out/r2u_prof.log 2>&1
grep "PASS\|FAIL\|cta    0" gpurun_out/r2u_prof.log | cut -c1-60,100-330
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_tests.log 2>&1; tail -4 gpurun_out/r2u_tests.log | cut -c1-300
for prec in fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2u_bench_${prec}_err.log | tail -1 > gpurun_out/r2u_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2u_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2u_bench_${prec}_err.log
done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2u_launches_fp16.csv python tools/profile_step.py --steps 3 --precision fp16 > gpurun_out/r2u_ncu.log 2>&1
tail -2 gpurun_out/r2u_ncu.log
