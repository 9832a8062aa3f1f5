#!/bin/bash
# round-2 call C (re-entry): state of the committed build -- full GPU tests, smoke, A/B bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; tail -30 gpurun_out/r2c_tests.log | cut -c1-300
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -2 gpurun_out/r2c_smoke.log
for prec in tf32 fp16g fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2c_bench_${prec}_err.log | tail -1 > gpurun_out/r2c_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2c_bench_${prec}_err.log
done
for prec in tf32 fp16; do
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/r2c_launches_${prec}.csv python tools/profile_step.py --steps 2 --precision $prec > gpurun_out/r2c_ncu_${prec}.log 2>&1
done
ls -la gpurun_out | cat
