#!/bin/bash
# round-2 call E: G2 (16-bit activation Generator) probe + parity tests + bench
mkdir -p gpurun_out
timeout 300 tests/cuda/g2_probe perf > gpurun_out/r2e_g2_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/r2e_g2_probe.log | cut -c1-250
timeout 600 python -m pytest tests -m gpu -q -x -k "generator or config5 or config2 or full_infer or smoke" > gpurun_out/r2e_tests.log 2>&1; tail -15 gpurun_out/r2e_tests.log | cut -c1-300
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -2 gpurun_out/r2e_smoke.log
for prec in fp16; do
  timeout 200 python bench.py --precision $prec --steps 10 --cpu-baseline-steps 0 --extras 0 2> gpurun_out/r2e_bench_${prec}_err.log | tail -1 > gpurun_out/r2e_bench_${prec}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2e_bench_${prec}.json"))
    print("${prec}", "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"], "launches", d["gpu_launches"], "frames", d["config"].get("frames_per_utterance"))
except Exception as ex:
    print("${prec} bench failed", ex)
PY
  tail -2 gpurun_out/r2e_bench_${prec}_err.log
done
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/r2e_launches_fp16.csv python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r2e_ncu_fp16.log 2>&1
ls -la gpurun_out | cat
