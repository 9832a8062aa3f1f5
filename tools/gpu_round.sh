#!/bin/bash
# One GPU-box visit that produces everything a round needs (run it THROUGH gpurun; ~4 min of box time on one B200):
#   gpurun --timeout 420 -- 'bash tools/gpu_round.sh r02 k_tc_pair_persist 20'
# outputs (merged back into gpurun_out/): <tag>_tests.log, <tag>_smoke.log, <tag>_bench.json, <tag>_launches.csv,
# <tag>_<kernel>.ncu-rep.  Summarise here with tools/ncu_summary.py and copy what should be judged into profiles/.
# Every step has its own timeout: a hung kernel must never hold the box (bounded mbarrier waits trap, but be safe).
tag=${1:-rNN}; kernel=${2:-k_tc_conv1d_persist}; skip=${3:-20}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 150 python bench.py 2> gpurun_out/${tag}_bench_err.log | tail -1 > gpurun_out/${tag}_bench.json
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("bench", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["stage_ms"])
PY
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py --steps 2 > /dev/null 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:${kernel} -s ${skip} -c 1 -f \
    -o gpurun_out/${tag}_${kernel} python tools/profile_step.py --steps 2 > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_* | cat
