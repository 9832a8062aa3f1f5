#!/bin/bash
# round-2 call D: ncu --set full over every launch of one fp16 step (no source import: report size), read here with ncu_summary.py
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --launch-skip 263 --launch-count 263 -f -o gpurun_out/r2d_step_fp16 python tools/profile_step.py --steps 2 --precision fp16 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/ | cat
