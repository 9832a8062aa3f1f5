#!/bin/bash
# round-2 call H: k_g2_conv with the warp-uniform MMA issuer: correctness + per-shape timings + timeline
mkdir -p gpurun_out
timeout 300 tests/cuda/g2_probe perf > gpurun_out/r2h_g2_probe.log 2>&1; echo "probe exit $?"; grep -v "^PASS.*T=   \|^PASS.*T=  [0-9][0-9][0-9][0-9] B=[23]" gpurun_out/r2h_g2_probe.log | cut -c1-60,125-250
C="tests/cuda/g2_probe case"
G2_PROF=1 $C 128 128 7 1 65472 4 20 0  16 16 7 1 523776 0 20 0  64 64 7 1 130944 7 20 1 2>&1 | tee gpurun_out/r2h_prof.log | cut -c1-230
