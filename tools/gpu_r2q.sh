#!/bin/bash
# round-2 call Q: MMA issue-rate microbenchmark + per-CTA phase accounting of k_g2_conv on the five Generator stage shapes
mkdir -p gpurun_out
timeout 120 tests/cuda/mma_rate > gpurun_out/r2q_mma_rate.log 2>&1; echo "mma_rate exit $?"; cat gpurun_out/r2q_mma_rate.log | cut -c1-200
timeout 300 tests/cuda/g2_probe perf > gpurun_out/r2q_g2_probe.log 2>&1; echo "probe exit $?"; grep "ms \|PROBE" gpurun_out/r2q_g2_probe.log | cut -c1-60,125-250
C="tests/cuda/g2_probe case"
G2_PROF=1 timeout 200 $C 256 256 11 1 8184 0 20 1  256 256 3 1 8184 0 20 1  128 128 11 1 32736 0 20 1  128 128 3 1 32736 0 20 1  64 64 11 1 130944 0 20 1  32 32 11 1 261888 0 20 1  16 16 11 1 523776 0 20 1  16 16 3 1 523776 0 20 1 > gpurun_out/r2q_prof.log 2>&1
grep -A3 "^PASS\|^FAIL" gpurun_out/r2q_prof.log | cut -c1-330
