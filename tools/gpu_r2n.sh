#!/bin/bash
# round-2 call N: new epilogue correctness + MMA-phase cycle accounting, with / without weight-stage commits
mkdir -p gpurun_out
timeout 300 tests/cuda/g2_probe > gpurun_out/r2n_g2_probe.log 2>&1; echo "probe exit $?"; grep -c PASS gpurun_out/r2n_g2_probe.log; grep "FAIL\|PROBE\|error" gpurun_out/r2n_g2_probe.log | head
C="tests/cuda/g2_probe case"
G2_PROF=1 $C 128 128 3 1 65472 4 20 0  128 128 7 1 65472 4 20 0  16 16 7 1 523776 0 20 0  64 64 7 1 130944 7 20 1 2>&1 | tee gpurun_out/r2n_prof.log | cut -c1-300
echo "---- without weight-stage commits (K = 3: 12 stages fit the ring)"
G2_SKIP_WCOMMIT=1 G2_PROF=1 $C 128 128 3 1 65472 4 20 0 2>&1 | tee -a gpurun_out/r2n_prof.log | cut -c1-300
