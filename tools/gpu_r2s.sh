#!/bin/bash
# round-2 call S: what bounds the MMA issue rate on the uniform path: descriptor study microbenchmark + in-situ knobs of k_g2_conv
mkdir -p gpurun_out
MMA_RATE_DESC=1 timeout 120 tests/cuda/mma_rate > gpurun_out/r2s_desc.log 2>&1; echo "mma_rate exit $?"; cat gpurun_out/r2s_desc.log | cut -c1-160
C="tests/cuda/g2_probe case"
CASES="128 128 11 1 32736 0 20 1  64 64 11 1 130944 0 20 1  16 16 11 1 523776 0 20 1"
for dbg in 0 1 2 3; do
  echo "==== G2_DBG=$dbg" >> gpurun_out/r2s_prof.log
  G2_DBG=$dbg G2_PROF=1 timeout 100 $C $CASES >> gpurun_out/r2s_prof.log 2>&1
done
echo "==== G2_SKIP_WCOMMIT K=3" >> gpurun_out/r2s_prof.log
G2_SKIP_WCOMMIT=1 G2_PROF=1 timeout 100 $C 128 128 3 1 32736 0 20 1 >> gpurun_out/r2s_prof.log 2>&1
grep "====\|PASS\|FAIL\|cta    0" gpurun_out/r2s_prof.log | cut -c1-60,100-330
