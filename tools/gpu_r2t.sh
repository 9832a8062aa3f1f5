#!/bin/bash
# round-2 call T: why is the in-situ MMA rate 3x the microbenchmark's?  contention knobs (4: no TMA traffic after the ring fill, 8: parked epilogue warps)
# + ncu --set full of k_g2_conv on three Generator shapes
mkdir -p gpurun_out
C="tests/cuda/g2_probe case"
CASES="128 128 11 1 32736 0 20 1  64 64 11 1 130944 0 20 1  16 16 11 1 523776 0 20 1"
for dbg in 4 8 12; do
  echo "==== G2_DBG=$dbg" >> gpurun_out/r2t_prof.log
  G2_DBG=$dbg G2_PROF=1 timeout 100 $C $CASES >> gpurun_out/r2t_prof.log 2>&1
done
grep "====\|PASS\|FAIL\|cta    0" gpurun_out/r2t_prof.log | cut -c1-60,100-330
for shape in "128 128 11 1 32736" "64 64 11 1 130944" "16 16 11 1 523776"; do
  tag=$(echo $shape | awk '{print "c"$1"k"$3}')
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_g2_conv -s 2 -c 1 -f -o gpurun_out/r2t_g2_$tag $C $shape 0 3 1 > gpurun_out/r2t_ncu_$tag.log 2>&1
  tail -2 gpurun_out/r2t_ncu_$tag.log
done
ls -la gpurun_out/*.ncu-rep
