#!/bin/bash
# round-2 call F: what bounds the G2 MMA phase?  tap alignment (dil = 8 -> every tap 128-byte aligned), pure GEMM (K = 1), ncu of one launch
mkdir -p gpurun_out
C="tests/cuda/g2_probe case"
$C 128 128 7 1 65472 4 20 0   128 128 7 8 65472 4 20 0   128 128 7 2 65472 4 20 0   128 128 7 4 65472 4 20 0   128 128 1 1 65472 4 20 0 \
   128 128 3 8 65472 4 20 0   128 128 3 1 65472 4 20 0 \
   16 16 7 1 523776 0 20 0   16 16 7 8 523776 0 20 0   16 16 1 1 523776 0 20 0 \
   64 64 7 1 130944 7 20 0   64 64 7 8 130944 7 20 0 \
   256 256 7 1 8184 1 20 0   256 256 7 8 8184 1 20 0 2>&1 | tee gpurun_out/r2f_cases.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_g2_conv -s 2 -c 1 -f -o gpurun_out/r2f_g2_c128k7 $C 128 128 7 1 65472 4 3 0 > gpurun_out/r2f_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_g2_conv -s 2 -c 1 -f -o gpurun_out/r2f_g2_c16k11 $C 16 16 11 1 523776 0 3 0 > gpurun_out/r2f_ncu2.log 2>&1
ls -la gpurun_out | cat
