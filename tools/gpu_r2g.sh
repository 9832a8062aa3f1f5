#!/bin/bash
# round-2 call G: per-CTA phase timeline of k_g2_conv (globaltimer stamps)
mkdir -p gpurun_out
C="tests/cuda/g2_probe case"
G2_PROF=1 $C 128 128 7 1 65472 4 20 0   128 128 1 1 65472 4 20 0   128 128 7 1 65472 4 20 1   16 16 7 1 523776 0 20 0  16 16 7 1 523776 0 20 1  64 64 7 1 130944 7 20 1 2>&1 | tee gpurun_out/r2g_prof.log | cut -c1-230
