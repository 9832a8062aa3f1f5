#!/bin/bash
# round-2 call V: ncu --set full of one flow layer's kernels (qkv, fused attention, conv_o+LN, FFN conv_1, FFN conv_2+LN) in the real step, and of k_g2_conv C=128 K=11 (new build)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tc_conv1d|k_flow_attn" -s 177 -c 5 -f -o gpurun_out/r2v_flow_layer python tools/profile_step.py --steps 3 --precision fp16 > gpurun_out/r2v_ncu_flow.log 2>&1
tail -3 gpurun_out/r2v_ncu_flow.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_g2_conv -s 2 -c 1 -f -o gpurun_out/r2v_g2_c128k11 tests/cuda/g2_probe case 128 128 11 1 32736 0 3 1 > gpurun_out/r2v_ncu_g2.log 2>&1
tail -2 gpurun_out/r2v_ncu_g2.log
ls -la gpurun_out/r2v*
