// Fused FP32 kernels of the token-rate stage (text encoder, SDP, DP: everything that feeds ceil(durations) stays on FP32 FMA,
// SURVEY.md section 7 H1).  At T ~ 256 tokens this stage is pure launch / dependency latency (round 1: ~110 launches of 5-40 us),
// so the kernels here trade launches for on-chip fusion: a whole layer per launch, weights staged in shared memory by TMA.
#pragma once
#include "kernels_simt.cuh"
#include "tc_conv.cuh"

namespace bv2 {

// ------------------------------------------------------------------------------------------------
// One DDSConv layer in one launch (reference modules.py:118-130, channels C = 192):
//     y = convs_sep[i](x * x_mask)   depthwise k=3, dilation d      y = gelu(norms_1[i](y))
//     y = convs_1x1[i](y)            dense C x C                    y = gelu(norms_2[i](y))
//     x = x + y                      (* x_mask after the last layer, modules.py:130)
// One CTA = 16 time steps x all channels, 8 warps.  The 1x1 weight matrix (C*C*4 = 144 KB, packed [ci][co]) is staged in
// shared memory with ONE bulk TMA copy that overlaps the depthwise conv + first LayerNorm; both LayerNorms are warp-shuffle
// reductions (one warp owns a time step: phase 1 by channel group, phase 3 by the 6 output channels a lane computed).
// x is read with a halo of +-d from neighbouring tiles, so the layer writes to a different buffer than it reads.
// ------------------------------------------------------------------------------------------------
struct DdsArgs {
    const float* x; float* y;            // c4 [B][C/4][T][4], in / out (different buffers)
    const float* dw_w; const float* dw_b;  // depthwise [C][3], [C]
    const float* w1; const float* b1;      // 1x1 packed [C][C] (ci major, co fastest), bias [C]
    const float* g1; const float* be1; const float* g2; const float* be2;  // LayerNorm gamma / beta
    const int* lens;
    int T, B, dil, last;                 // last: multiply the result by x_mask
};

template <int C>
__global__ void __launch_bounds__(256, 1) k_dds_layer(DdsArgs a) {
    using namespace tc;
    static_assert(C == 192, "lane -> channel mapping below is written for 192 channels (48 c4 groups)");
    constexpr int TT = 16, NCG = C / 4;
    extern __shared__ __align__(128) uint8_t smem[];
    float* sW = reinterpret_cast<float*>(smem);                      // [C][C]
    float* sY = sW + C * C;                                          // [TT][C]: activations between LN1/GELU and the 1x1 conv
    uint64_t* bar = reinterpret_cast<uint64_t*>(sY + C * TT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * TT, b = blockIdx.y;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(smem_u32(bar), (uint32_t)(C * C * 4));
        bulk_g2s(smem_u32(sW), a.w1, (uint32_t)(C * C * 4), smem_u32(bar));  // weights: independent of the upstream kernel
    }
    pdl_wait();
    const int len = a.lens ? a.lens[b] : a.T;
    const float4* x4 = reinterpret_cast<const float4*>(a.x) + (size_t)b * NCG * a.T;
    // ---- phase 1: depthwise conv + LayerNorm 1 + GELU; warp w owns time steps 2w, 2w+1; lane owns c4 groups lane and 32+lane (<16)
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int tt = 2 * warp + k, t = t0 + tt;
        float4 v[2];
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int cg = lane + 32 * q;
            v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cg < NCG && t < a.T) {
                float4 acc = *reinterpret_cast<const float4*>(a.dw_b + cg * 4);
                const float* wc = a.dw_w + cg * 12;
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const int u = t + (j - 1) * a.dil;
                    if (u >= 0 && u < a.T && u < len) {
                        const float4 xv = x4[(size_t)cg * a.T + u];
                        acc.x = fmaf(xv.x, wc[j], acc.x); acc.y = fmaf(xv.y, wc[3 + j], acc.y);
                        acc.z = fmaf(xv.z, wc[6 + j], acc.z); acc.w = fmaf(xv.w, wc[9 + j], acc.w);
                    }
                }
                v[q] = acc;
                s += (acc.x + acc.y) + (acc.z + acc.w);
            }
        }
        s = warp_sum(s);
        const float mean = s / (float)C;
        float qv = 0.f;
#pragma unroll
        for (int q = 0; q < 2; q++)
            if (lane + 32 * q < NCG) {
                const float dx = v[q].x - mean, dy = v[q].y - mean, dz = v[q].z - mean, dw = v[q].w - mean;
                qv += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        qv = warp_sum(qv);
        const float rstd = rsqrtf(qv / (float)C + 1e-5f);
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int cg = lane + 32 * q;
            if (cg < NCG) {
                const float4 g = *reinterpret_cast<const float4*>(a.g1 + cg * 4), be = *reinterpret_cast<const float4*>(a.be1 + cg * 4);
                float4 o;
                o.x = gelu_erf((v[q].x - mean) * rstd * g.x + be.x); o.y = gelu_erf((v[q].y - mean) * rstd * g.y + be.y);
                o.z = gelu_erf((v[q].z - mean) * rstd * g.z + be.z); o.w = gelu_erf((v[q].w - mean) * rstd * g.w + be.w);
                *reinterpret_cast<float4*>(&sY[tt * C + cg * 4]) = o;  // consecutive lanes -> consecutive 16 bytes
            }
        }
    }
    __syncthreads();
    mbar_wait(smem_u32(bar), 0);
    // ---- phase 2: 1x1 conv.  warp w: time steps 2w, 2w+1; lane: output channels 4*lane..4*lane+3 and 128+2*lane, 129+2*lane
    float acc[2][6];
    {
        const float4 bA = *reinterpret_cast<const float4*>(a.b1 + 4 * lane);
        const float2 bB = *reinterpret_cast<const float2*>(a.b1 + 128 + 2 * lane);
#pragma unroll
        for (int k = 0; k < 2; k++) { acc[k][0] = bA.x; acc[k][1] = bA.y; acc[k][2] = bA.z; acc[k][3] = bA.w; acc[k][4] = bB.x; acc[k][5] = bB.y; }
    }
    for (int c0 = 0; c0 < C; c0 += 4) {
        // 4 input channels per step: the two time steps' activations are warp-uniform (broadcast) 16-byte loads
        const float4 ya = *reinterpret_cast<const float4*>(&sY[(2 * warp) * C + c0]), yb = *reinterpret_cast<const float4*>(&sY[(2 * warp + 1) * C + c0]);
        const float y0[4] = {ya.x, ya.y, ya.z, ya.w}, y1[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {  // ascending ci: the summation order is fixed
            const float4 wA = *reinterpret_cast<const float4*>(&sW[(c0 + e) * C + 4 * lane]);
            const float2 wB = *reinterpret_cast<const float2*>(&sW[(c0 + e) * C + 128 + 2 * lane]);
            acc[0][0] = fmaf(y0[e], wA.x, acc[0][0]); acc[0][1] = fmaf(y0[e], wA.y, acc[0][1]); acc[0][2] = fmaf(y0[e], wA.z, acc[0][2]);
            acc[0][3] = fmaf(y0[e], wA.w, acc[0][3]); acc[0][4] = fmaf(y0[e], wB.x, acc[0][4]); acc[0][5] = fmaf(y0[e], wB.y, acc[0][5]);
            acc[1][0] = fmaf(y1[e], wA.x, acc[1][0]); acc[1][1] = fmaf(y1[e], wA.y, acc[1][1]); acc[1][2] = fmaf(y1[e], wA.z, acc[1][2]);
            acc[1][3] = fmaf(y1[e], wA.w, acc[1][3]); acc[1][4] = fmaf(y1[e], wB.x, acc[1][4]); acc[1][5] = fmaf(y1[e], wB.y, acc[1][5]);
        }
    }
    // ---- phase 3: LayerNorm 2 + GELU + residual (+ mask), straight from the accumulators
    const float4 gA = *reinterpret_cast<const float4*>(a.g2 + 4 * lane), eA = *reinterpret_cast<const float4*>(a.be2 + 4 * lane);
    const float2 gB = *reinterpret_cast<const float2*>(a.g2 + 128 + 2 * lane), eB = *reinterpret_cast<const float2*>(a.be2 + 128 + 2 * lane);
    float4* y4 = reinterpret_cast<float4*>(a.y) + (size_t)b * NCG * a.T;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int t = t0 + 2 * warp + k;
        float s = ((acc[k][0] + acc[k][1]) + (acc[k][2] + acc[k][3])) + (acc[k][4] + acc[k][5]);
        s = warp_sum(s);
        const float mean = s / (float)C;
        float qv = 0.f;
#pragma unroll
        for (int e = 0; e < 6; e++) { const float d = acc[k][e] - mean; qv += d * d; }
        qv = warp_sum(qv);
        const float rstd = rsqrtf(qv / (float)C + 1e-5f);
        if (t >= a.T) continue;
        const float m = (a.last && t >= len) ? 0.f : 1.f;
        const float4 rA = x4[(size_t)lane * a.T + t];
        float4 oA;
        oA.x = (rA.x + gelu_erf((acc[k][0] - mean) * rstd * gA.x + eA.x)) * m;
        oA.y = (rA.y + gelu_erf((acc[k][1] - mean) * rstd * gA.y + eA.y)) * m;
        oA.z = (rA.z + gelu_erf((acc[k][2] - mean) * rstd * gA.z + eA.z)) * m;
        oA.w = (rA.w + gelu_erf((acc[k][3] - mean) * rstd * gA.w + eA.w)) * m;
        y4[(size_t)lane * a.T + t] = oA;
        // channels 128 + 2*lane, +1 = half of c4 group 32 + lane/2
        const float2* xr = reinterpret_cast<const float2*>(x4 + (size_t)(32 + (lane >> 1)) * a.T + t) + (lane & 1);
        const float2 rB = *xr;
        float2 oB;
        oB.x = (rB.x + gelu_erf((acc[k][4] - mean) * rstd * gB.x + eB.x)) * m;
        oB.y = (rB.y + gelu_erf((acc[k][5] - mean) * rstd * gB.y + eB.y)) * m;
        *(reinterpret_cast<float2*>(y4 + (size_t)(32 + (lane >> 1)) * a.T + t) + (lane & 1)) = oB;
    }
}

inline size_t dds_layer_smem(int C) { return (size_t)C * C * 4 + (size_t)C * 16 * 4 + 16; }
inline void tok_init_device() {
    BV2_CUDA(cudaFuncSetAttribute(k_dds_layer<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dds_layer_smem(192)));
}
inline void launch_dds_layer(const DdsArgs& a, int C, cudaStream_t st) {
    BV2_CHECK(C == 192, "fused DDS layer is instantiated for 192 channels");
    launch_pdl(k_dds_layer<192>, dim3(cdiv(a.T, 16), a.B), dim3(256), dds_layer_smem(C), st, a);
}

}  // namespace bv2
