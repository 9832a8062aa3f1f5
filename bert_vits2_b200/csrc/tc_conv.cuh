// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d on the c4 activation layout, TF32 operands, fp32 accumulate
// in TMEM.  Replaces the 90 dilated MRF convolutions of the HiFi-GAN Generator (reference modules.py:296-309 via
// models.py:546-552) = 96 % of the MACs of SynthesizerTrn.infer.
//
// GEMM view of one CTA tile:  D[128 time steps, N = Cout] += sum_{tap j} sum_{ci}  A_j[t, ci] * W_j[ci, co]
//   A_j[t, ci] = act(x[ci][t0 + t + j*dil - pad])  -- a time-shifted view of ONE staged activation tile.
// With the c4 layout a tile chunk is staged as [KC/4 channel groups][R = 128 + (K-1)*dil rows][4 ch] fp32, i.e. the
// K-major / no-swizzle UMMA canonical layout with SBO = 128 B (8 rows x 16 B) and LBO = R*16 B, so tap j is just the
// smem-descriptor start address advanced by j*dil*16 bytes: the im2col matrix is never materialised and each
// activation byte is fetched from HBM/L2 once per conv instead of K times.
//
// Warp roles (192 threads): warp 0 = TMA producer (cp.async.bulk, mbarrier complete_tx), warp 1 = TMEM allocator +
// single-thread tcgen05.mma issuer, warps 2-5 = operand prologue (leaky-relu + round-to-nearest TF32 in smem, zero
// fill of the conv padding rows, fence.proxy.async) and TMEM epilogue (tcgen05.ld -> +bias, +residual, MRF
// accumulate/scale -> coalesced 16-byte stores).
#pragma once
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include "common.cuh"

namespace bv2 {

struct TcConvW {
    float* w = nullptr;  // packed [Cout/nt N tiles][nchunks][K][KC/4][nt][4], TF32-rounded (RN): one contiguous smem image per stage
    int Cin = 0, Cout = 0, K = 0, KC = 0, nchunks = 0, nt = 0;
    int ups_u = 0, ups_cout = 0;  // polyphase ConvTranspose1d: Cout = ups_u * ups_cout columns (phase-major)
    int x3 = 0;  // error-compensated 3xTF32 (hi/lo operand split, fp32-class accuracy) for the stages that feed ceil(durations)
};
struct TcEpi {
    float in_slope = 1.f;        // leaky-relu slope applied to the conv INPUT (1 = identity)
    int in_mask = 0;             // input rows t >= lens[b] read as zero
    int relu = 0;
    int res_mode = 0;            // 1: v += res ; 2: v = res - v
    const float* res = nullptr;  // residual, c4, res_C_total channels, first channel res_c_off
    int res_C_total = 0, res_c_off = 0;
    int accumulate = 0;          // v += y_old
    float out_scale = 1.f;
    int out_mask = 0;            // v *= (t < lens[b])
    const int* lens = nullptr;
    const float* bias_b = nullptr;  // per-batch bias row (speaker conditioning)
    int bias_b_stride = 0;
    int cin_off = 0, cout_off = 0;  // channel windows inside x / y (multiples of 4)
    int dil = 1;
    int out_tf32 = 0;    // round the stored output to TF32 (RN): the consumer may then skip its operand prologue
    int skip_xform = 0;  // input already TF32-exact, no activation / mask / padding needed (K == 1): prologue warps only forward the barrier
};

inline float tf32_rn_host(float x) {
    uint32_t u; std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & 0xffffe000u;  // round to nearest, ties away (== cvt.rna.tf32.f32)
    float r; std::memcpy(&r, &u, 4);
    return r;
}

// w: [Cout][Cin][K] fp32 (weight-norm already folded)
// nt = N tile (0: largest divisor of Cout that is a multiple of 16 and <= 256)
// kc = K chunk (channels per pipeline stage): 16 for convs launched with thousands of tiles (small stages -> more
// resident CTAs per SM hide the per-tile latency chain), 32 for few-tile launches (shorter chunk loop per CTA).
inline TcConvW tc_pack_weights(std::function<float*(const std::vector<float>&)>& up, const std::vector<float>& w, int Cout, int Cin, int K, int nt = 0,
                               int x3 = 0, int kc = 0) {
    TcConvW t; t.Cin = Cin; t.Cout = Cout; t.K = K; t.x3 = x3;
    if (!nt) { nt = std::min(Cout, 256); while (Cout % nt || nt % 16) nt -= 16; }
    static const int kc_env = getenv("BV2_TC_KC") ? atoi(getenv("BV2_TC_KC")) : 0;  // tuning knob (experiments)
    if (kc_env) kc = kc_env;
    if (!kc) kc = 32;
    t.KC = Cin >= kc ? kc : Cin;
    if (Cin % t.KC != 0 || t.KC % 8 != 0 || Cout % 16 != 0 || Cout < 16)
        throw Error(-2, "tc_conv: unsupported channel counts " + std::to_string(Cin) + "->" + std::to_string(Cout));
    if (nt < 16 || nt > 256 || nt % 16 || Cout % nt) throw Error(-2, "tc_conv: bad N tile");
    t.nt = nt;
    t.nchunks = Cin / t.KC;
    const int parts = x3 ? 2 : 1;  // x3: every stage image is [hi | lo]
    std::vector<float> p((size_t)Cin * K * Cout * parts);
    const int ncg = t.KC / 4;
    for (int tile = 0; tile < Cout / nt; tile++)
        for (int c = 0; c < t.nchunks; c++)
            for (int j = 0; j < K; j++)
                for (int g = 0; g < ncg; g++)
                    for (int n = 0; n < nt; n++)
                        for (int e = 0; e < 4; e++) {
                            int ci = c * t.KC + g * 4 + e;
                            const float v = w[((size_t)(tile * nt + n) * Cin + ci) * K + j];
                            const float hi = tf32_rn_host(v);
                            const size_t stage = (((size_t)tile * t.nchunks + c) * K + j) * parts;
                            p[((stage * ncg + g) * nt + n) * 4 + e] = hi;
                            if (x3) p[(((stage + 1) * ncg + g) * nt + n) * 4 + e] = tf32_rn_host(v - hi);
                        }
    t.w = up(p);
    return t;
}

// ConvTranspose1d(Cin->Cout, K, stride u, padding (K-u)/2) as a polyphase stride-1 conv (reference models.py:543-545):
//   out[t*u + r][co] = sum_m sum_ci x[t + floor((r+p)/u) - m][ci] * w[ci][co][(r+p)%u + m*u]
// -> an ordinary conv over input-rate time with Kp taps (union of the per-phase offsets), N = u*Cout columns ordered
// (r, co), structural zeros where a phase does not use a tap.  wT: [Cin][Cout][K] (weight-norm folded).
inline TcConvW tc_pack_upsample(std::function<float*(const std::vector<float>&)>& up, const std::vector<float>& wT, int Cin, int Cout, int K, int u,
                                int kc = 0) {
    const int p = (K - u) / 2, taps = K / u;
    int omin = 1 << 30, omax = -(1 << 30);
    for (int r = 0; r < u; r++)
        for (int m = 0; m < taps; m++) { int o = (r + p) / u - m; omin = std::min(omin, o); omax = std::max(omax, o); }
    int half = std::max(-omin, omax);
    const int Kp = 2 * half + 1;  // symmetric so that pad = (Kp-1)/2
    std::vector<float> w((size_t)u * Cout * Cin * Kp, 0.f);  // [N = u*Cout][Cin][Kp]
    for (int r = 0; r < u; r++)
        for (int m = 0; m < taps; m++) {
            const int o = (r + p) / u - m, j = (r + p) % u + m * u, tap = o + half;
            for (int co = 0; co < Cout; co++)
                for (int ci = 0; ci < Cin; ci++)
                    w[(((size_t)(r * Cout + co)) * Cin + ci) * Kp + tap] = wT[((size_t)ci * Cout + co) * K + j];
        }
    int nt = std::min(u * Cout, 256);
    TcConvW t = tc_pack_weights(up, w, u * Cout, Cin, Kp, nt, 0, kc);
    t.ups_u = u; t.ups_cout = Cout;
    return t;
}

struct TcParams {
    const float* x; float* y; const float* w; const float* bias; const float* res; const float* bias_b; const int* lens;
    int Cin_total, cin_off, Cout_total, cout_off, res_C_total, res_c_off, bias_b_stride;
    int nt;           // columns per N tile
    int T, B, K, dil, pad, KC, nchunks, R, nws, nas, MT, x3;
    uint32_t a_stage_bytes, w_stage_bytes, tmem_cols, idesc;
    float in_slope, out_scale;
    int accumulate, relu, res_mode, in_mask, out_mask, ups_u, ups_cout;
    int out_tf32, skip_xform;
    // batched-GEMM extensions (attention): grid z = b * zsplit + h
    int zsplit;                 // 0/1: z == batch
    int x_batch_z, y_batch_z;   // 1: tensor's batch index is z (else b)
    int x_c_zstride, y_c_zstride;  // channel offset added per h
    long long w_zstride;        // packed-weight offset per z (floats)
    int w_mode;                 // 1: B operand rows come from a c4 activation tensor (K == 1): w = tensor base
    int w_ld, w_rows, w_c_total, w_c_off, w_c_zstride;
};

namespace tc {
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
}  // namespace tc

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Bounded wait: a protocol bug must surface as a trap (CUDA error at the next sync), never as a hung GPU.
// try_wait carries a suspend-time hint so a waiting thread sleeps in hardware instead of polling: with ~10 mostly-idle
// role warps per CTA, un-hinted polling consumed > 50 % of the SM issue slots (ncu: smsp__issue_active 55 %, instruction
// mix dominated by SYNCS.TRYWAIT loops) and starved the warps doing real work.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (done) return;
        if ((it & 0x3f) == 0x3f) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) {  // ~2 s
                printf("bv2 tc_conv: mbarrier wait timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                       threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace tc

namespace tc {

// Pre-load one 128-row x nt accumulator tile: bias (+ per-batch bias) (+/- residual) (+ previous output).
// All global loads of a 32-column batch are issued before the first tcgen05.st (memory-level parallelism: the
// epilogue warps are the only threads touching residual/output tensors).
// GEN = 1: generic epilogue (polyphase ConvTranspose scatter, per-batch bias, relu); GEN = 0: plain conv epilogue.  The plain
// instantiation is 10-20 % faster on the MRF convs (measured): these kernels run at their register caps.
template <int NG, int GEN>
__device__ __forceinline__ void acc_init_tile(const TcParams& p, uint32_t trow, int b, int t, int n0, int nt, int yb = -1, int coff = -1) {
    const bool ok = t < p.T;
    const size_t tstride = (size_t)p.T * ((GEN && p.ups_u) ? p.ups_u : 1);
    if (yb < 0) yb = b;
    if (coff < 0) coff = p.cout_off;
    const float4* resb = p.res ? reinterpret_cast<const float4*>(p.res) + (size_t)yb * (p.res_C_total / 4) * tstride : nullptr;
    const float4* ybp = reinterpret_cast<const float4*>(p.y) + (size_t)yb * (p.Cout_total / 4) * tstride;
    const float sg = p.res_mode == 1 ? 1.f : -1.f;  // mode 2: tail negates -> res - (conv + bias)
    for (int col0 = 0; col0 < nt; col0 += 4 * NG) {
        float4 o[NG];
        int cos[NG], tts[NG];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int n = n0 + col0 + 4 * g;
            int co = n, tt = t;
            if (GEN && p.ups_u) { const int r = n / p.ups_cout; co = n - r * p.ups_cout; tt = t * p.ups_u + r; }
            cos[g] = co; tts[g] = tt;
            if (col0 + 4 * g < nt) {
                o[g] = p.bias ? *reinterpret_cast<const float4*>(p.bias + co) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (GEN && p.bias_b) {
                    const float4 b2 = *reinterpret_cast<const float4*>(p.bias_b + (size_t)b * p.bias_b_stride + co);
                    o[g].x += b2.x; o[g].y += b2.y; o[g].z += b2.z; o[g].w += b2.w;
                }
            }
        }
        if (ok && p.res_mode) {
            float4 r[NG];
#pragma unroll
            for (int g = 0; g < NG; g++) if (col0 + 4 * g < nt) r[g] = resb[(size_t)((p.res_c_off + cos[g]) / 4) * tstride + tts[g]];
#pragma unroll
            for (int g = 0; g < NG; g++) if (col0 + 4 * g < nt) { o[g].x += sg * r[g].x; o[g].y += sg * r[g].y; o[g].z += sg * r[g].z; o[g].w += sg * r[g].w; }
        }
        if (ok && p.accumulate) {
            float4 a[NG];
#pragma unroll
            for (int g = 0; g < NG; g++) if (col0 + 4 * g < nt) a[g] = ybp[(size_t)((coff + cos[g]) / 4) * tstride + tts[g]];
#pragma unroll
            for (int g = 0; g < NG; g++) if (col0 + 4 * g < nt) { o[g].x += a[g].x; o[g].y += a[g].y; o[g].z += a[g].z; o[g].w += a[g].w; }
        }
#pragma unroll
        for (int h = 0; h < NG / 4; h++) {
            if (col0 + 16 * h < nt) {
                uint32_t v[16];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    v[4 * g] = __float_as_uint(o[4 * h + g].x); v[4 * g + 1] = __float_as_uint(o[4 * h + g].y);
                    v[4 * g + 2] = __float_as_uint(o[4 * h + g].z); v[4 * g + 3] = __float_as_uint(o[4 * h + g].w);
                }
                tmem_st16(trow + (uint32_t)(col0 + 16 * h), v);
            }
        }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Drain one accumulator tile: TMEM -> [relu] -> scale/mask -> c4 global (16-byte stores, coalesced across a warp).
template <int NG, int GEN>
__device__ __forceinline__ void acc_tail_tile(const TcParams& p, uint32_t trow, int b, int t, int n0, int nt, int len, int yb = -1, int coff = -1) {
    const bool ok = t < p.T;
    const size_t tstride = (size_t)p.T * ((GEN && p.ups_u) ? p.ups_u : 1);
    if (yb < 0) yb = b;
    if (coff < 0) coff = p.cout_off;
    float4* ybp = reinterpret_cast<float4*>(p.y) + (size_t)yb * (p.Cout_total / 4) * tstride;
    const float s = ((p.out_mask && t >= len) ? 0.f : p.out_scale) * (p.res_mode == 2 ? -1.f : 1.f);
    for (int col0 = 0; col0 < nt; col0 += 4 * NG) {
        uint32_t v[NG / 4][16];
#pragma unroll
        for (int h = 0; h < NG / 4; h++) if (col0 + 16 * h < nt) tmem_ld16(trow + (uint32_t)(col0 + 16 * h), v[h]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!ok) continue;
#pragma unroll
        for (int h = 0; h < NG / 4; h++) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                if (col0 + 16 * h + 4 * g < nt) {
                    const int n = n0 + col0 + 16 * h + 4 * g;
                    int co = n, tt = t;
                    if (GEN && p.ups_u) { const int r = n / p.ups_cout; co = n - r * p.ups_cout; tt = t * p.ups_u + r; }
                    float4 o = make_float4(__uint_as_float(v[h][4 * g]), __uint_as_float(v[h][4 * g + 1]), __uint_as_float(v[h][4 * g + 2]),
                                           __uint_as_float(v[h][4 * g + 3]));
                    if (GEN && p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    o.x *= s; o.y *= s; o.z *= s; o.w *= s;
                    if (p.out_tf32) { o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w); }
                    ybp[(size_t)((coff + co) / 4) * tstride + tt] = o;
                }
            }
        }
    }
}

}  // namespace tc

namespace tc {
// Operand prologue of one staged activation chunk [ncg][R][4] (generic proxy): leaky-relu + RN-TF32, zero rows outside
// [r_lo, r_hi).  The stage is contiguous, so the loop runs over flat 16-byte elements with 4 independent load->store
// chains per thread (the un-unrolled per-row loop exposed the full LDS latency on every element).
__device__ __forceinline__ void xform_stage(float4* A, int ncg, int R, int r_lo, int r_hi, float slope, int tid2) {
    const int total = ncg * R;
    int r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) r[u] = (tid2 + 128 * u) % R;
    const int step = 512 % R;  // row advance per iteration (512 elements), R >= 128
    const int wraps = 512 / R;
    (void)wraps;
    for (int i0 = tid2; i0 < total; i0 += 512) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + 128 * u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < total && r[u] >= r_lo && r[u] < r_hi) v[u] = A[i];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + 128 * u;
            if (i < total) {
                float4 o;
                o.x = to_tf32(lrelu(v[u].x, slope)); o.y = to_tf32(lrelu(v[u].y, slope));
                o.z = to_tf32(lrelu(v[u].z, slope)); o.w = to_tf32(lrelu(v[u].w, slope));
                A[i] = o;
            }
            r[u] += step; if (r[u] >= R) r[u] -= R;
        }
    }
}
}  // namespace tc

// grid: (M blocks of MT*128 time steps, N tiles, B)
//
// Accumulator-init fusion: before the first MMA the epilogue warps pre-load  bias (+ per-batch bias) (+/- residual)
// (+ previous output when accumulating)  into the TMEM accumulator with tcgen05.st while the first TMA loads are in
// flight; every MMA then accumulates, and the tail is only  TMEM -> [relu] -> scale/mask -> store.
template <int GEN, int X3 = 0>
__global__ void __launch_bounds__(224, 4) k_tc_conv1d(TcParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = p.MT;
    const int t0 = blockIdx.x * 128 * MT, n0 = blockIdx.y * p.nt, z = blockIdx.z;
    const int zs = p.zsplit > 1 ? p.zsplit : 1;
    const int b = z / zs, hz = z - b * zs;
    const int xb = p.x_batch_z ? z : b, yb = p.y_batch_z ? z : b;
    const int cin_off = p.cin_off + hz * p.x_c_zstride, cout_off = p.cout_off + hz * p.y_c_zstride;
    const int nt = p.nt;
    uint8_t* sA = smem;
    const int NAS = p.nas;
    uint8_t* sW = smem + (size_t)NAS * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)p.nws * p.w_stage_bytes);
    // barrier map: a_full[NAS], a_ready[NAS], a_empty[NAS], w_full[nws], w_empty[nws], acc_full, acc_init
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_AFULL = 0, B_AREADY = NAS, B_AEMPTY = 2 * NAS, B_WFULL = 3 * NAS, B_WEMPTY = 3 * NAS + p.nws, B_ACC = 3 * NAS + 2 * p.nws,
              B_INIT = B_ACC + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_INIT + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NAS; i++) { mbar_init(BAR(B_AFULL + i), 1); mbar_init(BAR(B_AREADY + i), 128); mbar_init(BAR(B_AEMPTY + i), 1); }
        for (int i = 0; i < p.nws; i++) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
        mbar_init(BAR(B_ACC), 1);
        mbar_init(BAR(B_INIT), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the
    // previous kernel in the stream; from here on this grid reads activations that kernel produced.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int R = p.R, ncg = p.KC / 4;
    const int len = p.lens ? p.lens[b] : p.T;
    // rows r of the staged tile map to t = t0 - pad + r; rows outside [0, T) are the conv's zero padding
    const int r_lo = max(0, p.pad - t0);
    const int r_hi = min(R, p.T - (t0 - p.pad));
    const int r_mask_hi = p.in_mask ? min(r_hi, len - (t0 - p.pad)) : r_hi;  // rows >= this read as zero (x * x_mask)

    if (warp == 0) {
        // ===== activation producer: lane 0 owns the mbarrier protocol, lanes 0..ncg-1 each issue one TMA bulk copy (one
        // contiguous run per channel group) so a chunk's copies are issued in parallel instead of serially by one thread
        const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
        for (int c = 0; c < p.nchunks; c++) {
            const int sa = c % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((c / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg);
            }
            __syncwarp();
            if (lane < ncg) {
                const float* src = p.x + (((size_t)xb * (p.Cin_total / 4) + cin_off / 4 + (size_t)c * ncg + lane) * p.T + (t0 - p.pad + r_lo)) * 4;
                const uint32_t dst = smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u;
                bulk_g2s(dst, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 6) {
        if (p.w_mode) {
            // B operand = rows n0.. of a c4 activation tensor (attention keys): lanes issue one bulk copy per channel group
            const int nvalid = max(0, min(nt, p.w_rows - n0));
            const uint32_t rb = (uint32_t)nvalid * 16u;
            const float* wb = p.w + (((size_t)b * (p.w_c_total / 4) + (p.w_c_off + hz * p.w_c_zstride) / 4) * p.w_ld + n0) * 4;
            for (int c = 0; c < p.nchunks; c++) {
                const int sw = c % p.nws;
                if (lane == 0) {
                    mbar_wait(BAR(B_WEMPTY + sw), ((c / p.nws) & 1) ^ 1);
                    mbar_expect_tx(BAR(B_WFULL + sw), rb * ncg);
                }
                __syncwarp();
                if (nvalid && lane < ncg)
                    bulk_g2s(smem_u32(sW + (size_t)sw * p.w_stage_bytes) + (uint32_t)lane * nt * 16u, wb + (size_t)(c * ncg + lane) * p.w_ld * 4, rb,
                             BAR(B_WFULL + sw));
            }
        } else if (lane == 0) {
            // ===== weight producer: its own thread so the weight ring runs ahead across chunk boundaries
            int wi = 0;
            if (!p.w_mode) {
                const float* wtile = p.w + (size_t)z * p.w_zstride + (size_t)blockIdx.y * p.nchunks * p.K * (p.w_stage_bytes / 4);
                for (int c = 0; c < p.nchunks; c++) {
                    for (int j = 0; j < p.K; j++, wi++) {
                        const int sw = wi % p.nws;
                        mbar_wait(BAR(B_WEMPTY + sw), ((wi / p.nws) & 1) ^ 1);
                        mbar_expect_tx(BAR(B_WFULL + sw), p.w_stage_bytes);
                        bulk_g2s(smem_u32(sW + (size_t)sw * p.w_stage_bytes), wtile + ((size_t)c * p.K + j) * (p.w_stage_bytes / 4), p.w_stage_bytes,
                                 BAR(B_WFULL + sw));
                    }
                }
            } else {
                // B operand = rows n0.. of a c4 activation tensor (attention keys): one bulk copy per channel group
                const int nvalid = max(0, min(nt, p.w_rows - n0));
                const uint32_t rb = (uint32_t)nvalid * 16u;
                const float* wb = p.w + (((size_t)b * (p.w_c_total / 4) + (p.w_c_off + hz * p.w_c_zstride) / 4) * p.w_ld + n0) * 4;
                for (int c = 0; c < p.nchunks; c++, wi++) {
                    const int sw = wi % p.nws;
                    mbar_wait(BAR(B_WEMPTY + sw), ((wi / p.nws) & 1) ^ 1);
                    mbar_expect_tx(BAR(B_WFULL + sw), rb * ncg);
                    const uint32_t dst = smem_u32(sW + (size_t)sw * p.w_stage_bytes);
                    if (nvalid)
                        for (int g = 0; g < ncg; g++)
                            bulk_g2s(dst + (uint32_t)g * nt * 16u, wb + (size_t)(c * ncg + g) * p.w_ld * 4, rb, BAR(B_WFULL + sw));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: per weight tile, MT x KC/8 tcgen05.mma (M=128, N=nt, K=8 tf32), always accumulating.
            // Descriptors are advanced with 64-bit adds on the (addr >> 4) field: this single thread is the issue
            // bottleneck for narrow N, so the loop body is kept to a handful of integer instructions.
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);  // two channel groups per MMA
            const uint64_t a_lo = (uint64_t)((uint32_t)ncg * R), w_lo = (uint64_t)((uint32_t)ncg * nt);     // x3: lo halves
            const int nk = p.KC / 8;
            int wi = 0;
            mbar_wait(BAR(B_INIT), 0);
            fence_after();
            for (int c = 0; c < p.nchunks; c++) {
                const int sa = c % NAS;
                mbar_wait(BAR(B_AREADY + sa), (c / NAS) & 1);
                fence_after();
                const uint64_t a_desc0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes), a_lbo, 128u);
                for (int j = 0; j < p.K; j++, wi++) {
                    const int sw = wi % p.nws;
                    mbar_wait(BAR(B_WFULL + sw), (wi / p.nws) & 1);
                    fence_after();
                    const uint64_t b_desc0 = make_desc(smem_u32(sW + (size_t)sw * p.w_stage_bytes), b_lbo, 128u);
                    for (int mt = 0; mt < MT; mt++) {
                        uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(mt * 128 + j * p.dil), bd = b_desc0;
                        const uint32_t d = tmem + (uint32_t)(mt * nt);
                        if (!X3) {
                            for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(d, ad, bd, p.idesc, 1u);
                        } else {
                            // a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo   (lo*lo ~ 2^-22 relative, dropped)
                            for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) {
                                umma_tf32(d, ad + a_lo, bd, p.idesc, 1u);
                                umma_tf32(d, ad, bd + w_lo, p.idesc, 1u);
                                umma_tf32(d, ad, bd, p.idesc, 1u);
                            }
                        }
                    }
                    umma_commit(BAR(B_WEMPTY + sw));
                }
                umma_commit(BAR(B_AEMPTY + sa));
            }
            umma_commit(BAR(B_ACC));
        }
    } else {
        const int tid2 = threadIdx.x - 64;
        const int q = warp & 3;
        // ===== accumulator init (overlaps the first TMA loads)
        for (int mt = 0; mt < MT; mt++)
            acc_init_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * nt), b, t0 + mt * 128 + q * 32 + lane, n0, nt, yb, cout_off);
        fence_before();
        mbar_arrive(BAR(B_INIT));
        // ===== operand prologue on the staged tile (generic proxy), then hand over to the async proxy
        const float slope = p.in_slope;
        for (int c = 0; c < p.nchunks; c++) {
            const int sa = c % NAS;
            mbar_wait(BAR(B_AFULL + sa), (c / NAS) & 1);
            float4* A = reinterpret_cast<float4*>(sA + (size_t)sa * p.a_stage_bytes);
            if (!X3 && !p.skip_xform) xform_stage(A, ncg, R, r_lo, r_mask_hi, slope, tid2);
            for (int g = 0; g < ((p.skip_xform || !X3) ? 0 : ncg); g++) {
                float4* Ag = A + (size_t)g * R;
                for (int r = tid2; r < R; r += 128) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), lo = v;
                    if (r >= r_lo && r < r_mask_hi) {
                        const float4 a = Ag[r];
                        const float ax = lrelu(a.x, slope), ay = lrelu(a.y, slope), az = lrelu(a.z, slope), aw = lrelu(a.w, slope);
                        v.x = to_tf32(ax); v.y = to_tf32(ay); v.z = to_tf32(az); v.w = to_tf32(aw);
                        if (X3) { lo.x = to_tf32(ax - v.x); lo.y = to_tf32(ay - v.y); lo.z = to_tf32(az - v.z); lo.w = to_tf32(aw - v.w); }
                    }
                    Ag[r] = v;
                    if (X3) Ag[(size_t)ncg * R + r] = lo;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_AREADY + sa));
        }
        // ===== tail
        mbar_wait(BAR(B_ACC), 0);
        fence_after();
        for (int mt = 0; mt < MT; mt++)
            acc_tail_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * nt), b, t0 + mt * 128 + q * 32 + lane, n0, nt, len, yb, cout_off);
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent variant for narrow layers (Cin <= 32: Generator stages 3/4 = 36 of the 90 MRF convs, the last ups).
// Those launches have thousands of 128-row tiles with only K * Cin/8 tiny MMAs each, so the one-tile-per-CTA kernel
// is bound by its per-tile latency chain (launch, TMEM alloc, barrier init, TMA round trip, tail).  Here each CTA
//   * keeps ALL weight taps resident in shared memory (<= 45 KB, loaded once),
//   * walks tiles blockIdx.x, +gridDim.x, ... with a 3-deep TMA ring for the activation tiles (producer runs ahead),
//   * double-buffers the TMEM accumulator: the epilogue warps pre-load tile i+1's accumulator (bias/residual) and
//     drain tile i-1 while the MMA warp works on tile i.
// 320 threads: warp 0 producer, warp 1 MMA issuer, warps 2-5 operand prologue, warps 6-9 accumulator init + tail.
template <int GEN, int OCC = 3>
__global__ void __launch_bounds__(320, OCC) k_tc_conv1d_persist(TcParams p, int mtiles, int ntiles_total) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = p.nt, NAS = p.nas, R = p.R, ncg = p.KC / 4;
    uint8_t* sWt = smem;                                   // resident weights: [K][ncg][nt][4]
    const uint32_t w_bytes = (uint32_t)(p.K * p.KC * nt * 4);
    uint8_t* sA = smem + ((w_bytes + 127u) & ~127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)NAS * p.a_stage_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_WFULL = 0, B_AFULL = 1, B_AREADY = 1 + NAS, B_AEMPTY = 1 + 2 * NAS, B_INIT = 1 + 3 * NAS, B_ACC = 3 + 3 * NAS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACC + 2);

    if (threadIdx.x == 0) {
        mbar_init(BAR(B_WFULL), 1);
        for (int i = 0; i < NAS; i++) { mbar_init(BAR(B_AFULL + i), 1); mbar_init(BAR(B_AREADY + i), 128); mbar_init(BAR(B_AEMPTY + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(BAR(B_INIT + i), 128); mbar_init(BAR(B_ACC + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the
    // previous kernel in the stream; from here on this grid reads activations that kernel produced.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int n_mine = (ntiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(BAR(B_WFULL), w_bytes);
            bulk_g2s(smem_u32(sWt), p.w, w_bytes, BAR(B_WFULL));
        }
        for (int i = 0; i < n_mine; i++) {
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            const int r_lo = max(0, p.pad - t0), r_hi = min(R, p.T - (t0 - p.pad));
            const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
            const int sa = i % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((i / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg);
            }
            __syncwarp();
            if (lane < ncg) {
                const float* src = p.x + (((size_t)b * (p.Cin_total / 4) + p.cin_off / 4 + lane) * p.T + (t0 - p.pad + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);
            const uint64_t b_tap = (uint64_t)((uint32_t)ncg * nt);  // next tap's weight tile, in 16-byte units
            const int nk = p.KC / 8;
            const uint64_t b_desc0 = make_desc(smem_u32(sWt), b_lbo, 128u);
            mbar_wait(BAR(B_WFULL), 0);
            for (int i = 0; i < n_mine; i++) {
                const int sa = i % NAS, ab = i & 1;
                mbar_wait(BAR(B_INIT + ab), (i >> 1) & 1);
                mbar_wait(BAR(B_AREADY + sa), (i / NAS) & 1);
                fence_after();
                const uint64_t a_desc0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes), a_lbo, 128u);
                const uint32_t d = tmem + (uint32_t)(ab * nt);
                uint64_t bd_tap = b_desc0;
                for (int j = 0; j < p.K; j++, bd_tap += b_tap) {
                    uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(j * p.dil), bd = bd_tap;
                    for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(d, ad, bd, p.idesc, 1u);
                }
                umma_commit(BAR(B_AEMPTY + sa));
                umma_commit(BAR(B_ACC + ab));
            }
        }
    } else if (warp < 6) {
        // ===== operand prologue
        const int tid2 = threadIdx.x - 64;
        const float slope = p.in_slope;
        for (int i = 0; i < n_mine; i++) {
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            const int len = p.lens ? p.lens[b] : p.T;
            const int r_lo = max(0, p.pad - t0), r_hi = min(R, p.T - (t0 - p.pad));
            const int r_mask_hi = p.in_mask ? min(r_hi, len - (t0 - p.pad)) : r_hi;
            const int sa = i % NAS;
            mbar_wait(BAR(B_AFULL + sa), (i / NAS) & 1);
            float4* A = reinterpret_cast<float4*>(sA + (size_t)sa * p.a_stage_bytes);
            xform_stage(A, ncg, R, r_lo, r_mask_hi, slope, tid2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_AREADY + sa));
        }
    } else {
        // ===== accumulator init (tile i+1) and tail (tile i), double-buffered TMEM
        const int q = warp & 3;
        auto init_tile = [&](int i) {
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            const int n0 = 0;
            acc_init_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * nt), b, t0 + q * 32 + lane, n0, nt);
            fence_before();
            mbar_arrive(BAR(B_INIT + (i & 1)));
        };
        if (n_mine > 0) init_tile(0);
        for (int i = 0; i < n_mine; i++) {
            if (i + 1 < n_mine) init_tile(i + 1);
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            const int n0 = 0;
            const int len = p.lens ? p.lens[b] : p.T;
            mbar_wait(BAR(B_ACC + (i & 1)), (i >> 1) & 1);
            fence_after();
            acc_tail_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * nt), b, t0 + q * 32 + lane, n0, nt, len);
            fence_before();  // order this tile's tcgen05.ld before the next init's tcgen05.st on the same columns
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent variant with STREAMED weights for wide layers (Cin >= 64): the canonical Blackwell GEMM structure.
// One CTA per SM walks tiles (n-tile fastest, so CTAs that share an activation tile run together and hit L2); the
// TMA producer runs continuously over the flat (tile, chunk, tap) sequence, so the activation ring (NAS deep) and
// the weight ring (nws deep) stay full across tile boundaries; the TMEM accumulator is double-buffered so the
// epilogue warps pre-load tile i+1's accumulator and drain tile i-1 while the MMA warp is busy with tile i.
// 352 threads: warp 0 activation producer, warp 1 MMA issuer, warps 2-5 operand prologue, warps 6-9 accumulator init +
// tail, warp 10 weight producer.
template <int GEN>
__global__ void __launch_bounds__(352, 2) k_tc_conv1d_pstream(TcParams p, int mtiles, int ntiles, int tiles_total) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = p.nt, NAS = p.nas, NWS = p.nws, R = p.R, ncg = p.KC / 4, NCH = p.nchunks;
    uint8_t* sA = smem;
    uint8_t* sW = smem + (size_t)NAS * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)NWS * p.w_stage_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_AFULL = 0, B_AREADY = NAS, B_AEMPTY = 2 * NAS, B_WFULL = 3 * NAS, B_WEMPTY = 3 * NAS + NWS, B_INIT = 3 * NAS + 2 * NWS,
              B_ACC = B_INIT + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACC + 2);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NAS; i++) { mbar_init(BAR(B_AFULL + i), 1); mbar_init(BAR(B_AREADY + i), 128); mbar_init(BAR(B_AEMPTY + i), 1); }
        for (int i = 0; i < NWS; i++) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(BAR(B_INIT + i), 128); mbar_init(BAR(B_ACC + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the
    // previous kernel in the stream; from here on this grid reads activations that kernel produced.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int n_mine = (tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // tile -> (batch, m tile, n tile); n tile fastest
    auto decode = [&](int i, int& b, int& t0, int& ntile) {
        const int tile = blockIdx.x + i * gridDim.x;
        ntile = tile % ntiles;
        const int mm = tile / ntiles;
        b = mm / mtiles;
        t0 = (mm - b * mtiles) * 128;
    };

    if (warp == 0) {
        // ===== activation producer: runs up to NAS (tile, chunk) steps ahead of the MMA warp; copies issued by ncg lanes
        const int steps = n_mine * NCH;  // flat (tile, chunk) sequence
        for (int s_ = 0; s_ < steps; s_++) {
            int b, t0, ntile;
            decode(s_ / NCH, b, t0, ntile);
            const int c = s_ % NCH;
            const int r_lo = max(0, p.pad - t0), r_hi = min(R, p.T - (t0 - p.pad));
            const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
            const int sa = s_ % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((s_ / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg);
            }
            __syncwarp();
            if (lane < ncg) {
                const float* src = p.x + (((size_t)b * (p.Cin_total / 4) + p.cin_off / 4 + (size_t)c * ncg + lane) * p.T + (t0 - p.pad + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {
            // ===== weight producer: independent thread, so weight tiles stream up to NWS taps ahead across chunk and tile
            // boundaries (a single producer would serialise on the activation ring and bubble at every chunk)
            const int steps = n_mine * NCH;
            int wi = 0;
            const size_t wstage_f = p.w_stage_bytes / 4;
            for (int s_ = 0; s_ < steps; s_++) {
                int b, t0, ntile;
                decode(s_ / NCH, b, t0, ntile);
                const int c = s_ % NCH;
                const float* wsrc = p.w + ((size_t)ntile * NCH + c) * p.K * wstage_f;
                for (int j = 0; j < p.K; j++, wi++) {
                    const int sw = wi % NWS;
                    mbar_wait(BAR(B_WEMPTY + sw), ((wi / NWS) & 1) ^ 1);
                    mbar_expect_tx(BAR(B_WFULL + sw), p.w_stage_bytes);
                    bulk_g2s(smem_u32(sW + (size_t)sw * p.w_stage_bytes), wsrc + (size_t)j * wstage_f, p.w_stage_bytes, BAR(B_WFULL + sw));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);
            const int nk = p.KC / 8;
            int wi = 0, s_ = 0;
            for (int i = 0; i < n_mine; i++) {
                const int ab = i & 1;
                mbar_wait(BAR(B_INIT + ab), (i >> 1) & 1);
                fence_after();
                const uint32_t d = tmem + (uint32_t)(ab * nt);
                for (int c = 0; c < NCH; c++, s_++) {
                    const int sa = s_ % NAS;
                    mbar_wait(BAR(B_AREADY + sa), (s_ / NAS) & 1);
                    fence_after();
                    const uint64_t a_desc0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes), a_lbo, 128u);
                    for (int j = 0; j < p.K; j++, wi++) {
                        const int sw = wi % NWS;
                        mbar_wait(BAR(B_WFULL + sw), (wi / NWS) & 1);
                        fence_after();
                        uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(j * p.dil);
                        uint64_t bd = make_desc(smem_u32(sW + (size_t)sw * p.w_stage_bytes), b_lbo, 128u);
                        for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(d, ad, bd, p.idesc, 1u);
                        umma_commit(BAR(B_WEMPTY + sw));
                    }
                    umma_commit(BAR(B_AEMPTY + sa));
                }
                umma_commit(BAR(B_ACC + ab));
            }
        }
    } else if (warp < 6) {
        // ===== operand prologue over the flat (tile, chunk) sequence
        const int tid2 = threadIdx.x - 64;
        const float slope = p.in_slope;
        const int steps = n_mine * NCH;
        for (int s_ = 0; s_ < steps; s_++) {
            int b, t0, ntile;
            decode(s_ / NCH, b, t0, ntile);
            const int len = p.lens ? p.lens[b] : p.T;
            const int r_lo = max(0, p.pad - t0), r_hi = min(R, p.T - (t0 - p.pad));
            const int r_mask_hi = p.in_mask ? min(r_hi, len - (t0 - p.pad)) : r_hi;
            const int sa = s_ % NAS;
            mbar_wait(BAR(B_AFULL + sa), (s_ / NAS) & 1);
            float4* A = reinterpret_cast<float4*>(sA + (size_t)sa * p.a_stage_bytes);
            xform_stage(A, ncg, R, r_lo, r_mask_hi, slope, tid2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_AREADY + sa));
        }
    } else {
        // ===== accumulator init (tile i+1) and tail (tile i), double-buffered TMEM
        const int q = warp & 3;
        auto init_tile = [&](int i) {
            int b, t0, ntile;
            decode(i, b, t0, ntile);
            const int n0 = ntile * nt;
            acc_init_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * nt), b, t0 + q * 32 + lane, n0, nt);
            fence_before();
            mbar_arrive(BAR(B_INIT + (i & 1)));
        };
        if (n_mine > 0) init_tile(0);
        for (int i = 0; i < n_mine; i++) {
            if (i + 1 < n_mine) init_tile(i + 1);
            int b, t0, ntile;
            decode(i, b, t0, ntile);
            const int n0 = ntile * nt;
            const int len = p.lens ? p.lens[b] : p.T;
            mbar_wait(BAR(B_ACC + (i & 1)), (i >> 1) & 1);
            fence_after();
            acc_tail_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * nt), b, t0 + q * 32 + lane, n0, nt, len);
            fence_before();  // order this tile's tcgen05.ld before the next init's tcgen05.st on the same columns
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// x: c4 input [B][x.C/4][T][4]; y: c4 output ([B][y.C/4][T*max(1,ups_u)][4]).  Channel windows via e.cin_off/e.cout_off.
inline void tc_conv1d(const TcConvW& w, const float* bias, const Act& x, const Act& y, const TcEpi& e, cudaStream_t st, int num_sms) {
    const int u = w.ups_u ? w.ups_u : 1;
    BV2_CHECK(w.w && x.B == y.B && y.T == x.T * u, "tc_conv1d shapes");
    BV2_CHECK(e.cin_off % 4 == 0 && e.cout_off % 4 == 0 && e.cin_off + w.Cin <= x.C, "tc_conv1d channel window");
    TcParams p{};
    p.x = x.p; p.y = y.p; p.w = w.w; p.bias = bias; p.res = e.res; p.bias_b = e.bias_b; p.lens = e.lens;
    p.Cin_total = x.C; p.cin_off = e.cin_off; p.Cout_total = y.C; p.cout_off = e.cout_off;
    p.res_C_total = e.res_C_total ? e.res_C_total : y.C; p.res_c_off = e.res_c_off; p.bias_b_stride = e.bias_b_stride;
    p.T = x.T; p.B = x.B; p.K = w.K; p.dil = e.dil; p.pad = (w.K - 1) / 2 * e.dil;
    p.KC = w.KC; p.nchunks = w.nchunks;
    const int nt = w.nt;
    if (w.ups_u) BV2_CHECK(w.ups_cout % 4 == 0, "ups cout");
    p.nt = nt;
    const int ntiles = w.Cout / nt;
    const int MT = 1;  // weight-tile sharing across M tiles: measured slower than more resident CTAs at these sizes
    const int halo = (w.K - 1) * e.dil;
    p.MT = MT;
    p.R = MT * 128 + halo;
    p.x3 = w.x3;
    const int parts = w.x3 ? 2 : 1;
    p.a_stage_bytes = (uint32_t)(p.KC * p.R * 4 * parts);
    p.w_stage_bytes = (uint32_t)(p.KC * nt * 4 * parts);
    // shared memory per CTA is capped (~100 KB) so that two CTAs co-reside per SM: one CTA's accumulator init / tail
    // overlaps the other's MMA main loop
    const long long nctas = (long long)cdiv(p.T, 128 * MT) * ntiles * p.B;
    static const int smem_kb_env = getenv("BV2_TC_SMEM_KB") ? atoi(getenv("BV2_TC_SMEM_KB")) : 48;  // tuning knob (experiments)
    uint32_t budget = (nctas > num_sms && nt <= 128) ? (uint32_t)smem_kb_env * 1024 : 200 * 1024;
    if (nt > 128 && 2 * nctas > num_sms && 2ull * p.a_stage_bytes + 2ull * p.w_stage_bytes + 2048 <= 112 * 1024)
        budget = 112 * 1024;  // wide layer launched on three streams at once (MRF resblock chains): let two CTAs share an SM
    // activation pipeline depth: up to 4 stages when the K loop is long (hides TMA + prologue latency per chunk)
    int nas = std::min(3, std::max(2, p.nchunks));
    while (nas > 2 && (size_t)nas * p.a_stage_bytes + 4 * (size_t)p.w_stage_bytes + 1024 > budget) nas--;
    p.nas = nas;
    int nws = ((int)budget - nas * (int)p.a_stage_bytes - 1024) / (int)p.w_stage_bytes;
    p.nws = std::max(2, std::min(nws, 8));
    uint32_t cols = 32; while ((int)cols < MT * nt) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((128u >> 4) << 24);
    p.in_slope = e.in_slope; p.out_scale = e.out_scale; p.accumulate = e.accumulate; p.relu = e.relu; p.res_mode = e.res ? (e.res_mode ? e.res_mode : 1) : 0;
    p.in_mask = e.in_mask; p.out_mask = e.out_mask; p.ups_u = w.ups_u; p.ups_cout = w.ups_cout;
    p.out_tf32 = e.out_tf32; p.skip_xform = e.skip_xform;
    if (e.skip_xform) BV2_CHECK(w.K == 1 && e.in_slope == 1.f && !e.in_mask, "skip_xform needs a plain 1x1 conv input");
    if (p.in_mask || p.out_mask) BV2_CHECK(e.lens != nullptr, "mask needs lens");
    BV2_CHECK(!(p.relu && (p.res_mode || p.accumulate)), "relu cannot be combined with residual/accumulate (accumulator-init fusion)");
    const size_t smem = (size_t)p.nas * p.a_stage_bytes + (size_t)p.nws * p.w_stage_bytes + (size_t)(3 * p.nas + 2 * p.nws + 2) * 8 + 16;
    BV2_CHECK(smem <= 227 * 1024, "tc_conv1d shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const bool generic = w.ups_u || e.bias_b || e.relu;
    static const int persist_env = getenv("BV2_TC_PERSIST") ? atoi(getenv("BV2_TC_PERSIST")) : 1;
    const size_t w_all = (size_t)p.K * p.KC * nt * 4;
    if (persist_env && !e.skip_xform && p.nchunks == 1 && ntiles == 1 && !w.x3 && w_all <= 64 * 1024 && nctas >= 2 * num_sms) {
        // narrow layer with many tiles: persistent CTAs, resident weights, double-buffered TMEM
        p.nas = 3;
        const size_t wb = (w_all + 127) & ~(size_t)127;
        const size_t smem_p = wb + (size_t)p.nas * p.a_stage_bytes + (size_t)(3 * p.nas + 5) * 8 + 16;
        uint32_t pc = 32; while ((int)pc < 2 * nt) pc <<= 1;
        p.tmem_cols = pc;
        static bool attr2 = false;
        if (!attr2) {
            BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d_persist<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d_persist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d_persist<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr2 = true;
        }
        static const int occ_env = getenv("BV2_PERSIST_OCC") ? atoi(getenv("BV2_PERSIST_OCC")) : 2;  // measured: 2 CTAs/SM with 102 registers (no spills) beats 3 with 68
        int per_sm = smem_p <= 72 * 1024 ? 3 : (smem_p <= 110 * 1024 ? 2 : 1);
        if (occ_env == 2 && !generic) per_sm = std::min(per_sm, 2);
        const int mtiles = cdiv(p.T, 128);
        const int total = mtiles * p.B;
        const int grid_p = std::min(total, per_sm * num_sms);
        if (occ_env == 2 && !generic) launch_pdl(k_tc_conv1d_persist<0, 2>, dim3(grid_p), dim3(320), smem_p, st, p, mtiles, total);
        else launch_pdl(generic ? k_tc_conv1d_persist<1> : k_tc_conv1d_persist<0>, dim3(grid_p), dim3(320), smem_p, st, p, mtiles, total);
        return;
    }
    static const int pstream_env = getenv("BV2_TC_PSTREAM") ? atoi(getenv("BV2_TC_PSTREAM")) : 1;
    if (pstream_env && !e.skip_xform && !w.x3 && nctas >= num_sms && nt >= 128 && 2 * nt <= 512) {  // measured: wins for wide N tiles only
        // wide layer with at least one tile per SM: persistent CTAs, continuously streamed weights, double-buffered TMEM
        const bool two_per_sm = 2 * nt <= 256 && (size_t)2 * p.a_stage_bytes + 3 * (size_t)p.w_stage_bytes + 2048 <= 104 * 1024;
        const uint32_t big = two_per_sm ? 104 * 1024 : 200 * 1024;
        int nas2 = std::min(4, std::max(2, p.nchunks * 2));
        while (nas2 > 2 && (size_t)nas2 * p.a_stage_bytes + 3 * (size_t)p.w_stage_bytes + 2048 > big) nas2--;
        int nws2 = (int)((big - (size_t)nas2 * p.a_stage_bytes - 2048) / p.w_stage_bytes);
        nws2 = std::max(2, std::min(nws2, 8));
        p.nas = nas2; p.nws = nws2;
        uint32_t pc = 32; while ((int)pc < 2 * nt) pc <<= 1;
        p.tmem_cols = pc;
        const size_t smem_s = (size_t)nas2 * p.a_stage_bytes + (size_t)nws2 * p.w_stage_bytes + (size_t)(3 * nas2 + 2 * nws2 + 4) * 8 + 16;
        BV2_CHECK(smem_s <= 227 * 1024, "tc_conv1d pstream shared memory");
        static bool attr3 = false;
        if (!attr3) {
            BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d_pstream<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d_pstream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr3 = true;
        }
        const int mtiles = cdiv(p.T, 128);
        const int total = mtiles * p.B * ntiles;
        const int grid_s = std::min(total, (two_per_sm ? 2 : 1) * num_sms);
        launch_pdl(generic ? k_tc_conv1d_pstream<1> : k_tc_conv1d_pstream<0>, dim3(grid_s), dim3(352), smem_s, st, p, mtiles, ntiles, total);
        return;
    }
    dim3 grid(cdiv(p.T, 128 * MT), ntiles, p.B);
    if (w.x3) {  // accuracy-study instantiation (tests/cuda/tc_probe.cu); not used by the engine
        static bool a3 = false;
        if (!a3) { BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); a3 = true; }
        launch_pdl(k_tc_conv1d<1, 1>, grid, dim3(224), smem, st, p);
        return;
    }
    launch_pdl(generic ? k_tc_conv1d<1> : k_tc_conv1d<0>, grid, dim3(224), smem, st, p);
}


inline void tc_launch_simple(TcParams& p, int ntiles, int zdim, cudaStream_t st) {
    const int halo = (p.K - 1) * p.dil;
    p.MT = 1; p.R = 128 + halo; p.pad = (p.K - 1) / 2 * p.dil;
    p.a_stage_bytes = (uint32_t)(p.KC * p.R * 4);
    p.w_stage_bytes = (uint32_t)(p.KC * p.nt * 4);
    const long long nctas = (long long)cdiv(p.T, 128) * ntiles * zdim;
    const uint32_t budget = nctas > 148 ? 100 * 1024 : 200 * 1024;
    int nas = std::min(3, std::max(2, p.nchunks));
    while (nas > 2 && (size_t)nas * p.a_stage_bytes + 3 * (size_t)p.w_stage_bytes + 1024 > budget) nas--;
    p.nas = nas;
    int nws = ((int)budget - nas * (int)p.a_stage_bytes - 1024) / (int)p.w_stage_bytes;
    p.nws = std::max(2, std::min(nws, 8));
    uint32_t cols = 32; while ((int)cols < p.nt) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.nt >> 3) << 17) | ((128u >> 4) << 24);
    const size_t smem = (size_t)p.nas * p.a_stage_bytes + (size_t)p.nws * p.w_stage_bytes + (size_t)(3 * p.nas + 2 * p.nws + 2) * 8 + 16;
    BV2_CHECK(smem <= 227 * 1024, "tc gemm shared memory");
    static bool attr_set = false;
    if (!attr_set) { BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set = true; }
    dim3 grid(cdiv(p.T, 128), ntiles, zdim);
    launch_pdl(k_tc_conv1d<0>, grid, dim3(224), smem, st, p);
}

// S[z][keys][queries] (c4 over keys) = Q . K^T for every (batch, head): qkv c4 [B][3H/4][T][4], q pre-scaled.
// Keys are padded to Fp (multiple of 128); columns >= T hold garbage and are never read by the softmax.
inline void tc_attn_qk(const Act& qkv, int H, int heads, const Act& S, cudaStream_t st) {
    const int dk = H / heads;
    TcParams p{};
    p.x = qkv.p; p.y = S.p; p.w = qkv.p; p.bias = nullptr;
    p.Cin_total = qkv.C; p.cin_off = 0; p.x_c_zstride = dk; p.Cout_total = S.C; p.cout_off = 0; p.res_C_total = S.C;
    p.T = qkv.T; p.B = qkv.B; p.K = 1; p.dil = 1; p.KC = 32; p.nchunks = dk / 32; p.nt = 128;
    p.in_slope = 1.f; p.out_scale = 1.f;
    p.zsplit = heads; p.x_batch_z = 0; p.y_batch_z = 1;
    p.w_mode = 1; p.w_ld = qkv.T; p.w_rows = qkv.T; p.w_c_total = qkv.C; p.w_c_off = H; p.w_c_zstride = dk;
    p.skip_xform = 1;  // q/k/v were rounded to TF32 by the QKV projection's tail
    BV2_CHECK(dk % 32 == 0 && S.C % 128 == 0 && S.T == qkv.T && S.B == qkv.B * heads, "tc_attn_qk shapes");
    tc_launch_simple(p, S.C / 128, qkv.B * heads, st);
}

// att[b][h*dk + d][i] += sum_j P[z][j][i] * V[j][d]  (P = c4 over keys, vt = packed V^T [z][Fp/32][8][dk][4])
inline void tc_attn_pv(const Act& P, const float* vt, int H, int heads, const Act& att, cudaStream_t st) {
    const int dk = H / heads;
    TcParams p{};
    p.x = P.p; p.y = att.p; p.w = vt; p.bias = nullptr;
    p.Cin_total = P.C; p.cin_off = 0; p.Cout_total = att.C; p.cout_off = 0; p.y_c_zstride = dk; p.res_C_total = att.C;
    p.T = P.T; p.B = att.B; p.K = 1; p.dil = 1; p.KC = 64; p.nchunks = P.C / 64; p.nt = dk;  // long reduction (keys): big chunks
    p.in_slope = 1.f; p.out_scale = 1.f; p.accumulate = 1;
    p.zsplit = heads; p.x_batch_z = 1; p.y_batch_z = 0;
    p.w_mode = 0; p.w_zstride = (long long)P.C * dk;
    p.skip_xform = 1;  // P rounded by k_attn_softmax, V^T is a copy of the rounded v
    p.out_tf32 = 1;    // conv_o consumes it without a prologue
    BV2_CHECK(dk % 16 == 0 && dk <= 256 && P.C % 64 == 0 && P.B == att.B * heads && att.T == P.T, "tc_attn_pv shapes");
    tc_launch_simple(p, 1, P.B, st);
}

// ------------------------------------------------------------------------------------------------------------
// Fused ResBlock pair (reference modules.py:296-309):  y = conv2(lrelu(conv1(lrelu(x)))) + x  [(+ y_old) * scale]
// conv1: K taps, dilation d;  conv2: K taps, dilation 1;  C channels in and out.  One CTA produces TO = 128-(K-1) output
// rows: phase 1 accumulates conv1 for 128 rows in TMEM (D1), the epilogue warps turn D1 into the TF32 A-operand image
// XT[C/4][128+K-1][4] in SHARED memory (zero outside the sequence = conv2's padding), phase 2 runs conv2 straight from
// XT into a second accumulator (D2) that was pre-loaded with bias2 + residual.  The intermediate never touches HBM:
// 5 activation round trips per pair become ~2.5 and two launches become one.
struct TcPairParams {
    const float* x; float* y; const float* w1; const float* w2; const float* b1; const float* b2;
    int C, T, B, K, dil, KC, nchunks, R1, RT, TO, nws, nas;
    uint32_t a_stage_bytes, w_stage_bytes, xt_bytes, tmem_cols, idesc;
    float out_scale; int accumulate;
};

__global__ void __launch_bounds__(256, 2) k_tc_pair(TcPairParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, NAS = p.nas, NWS = p.nws, R = p.R1, RT = p.RT, ncg = p.KC / 4, NCH = p.nchunks;
    const int t0 = blockIdx.x * p.TO, b = blockIdx.y;
    const int p2 = (p.K - 1) / 2, p1 = p2 * p.dil;
    uint8_t* sA = smem;
    uint8_t* sW = sA + (size_t)NAS * p.a_stage_bytes;
    float4* XT = reinterpret_cast<float4*>(sW + (size_t)NWS * p.w_stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(XT) + p.xt_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_AFULL = 0, B_AREADY = NAS, B_AEMPTY = 2 * NAS, B_WFULL = 3 * NAS, B_WEMPTY = 3 * NAS + NWS, B_INIT = 3 * NAS + 2 * NWS,
              B_ACC1 = B_INIT + 1, B_XT = B_INIT + 2, B_ACC2 = B_INIT + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_ACC2 + 1);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NAS; i++) { mbar_init(BAR(B_AFULL + i), 1); mbar_init(BAR(B_AREADY + i), 128); mbar_init(BAR(B_AEMPTY + i), 1); }
        for (int i = 0; i < NWS; i++) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
        mbar_init(BAR(B_INIT), 128); mbar_init(BAR(B_ACC1), 1); mbar_init(BAR(B_XT), 128); mbar_init(BAR(B_ACC2), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // staged x rows r <-> t = t0 - p2 - p1 + r
    const int tx0 = t0 - p2 - p1;
    const int r_lo = max(0, -tx0), r_hi = min(R, p.T - tx0);

    if (warp == 0) {
        const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
        for (int c = 0; c < NCH; c++) {
            const int sa = c % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((c / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg);
            }
            __syncwarp();
            if (lane < ncg) {
                const float* src = p.x + (((size_t)b * (C / 4) + (size_t)c * ncg + lane) * p.T + (tx0 + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            const size_t wst = p.w_stage_bytes / 4;
            int wi = 0;
            for (int ph = 0; ph < 2; ph++) {
                const float* wsrc = ph ? p.w2 : p.w1;
                for (int c = 0; c < NCH; c++)
                    for (int j = 0; j < p.K; j++, wi++) {
                        const int sw = wi % NWS;
                        mbar_wait(BAR(B_WEMPTY + sw), ((wi / NWS) & 1) ^ 1);
                        mbar_expect_tx(BAR(B_WFULL + sw), p.w_stage_bytes);
                        bulk_g2s(smem_u32(sW + (size_t)sw * p.w_stage_bytes), wsrc + ((size_t)c * p.K + j) * wst, p.w_stage_bytes, BAR(B_WFULL + sw));
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t b_lbo = (uint32_t)C * 16u;
            const uint64_t b_kstep = (uint64_t)(2u * (uint32_t)C);
            const int nk = p.KC / 8;
            int wi = 0;
            mbar_wait(BAR(B_INIT), 0);
            fence_after();
            // ---- phase 1: D1 = conv1 over the staged, activated x chunks
            {
                const uint32_t a_lbo = (uint32_t)R * 16u;
                const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R);
                for (int c = 0; c < NCH; c++) {
                    const int sa = c % NAS;
                    mbar_wait(BAR(B_AREADY + sa), (c / NAS) & 1);
                    fence_after();
                    const uint64_t a_desc0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes), a_lbo, 128u);
                    for (int j = 0; j < p.K; j++, wi++) {
                        const int sw = wi % NWS;
                        mbar_wait(BAR(B_WFULL + sw), (wi / NWS) & 1);
                        fence_after();
                        uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(j * p.dil);
                        uint64_t bd = make_desc(smem_u32(sW + (size_t)sw * p.w_stage_bytes), b_lbo, 128u);
                        for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(tmem, ad, bd, p.idesc, 1u);
                        umma_commit(BAR(B_WEMPTY + sw));
                    }
                    umma_commit(BAR(B_AEMPTY + sa));
                }
                umma_commit(BAR(B_ACC1));
            }
            // ---- phase 2: D2 += conv2 over XT (resident in smem)
            mbar_wait(BAR(B_XT), 0);
            fence_after();
            {
                const uint32_t a_lbo = (uint32_t)RT * 16u;
                const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)RT);
                const uint64_t xt_desc0 = make_desc(smem_u32(XT), a_lbo, 128u);
                for (int c = 0; c < NCH; c++) {
                    const uint64_t a_desc0 = xt_desc0 + (uint64_t)((uint32_t)(c * ncg) * (uint32_t)RT);
                    for (int j = 0; j < p.K; j++, wi++) {
                        const int sw = wi % NWS;
                        mbar_wait(BAR(B_WFULL + sw), (wi / NWS) & 1);
                        fence_after();
                        uint64_t ad = a_desc0 + (uint64_t)(uint32_t)j;
                        uint64_t bd = make_desc(smem_u32(sW + (size_t)sw * p.w_stage_bytes), b_lbo, 128u);
                        for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(tmem + (uint32_t)C, ad, bd, p.idesc, 1u);
                        umma_commit(BAR(B_WEMPTY + sw));
                    }
                }
                umma_commit(BAR(B_ACC2));
            }
        }
    } else if (warp >= 4) {
        const int tid2 = threadIdx.x - 128;
        const int q = warp & 3, m = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        const int t_out = t0 + m;
        const bool ok_out = m < p.TO && t_out < p.T;
        const float4* xb = reinterpret_cast<const float4*>(p.x) + (size_t)b * (C / 4) * p.T;
        float4* yb = reinterpret_cast<float4*>(p.y) + (size_t)b * (C / 4) * p.T;
        // ---- accumulator init: D1 = bias1 ; D2 = bias2 + x (+ y_old)
        for (int col = 0; col < C; col += 16) {
            uint32_t v1[16], v2[16];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int cg = (col >> 2) + g;
                const float4 bb1 = *reinterpret_cast<const float4*>(p.b1 + cg * 4);
                float4 o = *reinterpret_cast<const float4*>(p.b2 + cg * 4);
                if (ok_out) {
                    const float4 r = xb[(size_t)cg * p.T + t_out];
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                    if (p.accumulate) { const float4 a = yb[(size_t)cg * p.T + t_out]; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
                }
                v1[4 * g] = __float_as_uint(bb1.x); v1[4 * g + 1] = __float_as_uint(bb1.y); v1[4 * g + 2] = __float_as_uint(bb1.z); v1[4 * g + 3] = __float_as_uint(bb1.w);
                v2[4 * g] = __float_as_uint(o.x); v2[4 * g + 1] = __float_as_uint(o.y); v2[4 * g + 2] = __float_as_uint(o.z); v2[4 * g + 3] = __float_as_uint(o.w);
            }
            tmem_st16(trow + (uint32_t)col, v1);
            tmem_st16(trow + (uint32_t)(C + col), v2);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_before();
        mbar_arrive(BAR(B_INIT));
        // ---- operand prologue of the x chunks (phase 1)
        for (int c = 0; c < NCH; c++) {
            const int sa = c % NAS;
            mbar_wait(BAR(B_AFULL + sa), (c / NAS) & 1);
            xform_stage(reinterpret_cast<float4*>(sA + (size_t)sa * p.a_stage_bytes), ncg, R, r_lo, r_hi, 0.1f, tid2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_AREADY + sa));
        }
        // ---- epilogue 1: D1 -> lrelu -> TF32 -> XT (A-operand image of conv2); rows outside the sequence are conv2's zero padding
        mbar_wait(BAR(B_ACC1), 0);
        fence_after();
        const int t_xt = t0 - p2 + m;
        const bool xt_in = t_xt >= 0 && t_xt < p.T;
        for (int col = 0; col < C; col += 16) {
            uint32_t v[16];
            tmem_ld16(trow + (uint32_t)col, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int g = 0; g < 4; g++) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (xt_in) {
                    o.x = to_tf32(lrelu(__uint_as_float(v[4 * g]), 0.1f)); o.y = to_tf32(lrelu(__uint_as_float(v[4 * g + 1]), 0.1f));
                    o.z = to_tf32(lrelu(__uint_as_float(v[4 * g + 2]), 0.1f)); o.w = to_tf32(lrelu(__uint_as_float(v[4 * g + 3]), 0.1f));
                }
                XT[(size_t)((col >> 2) + g) * RT + m] = o;
            }
        }
        if (m < p.K - 1)  // rows 128..RT-1 are only read by the discarded output rows; keep them finite
            for (int cg = 0; cg < C / 4; cg++) XT[(size_t)cg * RT + 128 + m] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fence_before();
        mbar_arrive(BAR(B_XT));
        // ---- tail: D2 -> scale -> store the TO valid rows
        mbar_wait(BAR(B_ACC2), 0);
        fence_after();
        for (int col = 0; col < C; col += 16) {
            uint32_t v[16];
            tmem_ld16(trow + (uint32_t)(C + col), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (!ok_out) continue;
#pragma unroll
            for (int g = 0; g < 4; g++)
                yb[(size_t)((col >> 2) + g) * p.T + t_out] = make_float4(__uint_as_float(v[4 * g]) * p.out_scale, __uint_as_float(v[4 * g + 1]) * p.out_scale,
                                                                         __uint_as_float(v[4 * g + 2]) * p.out_scale, __uint_as_float(v[4 * g + 3]) * p.out_scale);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// returns false when the shapes do not fit (caller falls back to two launches)
inline bool tc_pair(const TcConvW& w1, const TcConvW& w2, const float* b1, const float* b2, const Act& x, const Act& y, int dil, float out_scale,
                    int accumulate, cudaStream_t st) {
    const int C = w1.Cin;
    if (!(w1.Cout == C && w2.Cin == C && w2.Cout == C && w1.K == w2.K && w1.KC == w2.KC && w1.nt == C && w2.nt == C && !w1.x3 && !w1.ups_u && C <= 128 && (w1.K & 1)))
        return false;
    TcPairParams p{};
    p.x = x.p; p.y = y.p; p.w1 = w1.w; p.w2 = w2.w; p.b1 = b1; p.b2 = b2;
    p.C = C; p.T = x.T; p.B = x.B; p.K = w1.K; p.dil = dil; p.KC = w1.KC; p.nchunks = w1.nchunks;
    p.R1 = 128 + (p.K - 1) * dil; p.RT = 128 + p.K - 1; p.TO = 128 - (p.K - 1);
    p.a_stage_bytes = (uint32_t)(p.KC * p.R1 * 4); p.w_stage_bytes = (uint32_t)(p.KC * C * 4); p.xt_bytes = (uint32_t)(C * p.RT * 4);
    p.nas = std::min(2, std::max(2, p.nchunks));
    const long long fixed = (long long)p.nas * p.a_stage_bytes + p.xt_bytes + 1024;
    // shared-memory budgets for 3 / 2 / 1 resident CTAs per SM; take the first that leaves a >=3-deep weight ring
    int nws = 0;
    for (long long budget : {74LL * 1024, 112LL * 1024, 224LL * 1024}) {
        nws = (int)((budget - fixed) / (long long)p.w_stage_bytes);
        if (nws >= 3) break;
    }
    if (nws < 2) return false;
    p.nws = std::min(nws, 6);
    uint32_t cols = 32; while ((int)cols < 2 * C) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
    p.out_scale = out_scale; p.accumulate = accumulate;
    const size_t smem = (size_t)p.nas * p.a_stage_bytes + (size_t)p.nws * p.w_stage_bytes + p.xt_bytes + (size_t)(3 * p.nas + 2 * p.nws + 5) * 8 + 16;
    static bool attr = false;
    if (!attr) { BV2_CUDA(cudaFuncSetAttribute(k_tc_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
    launch_pdl(k_tc_pair, dim3(cdiv(p.T, p.TO), p.B), dim3(256), smem, st, p);
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// Persistent fused ResBlock pair (round-2 form of k_tc_pair; C <= 32): both weight tensors stay resident in shared
// memory, each CTA walks tiles g = blockIdx.x, +gridDim.x, ... and software-pipelines them across the roles:
//   warp 0      TMA producer: x tile i+2 while ...
//   warps 2-5   operand prologue (lrelu + RN-TF32) of x tile i+1
//   warp 1      MMA: conv1(tile i+1) into D1[(i+1)&1], then conv2(tile i) from XT[i&1] into D2[i&1]
//   warps 6-9   D2 init (bias2 + residual [+ y_old]) of tile i+1, D1 -> XT of tile i+1, tail (D2 -> HBM) of tile i
// D1, D2 and XT are double-buffered, so conv2 of a tile overlaps conv1 of the next and both overlap the epilogues.
struct TcPairPParams {
    const float* x; float* y; const float* w1; const float* w2; const float* b1; const float* b2;
    int C, T, B, K, dil, R1, RT, TO, nas, tiles_per_b, total_tiles;
    uint32_t a_stage_bytes, w_bytes, xt_bytes, tmem_cols, idesc;
    float out_scale; int accumulate;
};

__global__ void __launch_bounds__(320, 2) k_tc_pair_persist(TcPairPParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, NAS = p.nas, R = p.R1, RT = p.RT, ncg = C / 4;
    const int p2 = (p.K - 1) / 2, p1 = p2 * p.dil;
    uint8_t* sW1 = smem;
    uint8_t* sW2 = sW1 + p.w_bytes;
    uint8_t* sA = sW2 + p.w_bytes;
    uint8_t* sXT = sA + (size_t)NAS * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sXT + 2 * (size_t)p.xt_bytes);
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_AFULL = 0, B_AREADY = NAS, B_AEMPTY = 2 * NAS, B_W = 3 * NAS, B_D1FULL = B_W + 1, B_D1EMPTY = B_W + 3, B_XTFULL = B_W + 5,
              B_D2FULL = B_W + 7, NBARS = B_W + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NAS; i++) { mbar_init(BAR(B_AFULL + i), 1); mbar_init(BAR(B_AREADY + i), 128); mbar_init(BAR(B_AEMPTY + i), 1); }
        mbar_init(BAR(B_W), 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(BAR(B_D1FULL + i), 1); mbar_init(BAR(B_D1EMPTY + i), 128); mbar_init(BAR(B_XTFULL + i), 128); mbar_init(BAR(B_D2FULL + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const int ntl = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
    // tile i of this CTA -> (batch, first output row)
    auto tile_bt = [&](int i, int& b, int& t0) {
        const int g = (int)blockIdx.x + i * (int)gridDim.x;
        b = g / p.tiles_per_b;
        t0 = (g - b * p.tiles_per_b) * p.TO;
    };

    if (warp == 0) {
        // resident weights: independent of the upstream kernel, so they load before the PDL wait
        if (lane == 0) {
            mbar_expect_tx(BAR(B_W), 2 * p.w_bytes);
            bulk_g2s(smem_u32(sW1), p.w1, p.w_bytes, BAR(B_W));
            bulk_g2s(smem_u32(sW2), p.w2, p.w_bytes, BAR(B_W));
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        for (int i = 0; i < ntl; i++) {
            int b, t0; tile_bt(i, b, t0);
            const int tx0 = t0 - p2 - p1;
            const int r_lo = max(0, -tx0), r_hi = min(R, p.T - tx0);
            const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
            const int sa = i % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((i / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg);
            }
            __syncwarp();
            if (lane < ncg) {
                const float* src = p.x + (((size_t)b * ncg + lane) * p.T + (tx0 + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 1) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (lane == 0) {
            const uint32_t b_lbo = (uint32_t)C * 16u;
            const uint64_t b_kstep = (uint64_t)(2u * (uint32_t)C);
            const int nk = C / 8;
            const uint32_t tap_bytes = (uint32_t)C * (uint32_t)C * 4u;
            mbar_wait(BAR(B_W), 0);
            auto conv2 = [&](int i) {  // D2[i&1] += conv2(XT[i&1])
                const int buf = i & 1;
                mbar_wait(BAR(B_XTFULL + buf), (i >> 1) & 1);
                fence_after();
                const uint64_t xt0 = make_desc(smem_u32(sXT + (size_t)buf * p.xt_bytes), (uint32_t)RT * 16u, 128u);
                const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)RT);
                const uint32_t d2 = tmem + (uint32_t)(2 * C + buf * C);
                for (int j = 0; j < p.K; j++) {
                    uint64_t ad = xt0 + (uint64_t)(uint32_t)j;
                    uint64_t bd = make_desc(smem_u32(sW2) + (uint32_t)j * tap_bytes, b_lbo, 128u);
                    for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(d2, ad, bd, p.idesc, 1u);
                }
                umma_commit(BAR(B_D2FULL + buf));
            };
            for (int i = 0; i < ntl; i++) {
                const int sa = i % NAS, buf = i & 1;
                mbar_wait(BAR(B_AREADY + sa), (i / NAS) & 1);
                mbar_wait(BAR(B_D1EMPTY + buf), ((i >> 1) & 1) ^ 1);
                fence_after();
                const uint64_t a0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes), (uint32_t)R * 16u, 128u);
                const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R);
                const uint32_t d1 = tmem + (uint32_t)(buf * C);
                for (int j = 0; j < p.K; j++) {
                    uint64_t ad = a0 + (uint64_t)(uint32_t)(j * p.dil);
                    uint64_t bd = make_desc(smem_u32(sW1) + (uint32_t)j * tap_bytes, b_lbo, 128u);
                    for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(d1, ad, bd, p.idesc, (j | kk) ? 1u : 0u);
                }
                umma_commit(BAR(B_AEMPTY + sa));
                umma_commit(BAR(B_D1FULL + buf));
                if (i > 0) conv2(i - 1);
            }
            if (ntl > 0) conv2(ntl - 1);
        }
    } else if (warp < 6) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const int tid2 = threadIdx.x - 64;
        for (int i = 0; i < ntl; i++) {
            int b, t0; tile_bt(i, b, t0);
            const int tx0 = t0 - p2 - p1;
            const int r_lo = max(0, -tx0), r_hi = min(R, p.T - tx0);
            const int sa = i % NAS;
            mbar_wait(BAR(B_AFULL + sa), (i / NAS) & 1);
            xform_stage(reinterpret_cast<float4*>(sA + (size_t)sa * p.a_stage_bytes), ncg, R, r_lo, r_hi, 0.1f, tid2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_AREADY + sa));
        }
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const int q = warp & 3, m = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        // rows 128..RT-1 of both XT buffers feed only discarded output rows; zero them once so they stay finite
        if (m < p.K - 1)
            for (int buf = 0; buf < 2; buf++) {
                float4* XT = reinterpret_cast<float4*>(sXT + (size_t)buf * p.xt_bytes);
                for (int cg = 0; cg < ncg; cg++) XT[(size_t)cg * RT + 128 + m] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        auto tail = [&](int i) {  // D2[i&1] -> scale -> HBM
            int b, t0; tile_bt(i, b, t0);
            const int buf = i & 1, t_out = t0 + m;
            const bool ok_out = m < p.TO && t_out < p.T;
            float4* yb = reinterpret_cast<float4*>(p.y) + (size_t)b * ncg * p.T;
            mbar_wait(BAR(B_D2FULL + buf), (i >> 1) & 1);
            fence_after();
            for (int col = 0; col < C; col += 16) {
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)(2 * C + buf * C + col), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (!ok_out) continue;
#pragma unroll
                for (int g = 0; g < 4; g++)
                    yb[(size_t)((col >> 2) + g) * p.T + t_out] = make_float4(__uint_as_float(v[4 * g]) * p.out_scale, __uint_as_float(v[4 * g + 1]) * p.out_scale,
                                                                             __uint_as_float(v[4 * g + 2]) * p.out_scale, __uint_as_float(v[4 * g + 3]) * p.out_scale);
            }
            fence_before();
        };
        for (int i = 0; i < ntl; i++) {
            int b, t0; tile_bt(i, b, t0);
            const int buf = i & 1, t_out = t0 + m;
            const bool ok_out = m < p.TO && t_out < p.T;
            const float4* xb = reinterpret_cast<const float4*>(p.x) + (size_t)b * ncg * p.T;
            const float4* yb = reinterpret_cast<const float4*>(p.y) + (size_t)b * ncg * p.T;
            // ---- D2[buf] = bias2 + x (+ y_old): its previous user (tile i-2) was drained by tail(i-2) in program order
            for (int col = 0; col < C; col += 16) {
                uint32_t v2[16];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int cg = (col >> 2) + g;
                    float4 o = *reinterpret_cast<const float4*>(p.b2 + cg * 4);
                    if (ok_out) {
                        const float4 r = xb[(size_t)cg * p.T + t_out];
                        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                        if (p.accumulate) { const float4 a = yb[(size_t)cg * p.T + t_out]; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
                    }
                    v2[4 * g] = __float_as_uint(o.x); v2[4 * g + 1] = __float_as_uint(o.y); v2[4 * g + 2] = __float_as_uint(o.z); v2[4 * g + 3] = __float_as_uint(o.w);
                }
                tmem_st16(trow + (uint32_t)(2 * C + buf * C + col), v2);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            // ---- D1[buf] -> + bias1 -> lrelu -> TF32 -> XT[buf]   (XT[buf]'s previous reader, conv2(i-2), finished before tail(i-2) returned)
            float4* XT = reinterpret_cast<float4*>(sXT + (size_t)buf * p.xt_bytes);
            const int t_xt = t0 - p2 + m;
            const bool xt_in = t_xt >= 0 && t_xt < p.T;
            mbar_wait(BAR(B_D1FULL + buf), (i >> 1) & 1);
            fence_after();
            for (int col = 0; col < C; col += 16) {
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)(buf * C + col), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (xt_in) {
                        const float4 bb = *reinterpret_cast<const float4*>(p.b1 + ((col >> 2) + g) * 4);
                        o.x = to_tf32(lrelu(__uint_as_float(v[4 * g]) + bb.x, 0.1f)); o.y = to_tf32(lrelu(__uint_as_float(v[4 * g + 1]) + bb.y, 0.1f));
                        o.z = to_tf32(lrelu(__uint_as_float(v[4 * g + 2]) + bb.z, 0.1f)); o.w = to_tf32(lrelu(__uint_as_float(v[4 * g + 3]) + bb.w, 0.1f));
                    }
                    XT[(size_t)((col >> 2) + g) * RT + m] = o;
                }
            }
            fence_before();
            mbar_arrive(BAR(B_D1EMPTY + buf));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_XTFULL + buf));
            if (i > 0) tail(i - 1);
        }
        if (ntl > 0) tail(ntl - 1);
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

inline bool tc_pair_persist(const TcConvW& w1, const TcConvW& w2, const float* b1, const float* b2, const Act& x, const Act& y, int dil, float out_scale,
                            int accumulate, cudaStream_t st, int num_sms) {
    const int C = w1.Cin;
    if (!(w1.Cout == C && w2.Cin == C && w2.Cout == C && w1.K == w2.K && w1.KC == C && w2.KC == C && w1.nt == C && w2.nt == C && !w1.x3 && !w1.ups_u &&
          C <= 32 && C % 16 == 0 && (w1.K & 1)))
        return false;
    TcPairPParams p{};
    p.x = x.p; p.y = y.p; p.w1 = w1.w; p.w2 = w2.w; p.b1 = b1; p.b2 = b2;
    p.C = C; p.T = x.T; p.B = x.B; p.K = w1.K; p.dil = dil;
    p.R1 = 128 + (p.K - 1) * dil; p.RT = 128 + p.K - 1; p.TO = 128 - (p.K - 1);
    p.a_stage_bytes = (uint32_t)(C * p.R1 * 4); p.w_bytes = (uint32_t)(p.K * C * C * 4); p.xt_bytes = (uint32_t)(C * p.RT * 4);
    p.nas = 3;
    p.tiles_per_b = cdiv(p.T, p.TO); p.total_tiles = p.tiles_per_b * p.B;
    uint32_t cols = 32; while ((int)cols < 4 * C) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
    p.out_scale = out_scale; p.accumulate = accumulate;
    const size_t smem = 2 * (size_t)p.w_bytes + (size_t)p.nas * p.a_stage_bytes + 2 * (size_t)p.xt_bytes + (size_t)(3 * p.nas + 9) * 8 + 16;
    if (smem > 227 * 1024) return false;
    static bool attr = false;
    if (!attr) { BV2_CUDA(cudaFuncSetAttribute(k_tc_pair_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
    // resident CTAs per SM: shared memory (228 KB/SM, 1 KB reserved per CTA), registers (64K/SM), TMEM columns (512/SM)
    static cudaFuncAttributes fa = [] { cudaFuncAttributes a{}; cudaFuncGetAttributes(&a, k_tc_pair_persist); return a; }();
    static const int min_occ = getenv("BV2_PPAIR_MINOCC") ? atoi(getenv("BV2_PPAIR_MINOCC")) : 2;
    int occ = (int)((228 * 1024) / (smem + 1024));
    occ = std::min(occ, 65536 / (320 * std::max(fa.numRegs, 1)));
    occ = std::min(occ, (int)(512 / p.tmem_cols));
    occ = std::min(occ, 4);
    if (occ < min_occ) return false;  // one CTA per SM cannot hide the per-tile latency chain: the two-launch path is faster
    const int grid = std::min(p.total_tiles, num_sms * occ);
    launch_pdl(k_tc_pair_persist, dim3(grid), dim3(320), smem, st, p);
    return true;
}

}  // namespace bv2
