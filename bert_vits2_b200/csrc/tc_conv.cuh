// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d on the c4 activation layout, fp32 accumulate in TMEM.
// Replaces the 90 dilated MRF convolutions of the HiFi-GAN Generator (reference modules.py:296-309 via
// models.py:546-552) = 96 % of the MACs of SynthesizerTrn.infer, the ups (models.py:543-545) and the flow convs.
//
// GEMM view of one CTA tile:  D[128 time steps, N = Cout] += sum_{tap j} sum_{ci}  A_j[t, ci] * W_j[ci, co]
//   A_j[t, ci] = act(x[ci][t0 + t + j*dil - pad])  -- a time-shifted view of ONE staged activation tile.
// A staged operand chunk is [KC/G channel groups][R = 128*MT + (K-1)*dil rows][16 bytes], i.e. the K-major / no-swizzle
// UMMA canonical layout with SBO = 128 B (8 rows x 16 B) and LBO = R*16 B, so tap j is just the smem-descriptor start
// address advanced by j*dil*16 bytes: the im2col matrix is never materialised and each activation byte is fetched from
// HBM/L2 once per conv instead of K times.
//
// Operand types (template parameter F16):
//   F16 = 0  TF32 operands (kind::tf32, K = 8 per MMA, G = 4 fp32 channels per 16-byte group): the c4 activation tile is the
//            operand image; the prologue rounds it to TF32 in place (integer round-to-nearest: ALU pipe, not the XU pipe).
//   F16 = 1  FP16 operands (kind::f16, K = 16 per MMA, G = 8 channels per group), fp32 activations in HBM as before: the
//            prologue converts the staged fp32 c4 tile into a [KC/8][R][8 halves] image next to it.  FP16 has the same
//            11-bit significand as TF32 (identical rounding error), twice the tensor-pipe rate, half the shared-memory
//            operand bytes per MMA and half the L2->SM weight traffic -- the three things the round-1 ncu captures showed
//            these kernels to be bound by.  Out-of-range values saturate (cvt.rn.satfinite) instead of becoming inf.
//
// Warp roles: warp 0 = TMA producer (cp.async.bulk, mbarrier complete_tx), warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, 4 warps = operand prologue (leaky-relu + operand conversion in smem, zero fill of the conv padding
// rows, fence.proxy.async), 4 warps (the same ones in the one-tile kernel) = TMEM epilogue (accumulator init with
// bias/residual via tcgen05.st, tail tcgen05.ld -> scale/mask -> coalesced 16-byte stores), +1 weight-producer warp.
#pragma once
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <cuda_fp16.h>
#include "common.cuh"

namespace bv2 {

struct TcConvW {
    float* w = nullptr;  // packed [Cout/nt N tiles][nchunks][K][KC/G][nt][G] (fp32 TF32-rounded, or halves): one contiguous smem image per stage
    int Cin = 0, Cout = 0, K = 0, KC = 0, nchunks = 0, nt = 0;
    int ups_u = 0, ups_cout = 0;  // polyphase ConvTranspose1d: Cout = ups_u * ups_cout columns (phase-major)
    int f16 = 0;                  // operand type of the packed image
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
    int cin_off = 0, cout_off = 0;  // channel windows inside x / y (multiples of 4; of 8 for 16-bit tensors)
    int dil = 1;
    int out_tf32 = 0;    // round the stored output to TF32 (RN): a TF32 consumer may then skip its operand prologue
    int skip_xform = 0;  // TF32: input already TF32-exact, no activation / mask / padding needed (K == 1)
    int in_f16 = 0;      // FP16: x is a 16-bit c8 tensor [B][C/8][T][8] (the operand image itself: no prologue; K == 1)
    int out_f16 = 0;     // store y as a 16-bit c8 tensor
    int gate = 0;        // WN gate fused into the tail (reference commons.py:98-105): columns (2c, 2c+1) hold the tanh / sigmoid pre-activations of
                         // channel c (weights interleaved at load time); y gets tanh(a) * sigmoid(b) as a 16-bit c8 tensor with Cout/2 channels
    long long* prof = nullptr;  // probes: per-CTA phase timestamps (k_tc_conv1d)
    const float* ln_gamma = nullptr; const float* ln_beta = nullptr;  // LayerNorm over the Cout channels of each time step fused into the
                                                                      // tail (one N tile = all channels; combine with res for norm(x + conv))
};

inline float tf32_rn_host(float x) {
    uint32_t u; std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & 0xffffe000u;  // round to nearest, ties away (== cvt.rna.tf32.f32)
    float r; std::memcpy(&r, &u, 4);
    return r;
}
inline uint16_t f16_rn_host(float x) {
    if (x > 65504.f) x = 65504.f;
    if (x < -65504.f) x = -65504.f;
    __half h = __float2half_rn(x);
    uint16_t u; std::memcpy(&u, &h, 2);
    return u;
}
inline float f16_round_host(float x) {
    uint16_t u = f16_rn_host(x); __half h; std::memcpy(&h, &u, 2);
    return __half2float(h);
}

// Tuning knobs are compiled out of the product build (-DBV2_TUNING enables the BV2_* environment variables for probes).
inline int tune_env(const char* name, int dflt) {
#ifdef BV2_TUNING
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// w: [Cout][Cin][K] fp32 (weight-norm already folded)
// nt = N tile (0: largest divisor of Cout that is a multiple of 16 and <= 256)
// kc = K chunk (channels per pipeline stage)
// fill = false: only the sizes / layout fields are computed and an all-zero image of the right size is handed to `up` (the engine's
// measuring pass and bv2_load_packed, where the image comes from a file)
inline TcConvW tc_pack_weights(std::function<float*(const std::vector<float>&)>& up, const std::vector<float>& w, int Cout, int Cin, int K, int nt = 0,
                               int f16 = 0, int kc = 0, bool fill = true) {
    TcConvW t; t.Cin = Cin; t.Cout = Cout; t.K = K; t.f16 = f16;
    if (!nt) { nt = std::min(Cout, 256); while (Cout % nt || nt % 16) nt -= 16; }
    kc = tune_env("BV2_TC_KC", kc);
    if (!kc) kc = 32;
    t.KC = Cin >= kc ? kc : Cin;
    const int G = f16 ? 8 : 4;
    if (Cin % t.KC != 0 || t.KC % (2 * G) != 0 || Cout % 16 != 0 || Cout < 16)
        throw Error(-2, "tc_conv: unsupported channel counts " + std::to_string(Cin) + "->" + std::to_string(Cout));
    if (nt < 16 || nt > 256 || nt % 16 || Cout % nt) throw Error(-2, "tc_conv: bad N tile");
    t.nt = nt;
    t.nchunks = Cin / t.KC;
    const int ncg = t.KC / G;
    const size_t total = (size_t)Cin * K * Cout;
    std::vector<float> p(f16 ? (total + 1) / 2 : total);
    uint16_t* ph = reinterpret_cast<uint16_t*>(p.data());
    for (int tile = 0; fill && tile < Cout / nt; tile++)
        for (int c = 0; c < t.nchunks; c++)
            for (int j = 0; j < K; j++)
                for (int g = 0; g < ncg; g++)
                    for (int n = 0; n < nt; n++)
                        for (int e = 0; e < G; e++) {
                            const int ci = c * t.KC + g * G + e;
                            const float v = w[((size_t)(tile * nt + n) * Cin + ci) * K + j];
                            const size_t stage = ((size_t)tile * t.nchunks + c) * K + j;
                            const size_t idx = ((stage * ncg + g) * nt + n) * G + e;
                            if (f16) ph[idx] = f16_rn_host(v); else p[idx] = tf32_rn_host(v);
                        }
    t.w = up(p);
    return t;
}

// ConvTranspose1d(Cin->Cout, K, stride u, padding (K-u)/2) as a polyphase stride-1 conv (reference models.py:543-545):
//   out[t*u + r][co] = sum_m sum_ci x[t + floor((r+p)/u) - m][ci] * w[ci][co][(r+p)%u + m*u]
// -> an ordinary conv over input-rate time with Kp taps (union of the per-phase offsets), N = u*Cout columns ordered
// (r, co), structural zeros where a phase does not use a tap.  wT: [Cin][Cout][K] (weight-norm folded).
inline TcConvW tc_pack_upsample(std::function<float*(const std::vector<float>&)>& up, const std::vector<float>& wT, int Cin, int Cout, int K, int u,
                                int kc = 0, int f16 = 0, bool fill = true, int nt_max = 256) {
    const int p = (K - u) / 2, taps = K / u;
    int omin = 1 << 30, omax = -(1 << 30);
    for (int r = 0; r < u; r++)
        for (int m = 0; m < taps; m++) { int o = (r + p) / u - m; omin = std::min(omin, o); omax = std::max(omax, o); }
    int half = std::max(-omin, omax);
    const int Kp = 2 * half + 1;  // symmetric so that pad = (Kp-1)/2
    std::vector<float> w((size_t)u * Cout * Cin * Kp, 0.f);  // [N = u*Cout][Cin][Kp]
    for (int r = 0; fill && r < u; r++)
        for (int m = 0; m < taps; m++) {
            const int o = (r + p) / u - m, j = (r + p) % u + m * u, tap = o + half;
            for (int co = 0; co < Cout; co++)
                for (int ci = 0; ci < Cin; ci++)
                    w[(((size_t)(r * Cout + co)) * Cin + ci) * Kp + tap] = wT[((size_t)ci * Cout + co) * K + j];
        }
    int nt = std::min(u * Cout, nt_max);
    TcConvW t = tc_pack_weights(up, w, u * Cout, Cin, Kp, nt, f16, kc, fill);
    t.ups_u = u; t.ups_cout = Cout;
    return t;
}

struct TcParams {
    const float* x; float* y; const float* w; const float* bias; const float* res; const float* bias_b; const int* lens;
    int Cin_total, cin_off, Cout_total, cout_off, res_C_total, res_c_off, bias_b_stride;
    int nt;           // columns per N tile
    int T, B, K, dil, pad, KC, nchunks, R, nws, nas, MT;
    uint32_t a_stage_bytes, a_op_off, w_stage_bytes, tmem_cols, idesc;
    float in_slope, out_scale;
    int accumulate, relu, res_mode, in_mask, out_mask, ups_u, ups_cout;
    int out_tf32, skip_xform, in_f16, out_f16, gate;
    const float* ln_gamma; const float* ln_beta;
    uint32_t res_soff;          // LayerNorm tail only: byte offset (in dynamic shared memory) of the TMA-staged residual tile [nt/4][128 rows][16 B];
                                // 0 = residual pre-loaded into the accumulator by the epilogue warps (acc_init_tile)
    // batched-GEMM extensions (TF32 attention GEMMs of the fp32/tf32 engines): grid z = b * zsplit + h
    int zsplit;                 // 0/1: z == batch
    int x_batch_z, y_batch_z;   // 1: tensor's batch index is z (else b)
    int x_c_zstride, y_c_zstride;  // channel offset added per h
    long long w_zstride;        // packed-weight offset per z (floats)
    int w_mode;                 // 1: B operand rows come from a c4 activation tensor (K == 1): w = tensor base
    int w_ld, w_rows, w_c_total, w_c_off, w_c_zstride;
    long long* prof;            // probes only: per-CTA globaltimer stamps [ctas][8] (k_tc_conv1d); nullptr in the engine
};

// Device-side error flags: a barrier timeout raises both and lets the kernel run to completion instead of trapping the context
// ("never abort across the ABI").  g_tc_err_dev lives in device memory (what waiting threads poll: an L2 hit, never PCIe --
// polling the host-mapped copy from every waiting thread slowed every kernel ~100x in the first round-2 run);
// g_tc_err_flag points to a pinned, host-mapped int the host reads without a CUDA call (set by tc_init_device()).
__device__ int g_tc_err_dev = 0;
__device__ int* g_tc_err_flag = nullptr;

namespace tc {
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                   "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Bounded wait: a protocol bug must never hang the GPU.  try_wait carries a suspend-time hint so a waiting thread sleeps in
// hardware instead of polling (with ~10 mostly-idle role warps per CTA, un-hinted polling consumed > 50 % of the SM issue
// slots).  On a timeout (~2 s) the device error flag is raised and the wait is abandoned: the kernel finishes with
// undefined data, the host sees the flag at its next read-back and reports BV2_ERR_INTERNAL; once the flag is up every
// later wait gives up after one poll round, so a broken launch drains in milliseconds.
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity) {
    long long t0 = clock64();
    for (uint32_t it = 0;; it++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (done) return true;
        if ((it & 0xf) == 0xf) {
            if (*reinterpret_cast<volatile int*>(&g_tc_err_dev)) return false;  // another wait already timed out: drain quickly
            if (clock64() - t0 > 4000000000ll) {
                *reinterpret_cast<volatile int*>(&g_tc_err_dev) = 1;
                int* f = g_tc_err_flag;
                if (f) *reinterpret_cast<volatile int*>(f) = 1;
                printf("bv2 tc_conv: mbarrier wait timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
                       threadIdx.x, bar, parity);
                return false;
            }
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    if (!done) mbar_wait_slow(bar, parity);
}
// Bounded wait for the MMA-issuer warps, written as ONE asm block (no C++ control flow on a per-thread result, no call).
// Why: ptxas keeps the issuer's descriptors / loop state in uniform registers and updates them with UIADD3 only while the
// surrounding code is provably warp-uniform and free of calls (uniform registers do not survive a call).  mbar_wait() above
// branches on a per-thread predicate into a noinline function with printf: every value that is live across it is demoted to
// vector registers and each UTCHMMA then needs an ELECT + 5x R2UR.BROADCAST chain that re-uses one uniform-register set --
// measured 180-570 cycles per MMA in k_g2_conv against the 64-cycle tensor time of an N = 128 MMA (profiles/r02_g2_issue.md).
// Same timeout protocol as mbar_wait_slow (error flags raised, wait abandoned), minus the printf.
__device__ __forceinline__ void mbar_wait_u(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n, e;\n\t.reg .u64 t0, t1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "mov.u64 t0, %%clock64;\n\t"
        "mov.u32 n, 0;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "add.u32 n, n, 1;\n\t"
        "and.b32 e, n, 15;\n\t"
        "setp.ne.u32 p, e, 0;\n\t"
        "@p bra WAIT_%=;\n\t"
        "ld.volatile.global.u32 e, [%3];\n\t"
        "setp.ne.u32 p, e, 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "mov.u64 t1, %%clock64;\n\t"
        "sub.u64 t1, t1, t0;\n\t"
        "setp.lt.u64 p, t1, 4000000000;\n\t"
        "@p bra WAIT_%=;\n\t"
        "st.volatile.global.u32 [%3], 1;\n\t"
        "ld.global.u64 t1, [%4];\n\t"
        "setp.ne.u64 p, t1, 0;\n\t"
        "@p st.volatile.global.u32 [t1], 1;\n"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(1000000u), "l"(&g_tc_err_dev), "l"(&g_tc_err_flag) : "memory");
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
// One lane of a fully converged warp.  The MMA issuer runs its loops with ALL 32 lanes (descriptors and loop state stay in uniform
// registers) and only the tcgen05.mma / tcgen05.commit are guarded by this: issuing from inside an `if (lane == 0)` region makes
// nvcc treat every operand as divergent and wrap each UTCHMMA in an ELECT + 5x R2UR.BROADCAST + BRA.U.ANY waterfall loop
// (measured in round 2: ~185 cycles per MMA regardless of its size, 3x the 64-cycle MMA at N = 128).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// tcgen05.mma / tcgen05.commit with the lane election inside the asm block (statement stays in warp-uniform control flow)
template <int F16>
__device__ __forceinline__ void umma_el(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (F16)
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_el(uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
template <int F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (F16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int F16>
__device__ __forceinline__ void umma_e(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    umma_el<F16>(tmem_d, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar) { umma_commit_el(bar); }
// nk accumulating k-steps on one accumulator, unrolled for the common counts: a rolled loop re-uses one uniform-register set for the
// descriptors and serialises every UTCHMMA behind the R2UR round trip of its predecessor (~190 cycles per MMA of any size, round 2)
template <int F16, int NK>
__device__ __forceinline__ void umma_ksteps_n(uint32_t d, uint64_t ad, uint64_t bd, uint64_t a_kstep, uint64_t b_kstep, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
    for (int kk = 0; kk < NK; kk++) umma_e<F16>(d, ad + kk * a_kstep, bd + kk * b_kstep, idesc, kk ? 1u : acc_first);
}
template <int F16>
__device__ __forceinline__ void umma_ksteps(uint32_t d, uint64_t ad, uint64_t bd, uint64_t a_kstep, uint64_t b_kstep, uint32_t idesc, int nk, uint32_t acc_first) {
    // at most three alternatives: with four or more the compiler builds a jump table (BRX on a vector register) and ptxas then treats
    // everything after it as divergent, which takes the descriptors out of the uniform datapath (see mbar_wait_u)
    if (nk == 2) umma_ksteps_n<F16, 2>(d, ad, bd, a_kstep, b_kstep, idesc, acc_first);
    else if (nk == 4) umma_ksteps_n<F16, 4>(d, ad, bd, a_kstep, b_kstep, idesc, acc_first);
    else
        for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_e<F16>(d, ad, bd, idesc, kk ? 1u : acc_first);
}
// round-to-nearest (ties away) TF32 with two integer instructions: identical bits to cvt.rna.tf32.f32 for finite inputs,
// but issued on the ALU pipe (round 1 ncu: the cvt saturated the XU pipe at 94-98 % in the operand prologue)
__device__ __forceinline__ float to_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// two fp32 -> packed f16x2 (lo in bits [0,16)), round-to-nearest-even, saturating
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__host__ __device__ __forceinline__ uint32_t make_idesc(int f16, int n, int m = 128) {
    // c_format = F32 [4,6); a_format/b_format [7,10)/[10,13): 0 = F16, 2 = TF32; N>>3 [17,23); M>>4 [24,29)
    const uint32_t fmt = f16 ? 0u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Pre-load one 128-row x nt accumulator tile: bias (+ per-batch bias) (+/- residual) (+ previous output).
// All global loads of a 32-column batch are issued before the first tcgen05.st (memory-level parallelism: the
// epilogue warps are the only threads touching residual/output tensors).
// GEN = 1: generic epilogue (polyphase ConvTranspose scatter, per-batch bias, relu, 16-bit output); GEN = 0: plain conv
// epilogue.  The plain instantiation is 10-20 % faster on the MRF convs (measured): these kernels run at their register caps.
template <int NG, int GEN>
__device__ __forceinline__ void acc_init_tile(const TcParams& p, uint32_t trow, int b, int t, int n0, int nt, int yb = -1, int coff = -1, bool skip_res = false) {
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
        if (ok && p.res_mode && !(GEN && skip_res)) {
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
    tmem_wait_st();
}

// Drain one accumulator tile: TMEM -> [relu] -> scale/mask -> c4 global (16-byte stores, coalesced across a warp).
template <int NG, int GEN>
__device__ __forceinline__ void acc_tail_tile(const TcParams& p, uint32_t trow, int b, int t, int n0, int nt, int len, int yb = -1, int coff = -1) {
    const bool ok = t < p.T;
    const size_t tstride = (size_t)p.T * ((GEN && p.ups_u) ? p.ups_u : 1);
    if (yb < 0) yb = b;
    if (coff < 0) coff = p.cout_off;
    float4* ybp = reinterpret_cast<float4*>(p.y) + (size_t)yb * (p.Cout_total / 4) * tstride;
    uint4* yhp = reinterpret_cast<uint4*>(p.y) + (size_t)yb * (p.Cout_total / 8) * tstride;  // 16-bit c8 view
    const float s = ((p.out_mask && t >= len) ? 0.f : p.out_scale) * (p.res_mode == 2 ? -1.f : 1.f);
    for (int col0 = 0; col0 < nt; col0 += 4 * NG) {
        uint32_t v[NG / 4][16];
#pragma unroll
        for (int h = 0; h < NG / 4; h++) if (col0 + 16 * h < nt) tmem_ld16(trow + (uint32_t)(col0 + 16 * h), v[h]);
        tmem_wait_ld();
        if (!ok) continue;
#pragma unroll
        for (int h = 0; h < NG / 4; h++) {
            if (GEN && p.gate) {
                if (col0 + 16 * h < nt) {
                    float a[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const float ta = __uint_as_float(v[h][2 * e]), sb = __uint_as_float(v[h][2 * e + 1]);
                        a[e] = tanhf(ta) * (1.f / (1.f + expf(-sb))) * s;
                    }
                    uint4 o;
                    o.x = pack_h2(a[0], a[1]); o.y = pack_h2(a[2], a[3]); o.z = pack_h2(a[4], a[5]); o.w = pack_h2(a[6], a[7]);
                    yhp[(size_t)((coff + (n0 + col0 + 16 * h) / 2) / 8) * tstride + t] = o;
                }
                continue;
            }
            if (GEN && p.out_f16) {
                if (col0 + 16 * h < nt) {
                    float f[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) { f[e] = __uint_as_float(v[h][e]); if (p.relu) f[e] = fmaxf(f[e], 0.f); f[e] *= s; }
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int co = n0 + col0 + 16 * h + 8 * q;
                        uint4 o;
                        o.x = pack_h2(f[8 * q], f[8 * q + 1]); o.y = pack_h2(f[8 * q + 2], f[8 * q + 3]);
                        o.z = pack_h2(f[8 * q + 4], f[8 * q + 5]); o.w = pack_h2(f[8 * q + 6], f[8 * q + 7]);
                        yhp[(size_t)((coff + co) / 8) * tstride + t] = o;
                    }
                }
                continue;
            }
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

// Tail with a fused LayerNorm over the nt = Cout channels of each row (reference attentions.py:21-24 after the residual add of
// attentions.py:114,118): the accumulator already holds bias + residual + conv (accumulator-init fusion), so the row statistics
// are three cheap passes over TMEM (16 TB/s) instead of a separate kernel with an HBM round trip.
__device__ __forceinline__ void acc_tail_ln(const TcParams& p, uint32_t trow, int b, int t, int nt, int len, const float4* rs = nullptr) {
    // rs: this thread's row of the TMA-staged residual tile ([nt/4][128 rows] float4, already offset by the row); nullptr = the residual
    // was pre-loaded into the accumulator.  x = acc + residual is re-formed in each pass (shared-memory reads are cheap, TMEM is not written).
    const bool ok = t < p.T;
    float4* ybp = reinterpret_cast<float4*>(p.y) + (size_t)b * (p.Cout_total / 4) * p.T;
    float s = 0.f;
    for (int c0 = 0; c0 < nt; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)c0, v);
        tmem_wait_ld();
        if (rs) {
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const float4 r = rs[(size_t)(c0 / 4 + g) * 128];
                s += (__uint_as_float(v[4 * g]) + r.x) + (__uint_as_float(v[4 * g + 1]) + r.y) + (__uint_as_float(v[4 * g + 2]) + r.z) + (__uint_as_float(v[4 * g + 3]) + r.w);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 32; e++) s += __uint_as_float(v[e]);
        }
    }
    const float mean = s / (float)nt;
    float q = 0.f;
    for (int c0 = 0; c0 < nt; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(trow + (uint32_t)c0, v);
        tmem_wait_ld();
        if (rs) {
#pragma unroll
            for (int g = 0; g < 8; g++) {
                const float4 r = rs[(size_t)(c0 / 4 + g) * 128];
                float d;
                d = __uint_as_float(v[4 * g]) + r.x - mean; q = fmaf(d, d, q);
                d = __uint_as_float(v[4 * g + 1]) + r.y - mean; q = fmaf(d, d, q);
                d = __uint_as_float(v[4 * g + 2]) + r.z - mean; q = fmaf(d, d, q);
                d = __uint_as_float(v[4 * g + 3]) + r.w - mean; q = fmaf(d, d, q);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 32; e++) { const float d = __uint_as_float(v[e]) - mean; q = fmaf(d, d, q); }
        }
    }
    const float rstd = rsqrtf(q / (float)nt + 1e-5f);
    const float m = (p.out_mask && t >= len) ? 0.f : 1.f;
    for (int c0 = 0; c0 < nt; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        tmem_wait_ld();
        if (!ok) continue;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + c0 + 4 * g)), be = __ldg(reinterpret_cast<const float4*>(p.ln_beta + c0 + 4 * g));
            float4 x = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            if (rs) { const float4 r = rs[(size_t)(c0 / 4 + g) * 128]; x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w; }
            float4 o;
            o.x = ((x.x - mean) * rstd * ga.x + be.x) * m; o.y = ((x.y - mean) * rstd * ga.y + be.y) * m;
            o.z = ((x.z - mean) * rstd * ga.z + be.z) * m; o.w = ((x.w - mean) * rstd * ga.w + be.w) * m;
            ybp[(size_t)((p.cout_off + c0) / 4 + g) * p.T + t] = o;
        }
    }
}

// TF32 operand prologue of one staged activation chunk [ncg][R][4] (generic proxy), in place: leaky-relu + RN-TF32, zero
// rows outside [r_lo, r_hi).  The stage is contiguous, so the loop runs over flat 16-byte elements with 4 independent
// load->store chains per thread (the un-unrolled per-row loop exposed the full LDS latency on every element).
__device__ __forceinline__ void xform_stage(float4* A, int ncg, int R, int r_lo, int r_hi, float slope, int tid2) {
    const int total = ncg * R;
    int r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) r[u] = (tid2 + 128 * u) % R;
    const int step = 512 % R;  // row advance per iteration (512 elements), R >= 128
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
            r[u] += step;
            while (r[u] >= R) r[u] -= R;
        }
    }
}

// FP16 operand prologue: staged fp32 chunk S [2*ncg8][R][4] -> operand image O [ncg8][R][8 halves] (leaky-relu, RN-even
// saturating conversion, zero rows outside [r_lo, r_hi)).  Each thread owns rows tid2, tid2+128, ... and walks the channel
// groups four at a time (8 independent 16-byte shared loads in flight); consecutive lanes touch consecutive 16-byte
// elements on both sides: conflict free.
__device__ __forceinline__ void xform16_stage(const float4* S, uint4* O, int ncg8, int R, int r_lo, int r_hi, float slope, int tid2) {
    for (int r = tid2; r < R; r += 128) {
        const bool in = r >= r_lo && r < r_hi;
        for (int g0 = 0; g0 < ncg8; g0 += 4) {
            float4 v[4][2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[u][1] = v[u][0];
                if (in && g0 + u < ncg8) { v[u][0] = S[(size_t)(2 * (g0 + u)) * R + r]; v[u][1] = S[(size_t)(2 * (g0 + u) + 1) * R + r]; }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (g0 + u < ncg8) {
                    uint4 o;
                    o.x = pack_h2(lrelu(v[u][0].x, slope), lrelu(v[u][0].y, slope)); o.y = pack_h2(lrelu(v[u][0].z, slope), lrelu(v[u][0].w, slope));
                    o.z = pack_h2(lrelu(v[u][1].x, slope), lrelu(v[u][1].y, slope)); o.w = pack_h2(lrelu(v[u][1].z, slope), lrelu(v[u][1].w, slope));
                    O[(size_t)(g0 + u) * R + r] = o;
                }
            }
        }
    }
}
}  // namespace tc

// ------------------------------------------------------------------------------------------------------------
// One 128*MT-row tile per CTA.  grid: (M blocks of MT*128 time steps, N tiles, B [* heads])
//
// Accumulator-init fusion: before the first MMA the epilogue warps pre-load  bias (+ per-batch bias) (+/- residual)
// (+ previous output when accumulating)  into the TMEM accumulator with tcgen05.st while the first TMA loads are in
// flight; every MMA then accumulates, and the tail is only  TMEM -> [relu] -> scale/mask -> store.
template <int GEN, int F16>
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
    // barrier map: a_full[NAS], a_ready[NAS], a_empty[NAS], w_full[nws], w_empty[nws], acc_full, acc_init, res_full
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_AFULL = 0, B_AREADY = NAS, B_AEMPTY = 2 * NAS, B_WFULL = 3 * NAS, B_WEMPTY = 3 * NAS + p.nws, B_ACC = 3 * NAS + 2 * p.nws,
              B_INIT = B_ACC + 1, B_RES = B_ACC + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_RES + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NAS; i++) { mbar_init(BAR(B_AFULL + i), 1); mbar_init(BAR(B_AREADY + i), 128); mbar_init(BAR(B_AEMPTY + i), 1); }
        for (int i = 0; i < p.nws; i++) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), 1); }
        mbar_init(BAR(B_ACC), 1);
        mbar_init(BAR(B_INIT), 128);
        mbar_init(BAR(B_RES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // shfl: a warp-uniform value for ptxas (uniform registers in the MMA issuer)
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation) overlapped the tail of the previous kernel in
    // the stream.  Roles that touch only static data do not wait for it: the weight producer fills its ring and, when the accumulator
    // init is bias-only, the epilogue warps pre-load the accumulators while the upstream kernel is still running (timelines in
    // profiles/r02g_flow_conv_timelines.log: the init cost 5-6 us AFTER the wait before).  Everybody else waits, then releases the
    // dependents (releasing them before the wait let the whole rest of the stream become resident at once: measured slower).
    auto gtimer = [] { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    long long* prof = p.prof ? p.prof + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 : nullptr;
    if (prof && threadIdx.x == 0) prof[0] = gtimer();
    // LayerNorm tail with a TMA-staged residual (p.res_soff): the residual tile is copied into shared memory by the activation producer
    // right after the PDL wait and added by the tail, so the accumulator init is bias-only (static data, runs ahead of the wait) and the
    // MMAs never wait for residual loads.  (Pre-loading the residual into the accumulator took 15 us of a 24 us conv_o + LN launch: 12
    // dependent rounds of 4 float4 loads per thread AFTER the wait, with the MMA issuer parked on the init barrier,
    // profiles/r02h_flow_conv_timelines_f16_ffn.log.)
    const bool res_smem = GEN && p.res_soff != 0;
    const bool bias_only = (!p.res_mode || res_smem) && !p.accumulate;
    const bool static_role = (warp == 6 && !p.w_mode) || (warp >= 2 && warp <= 5 && bias_only);
    if (!static_role) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (prof && threadIdx.x == 0) prof[1] = gtimer();

    const int R = p.R;
    const int G = F16 ? 8 : 4;                                  // channels per 16-byte operand group
    const int ncg_in = (F16 && !p.in_f16) ? p.KC / 4 : p.KC / G;  // 16-byte groups per chunk in the GLOBAL tensor
    const int gdiv = (F16 && p.in_f16) ? 8 : 4;                 // channels per 16-byte group in the global tensor
    // rows r of the staged tile map to t = t0 - pad + r; rows outside [0, T) are the conv's zero padding
    const int r_lo = max(0, p.pad - t0);
    const int r_hi = min(R, p.T - (t0 - p.pad));

    if (warp == 0) {
        // ===== activation producer: lane 0 owns the mbarrier protocol, lanes 0..ncg_in-1 each issue one TMA bulk copy (one
        // contiguous run per channel group) so a chunk's copies are issued in parallel instead of serially by one thread
        const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
        const uint4* xg = reinterpret_cast<const uint4*>(p.x);
        for (int c = 0; c < p.nchunks; c++) {
            const int sa = c % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((c / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg_in);
            }
            __syncwarp();
            if (lane < ncg_in) {
                const uint4* src = xg + ((size_t)xb * (p.Cin_total / gdiv) + cin_off / gdiv + (size_t)c * ncg_in + lane) * p.T + (t0 - p.pad + r_lo);
                const uint32_t dst = smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u;
                bulk_g2s(dst, src, row_bytes, BAR(B_AFULL + sa));
            }
            if (res_smem && c == min(NAS, p.nchunks) - 1) {
                // residual tile -> shared memory, one bulk copy per channel group (behind the first ring fill: the operand tiles gate the MMAs)
                const int nrows = min(128, p.T - t0);
                if (lane == 0) mbar_expect_tx(BAR(B_RES), (uint32_t)nrows * 16u * (uint32_t)(nt / 4));
                __syncwarp();
                const float4* rg = reinterpret_cast<const float4*>(p.res) + ((size_t)yb * (p.res_C_total / 4) + p.res_c_off / 4) * p.T + t0;
                for (int g = lane; g < nt / 4; g += 32)
                    bulk_g2s(smem_u32(smem + p.res_soff) + (uint32_t)g * 2048u, rg + (size_t)g * p.T, (uint32_t)nrows * 16u, BAR(B_RES));
            }
        }
    } else if (warp == 6) {
        if (p.w_mode) {
            // B operand = rows n0.. of a c4 activation tensor (attention keys): lanes issue one bulk copy per channel group
            const int ncg = p.KC / 4;
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
            // slot / parity / addresses carried incrementally (one stage per ~20 instructions: the producer has to out-run the MMA issuer)
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w + (size_t)z * p.w_zstride + (size_t)blockIdx.y * p.nchunks * p.K * (p.w_stage_bytes / 4));
            const uint32_t bar_wf = BAR(B_WFULL), bar_we = BAR(B_WEMPTY), dst0 = smem_u32(sW), nws_u = (uint32_t)p.nws;
            uint32_t sw = 0, ph = 1u, dst = dst0;
            const int nst = p.nchunks * p.K;
            for (int i = 0; i < nst; i++, src += p.w_stage_bytes) {
                // (no griddepcontrol here: once the ring is full a slot only frees after MMAs ran, i.e. after the waiting roles saw the
                //  upstream kernel complete; the first launch_dependents of ANY thread releases the CTA's dependents, so this role stays silent)
                mbar_wait_u(bar_we + 8u * sw, ph);
                mbar_expect_tx(bar_wf + 8u * sw, p.w_stage_bytes);
                bulk_g2s(dst, src, p.w_stage_bytes, bar_wf + 8u * sw);
                dst += p.w_stage_bytes;
                if (++sw == nws_u) { sw = 0; ph ^= 1u; dst = dst0; }
            }
        }
    } else if (warp == 1) {
        {  // all 32 lanes run the issue loop convergently; uniform-register issue path: mbar_wait_u + in-asm election (see mbar_wait_u)
            // ===== MMA issuer: per weight tile, MT x KC/(2G) tcgen05.mma (M=128, N=nt), always accumulating.
            // Descriptors are advanced with 64-bit adds on the (addr >> 4) field: this single thread is the issue
            // bottleneck for narrow N, so the loop body is kept to a handful of integer instructions.
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);  // two channel groups per MMA
            const int nk = p.KC / (2 * G);
            mbar_wait_u(BAR(B_INIT), 0);
            fence_after();
            // ring slot / parity / descriptor state carried incrementally: no runtime division or multiply per stage (see tc_gen.cuh)
            const uint64_t a_stage16 = (uint64_t)(p.a_stage_bytes >> 4), w_stage16 = (uint64_t)(p.w_stage_bytes >> 4);
            const uint64_t a_desc_base = make_desc(smem_u32(sA) + p.a_op_off, a_lbo, 128u), b_desc_base = make_desc(smem_u32(sW), b_lbo, 128u);
            const uint32_t bar_ar = BAR(B_AREADY), bar_ae = BAR(B_AEMPTY), bar_wf = BAR(B_WFULL), bar_we = BAR(B_WEMPTY);
            const uint32_t nas_u = (uint32_t)NAS, nws_u = (uint32_t)p.nws;
            uint32_t sa = 0, aph = 0, sw = 0, wph = 0;
            uint64_t a_cur = a_desc_base, b_cur = b_desc_base;
            for (int c = 0; c < p.nchunks; c++) {
                mbar_wait_u(bar_ar + 8u * sa, aph);
                fence_after();
                if (prof && c == 0 && lane == 0) prof[2] = gtimer();
                uint64_t a_tap = a_cur;
                for (int j = 0; j < p.K; j++, a_tap += (uint64_t)(uint32_t)p.dil) {
                    mbar_wait_u(bar_wf + 8u * sw, wph);
                    fence_after();
                    for (int mt = 0; mt < MT; mt++)
                        umma_ksteps<F16>(tmem + (uint32_t)(mt * nt), a_tap + (uint64_t)(uint32_t)(mt * 128), b_cur, a_kstep, b_kstep, p.idesc, nk, 1u);
                    umma_commit_e(bar_we + 8u * sw);
                    b_cur += w_stage16;
                    if (++sw == nws_u) { sw = 0; wph ^= 1u; b_cur = b_desc_base; }
                }
                umma_commit_e(bar_ae + 8u * sa);
                a_cur += a_stage16;
                if (++sa == nas_u) { sa = 0; aph ^= 1u; a_cur = a_desc_base; }
            }
            umma_commit_e(BAR(B_ACC));
            if (prof && lane == 0) prof[3] = gtimer();
        }
    } else {
        const int tid2 = threadIdx.x - 64;
        const int q = warp & 3;
        // ===== accumulator init (overlaps the first TMA loads)
        for (int mt = 0; mt < MT; mt++)
            acc_init_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * nt), b, t0 + mt * 128 + q * 32 + lane, n0, nt, yb, cout_off, res_smem);
        fence_before();
        mbar_arrive(BAR(B_INIT));
        if (prof && tid2 == 0) prof[4] = gtimer();
        if (bias_only) {  // the init above touched static data only; everything below reads / overwrites tensors of the upstream kernel
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        const int len = p.lens ? p.lens[b] : p.T;
        const int r_mask_hi = p.in_mask ? min(r_hi, len - (t0 - p.pad)) : r_hi;  // rows >= this read as zero (x * x_mask)
        // ===== operand prologue on the staged tile (generic proxy), then hand over to the async proxy
        const float slope = p.in_slope;
        for (int c = 0; c < p.nchunks; c++) {
            const int sa = c % NAS;
            mbar_wait(BAR(B_AFULL + sa), (c / NAS) & 1);
            if (prof && tid2 == 0 && c == 0) prof[7] = gtimer();
            uint8_t* st = sA + (size_t)sa * p.a_stage_bytes;
            if (F16) {
                if (!p.in_f16) xform16_stage(reinterpret_cast<const float4*>(st), reinterpret_cast<uint4*>(st + p.a_op_off), p.KC / 8, R, r_lo, r_mask_hi, slope, tid2);
                else if (r_lo > 0 || r_hi < R) {
                    // 16-bit operand image with taps: the conv's zero padding is not part of the tensor; clear those rows of the staged
                    // tile (the TMA copies covered rows [r_lo, r_hi) only); rows t >= len are zero in the tensor itself (producer's mask)
                    uint4* A = reinterpret_cast<uint4*>(st);
                    const int ncg = p.KC / 8, nz = r_lo + (R - r_hi);
                    for (int i = tid2; i < ncg * nz; i += 128) {
                        const int g = i / nz, k = i - g * nz;
                        A[(size_t)g * R + (k < r_lo ? k : r_hi + (k - r_lo))] = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
            } else if (!p.skip_xform) {
                xform_stage(reinterpret_cast<float4*>(st), p.KC / 4, R, r_lo, r_mask_hi, slope, tid2);
            }
            fence_async_smem();
            mbar_arrive(BAR(B_AREADY + sa));
        }
        // ===== tail
        mbar_wait(BAR(B_ACC), 0);
        fence_after();
        if (prof && tid2 == 0) prof[5] = gtimer();
        if (GEN && p.ln_gamma) {
            const float4* rs = nullptr;
            if (res_smem) { mbar_wait(BAR(B_RES), 0); rs = reinterpret_cast<const float4*>(smem + p.res_soff) + (q * 32 + lane); }
            // (a 255-register instantiation that kept the 192-column row in registers -- one tcgen05.ld round trip instead of 24 -- ran
            //  11.1 us instead of 13.2 us per launch back to back but 1 us per layer SLOWER inside the flow: profiles/r02k_ab_ln_regs_weight_ring.jsonl)
            acc_tail_ln(p, tmem + ((uint32_t)(q * 32) << 16), b, t0 + q * 32 + lane, nt, len, rs);
        } else {
            for (int mt = 0; mt < MT; mt++)
                acc_tail_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * nt), b, t0 + mt * 128 + q * 32 + lane, n0, nt, len, yb, cout_off);
        }
        if (prof && tid2 == 0) prof[6] = gtimer();
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------
// Persistent variant for narrow layers (Cin <= 32: Generator stages 3/4 = 36 of the 90 MRF convs, the last ups).
// Those launches have thousands of 128-row tiles with only K * Cin/(2G) tiny MMAs each, so the one-tile-per-CTA kernel
// is bound by its per-tile latency chain (launch, TMEM alloc, barrier init, TMA round trip, tail).  Here each CTA
//   * keeps ALL weight taps resident in shared memory (<= 45 KB, loaded once),
//   * walks tiles blockIdx.x, +gridDim.x, ... with a 3-deep TMA ring for the activation tiles (producer runs ahead),
//   * double-buffers the TMEM accumulator: the epilogue warps pre-load tile i+1's accumulator (bias/residual) and
//     drain tile i-1 while the MMA warp works on tile i.
// 320 threads: warp 0 producer, warp 1 MMA issuer, warps 2-5 operand prologue, warps 6-9 accumulator init + tail.
template <int GEN, int OCC, int F16>
__global__ void __launch_bounds__(320, OCC) k_tc_conv1d_persist(TcParams p, int mtiles, int ntiles_total) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = p.nt, NAS = p.nas, R = p.R;
    const int G = F16 ? 8 : 4, ncg = p.KC / G, ncg_in = p.KC / 4;
    uint8_t* sWt = smem;                                   // resident weights: [K][ncg][nt][G]
    const uint32_t w_bytes = (uint32_t)(p.K * p.KC * nt * (F16 ? 2 : 4));
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
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // shfl: a warp-uniform value for ptxas (uniform registers in the MMA issuer)
    const int n_mine = (ntiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA

    if (warp == 0) {
        if (lane == 0) {  // weights do not depend on the upstream kernel: load them before the PDL wait
            mbar_expect_tx(BAR(B_WFULL), w_bytes);
            bulk_g2s(smem_u32(sWt), p.w, w_bytes, BAR(B_WFULL));
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        for (int i = 0; i < n_mine; i++) {
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            const int r_lo = max(0, p.pad - t0), r_hi = min(R, p.T - (t0 - p.pad));
            const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
            const int sa = i % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((i / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg_in);
            }
            __syncwarp();
            if (lane < ncg_in) {
                const float* src = p.x + (((size_t)b * (p.Cin_total / 4) + p.cin_off / 4 + lane) * p.T + (t0 - p.pad + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 1) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        {  // all 32 lanes run the issue loop convergently; uniform-register issue path: mbar_wait_u + in-asm election (see mbar_wait_u)
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);
            const uint64_t b_tap = (uint64_t)((uint32_t)ncg * nt);  // next tap's weight tile, in 16-byte units
            const int nk = p.KC / (2 * G);
            const uint64_t b_desc0 = make_desc(smem_u32(sWt), b_lbo, 128u);
            mbar_wait_u(BAR(B_WFULL), 0);
            for (int i = 0; i < n_mine; i++) {
                const int sa = i % NAS, ab = i & 1;
                mbar_wait_u(BAR(B_INIT + ab), (i >> 1) & 1);
                mbar_wait_u(BAR(B_AREADY + sa), (i / NAS) & 1);
                fence_after();
                const uint64_t a_desc0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + p.a_op_off, a_lbo, 128u);
                const uint32_t d = tmem + (uint32_t)(ab * nt);
                uint64_t bd_tap = b_desc0;
                for (int j = 0; j < p.K; j++, bd_tap += b_tap) {
                    uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(j * p.dil), bd = bd_tap;
                    umma_ksteps<F16>(d, ad, bd, a_kstep, b_kstep, p.idesc, nk, 1u);
                }
                umma_commit_e(BAR(B_AEMPTY + sa));
                umma_commit_e(BAR(B_ACC + ab));
            }
        }
    } else if (warp < 6) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
            uint8_t* st = sA + (size_t)sa * p.a_stage_bytes;
            if (F16) xform16_stage(reinterpret_cast<const float4*>(st), reinterpret_cast<uint4*>(st + p.a_op_off), ncg, R, r_lo, r_mask_hi, slope, tid2);
            else xform_stage(reinterpret_cast<float4*>(st), ncg, R, r_lo, r_mask_hi, slope, tid2);
            fence_async_smem();
            mbar_arrive(BAR(B_AREADY + sa));
        }
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        // ===== accumulator init (tile i+1) and tail (tile i), double-buffered TMEM
        const int q = warp & 3;
        auto init_tile = [&](int i) {
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            acc_init_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * nt), b, t0 + q * 32 + lane, 0, nt);
            fence_before();
            mbar_arrive(BAR(B_INIT + (i & 1)));
        };
        if (n_mine > 0) init_tile(0);
        for (int i = 0; i < n_mine; i++) {
            if (i + 1 < n_mine) init_tile(i + 1);
            const int tile = blockIdx.x + i * gridDim.x, b = tile / mtiles, t0 = (tile - b * mtiles) * 128;
            const int len = p.lens ? p.lens[b] : p.T;
            mbar_wait(BAR(B_ACC + (i & 1)), (i >> 1) & 1);
            fence_after();
            acc_tail_tile<4, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((i & 1) * nt), b, t0 + q * 32 + lane, 0, nt, len);
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
// Each CTA walks tiles of MT*128 rows (n-tile fastest, so CTAs that share an activation tile run together and hit L2);
// the TMA producers run continuously over the flat (tile, chunk, tap) sequence, so the activation ring (NAS deep) and
// the weight ring (nws deep) stay full across tile boundaries; the TMEM accumulator is double-buffered so the
// epilogue warps pre-load tile i+1's accumulator and drain tile i-1 while the MMA warp is busy with tile i.
// MT = 2 (256 rows per tile): every weight stage feeds two 128-row MMAs, which halves the L2->SM weight traffic per
// output row -- with one 128-row tile per weight pass these layers stream their whole weight tensor (0.36-1.4 MB) per
// tile and are bound by the ~42 B/clk/SM L2 path, not by the tensor pipe (round-1 ncu: tensor 16-21 %, DRAM 14-21 %).
// 352 threads: warp 0 activation producer, warp 1 MMA issuer, warps 2-5 operand prologue, warps 6-9 accumulator init +
// tail, warp 10 weight producer.
template <int GEN, int F16>
__global__ void __launch_bounds__(352, 1) k_tc_conv1d_pstream(TcParams p, int mtiles, int ntiles, int tiles_total) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = p.nt, NAS = p.nas, NWS = p.nws, R = p.R, NCH = p.nchunks, MT = p.MT;
    const int G = F16 ? 8 : 4, ncg = p.KC / G, ncg_in = p.KC / 4;
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
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // shfl: a warp-uniform value for ptxas (uniform registers in the MMA issuer)
    const int n_mine = (tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // tile -> (batch, m tile, n tile); n tile fastest
    auto decode = [&](int i, int& b, int& t0, int& ntile) {
        const int tile = blockIdx.x + i * gridDim.x;
        ntile = tile % ntiles;
        const int mm = tile / ntiles;
        b = mm / mtiles;
        t0 = (mm - b * mtiles) * 128 * MT;
    };
    (void)ncg;

    if (warp == 0) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        // ===== activation producer: runs up to NAS (tile, chunk) steps ahead of the MMA warp; copies issued by ncg_in lanes
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
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg_in);
            }
            __syncwarp();
            if (lane < ncg_in) {
                const float* src = p.x + (((size_t)b * (p.Cin_total / 4) + p.cin_off / 4 + (size_t)c * ncg_in + lane) * p.T + (t0 - p.pad + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 10) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (lane == 0) {
            // ===== weight producer: independent thread, so weight tiles stream up to NWS taps ahead across chunk and tile
            // boundaries (a single producer would serialise on the activation ring and bubble at every chunk).  Weights do
            // not depend on the upstream kernel: no PDL wait on this role.
            const int steps = n_mine * NCH;
            int wi = 0;
            const size_t wstage_f = p.w_stage_bytes / 4;
            const uint32_t bar_wf = BAR(B_WFULL), bar_we = BAR(B_WEMPTY), dst0 = smem_u32(sW), nws_u = (uint32_t)NWS;
            uint32_t sw = 0, ph = 1u, dst = dst0;
            for (int s_ = 0; s_ < steps; s_++) {
                int b, t0, ntile;
                decode(s_ / NCH, b, t0, ntile);
                const int c = s_ % NCH;
                const float* wsrc = p.w + ((size_t)ntile * NCH + c) * p.K * wstage_f;
                for (int j = 0; j < p.K; j++, wi++) {
                    mbar_wait_u(bar_we + 8u * sw, ph);
                    mbar_expect_tx(bar_wf + 8u * sw, p.w_stage_bytes);
                    bulk_g2s(dst, wsrc + (size_t)j * wstage_f, p.w_stage_bytes, bar_wf + 8u * sw);
                    dst += p.w_stage_bytes;
                    if (++sw == nws_u) { sw = 0; ph ^= 1u; dst = dst0; }
                }
            }
        }
    } else if (warp == 1) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        {  // all 32 lanes run the issue loop convergently; uniform-register issue path: mbar_wait_u + in-asm election (see mbar_wait_u)
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);
            const int nk = p.KC / (2 * G);
            const uint64_t a_stage16 = (uint64_t)(p.a_stage_bytes >> 4), w_stage16 = (uint64_t)(p.w_stage_bytes >> 4);
            const uint64_t a_desc_base = make_desc(smem_u32(sA) + p.a_op_off, a_lbo, 128u), b_desc_base = make_desc(smem_u32(sW), b_lbo, 128u);
            const uint32_t bar_ar = BAR(B_AREADY), bar_ae = BAR(B_AEMPTY), bar_wf = BAR(B_WFULL), bar_we = BAR(B_WEMPTY);
            const uint32_t nas_u = (uint32_t)NAS, nws_u = (uint32_t)NWS;
            uint32_t sa = 0, aph = 0, sw = 0, wph = 0;
            uint64_t a_cur = a_desc_base, b_cur = b_desc_base;
            for (int i = 0; i < n_mine; i++) {
                const int ab = i & 1;
                mbar_wait_u(BAR(B_INIT + ab), (i >> 1) & 1);
                fence_after();
                const uint32_t d0 = tmem + (uint32_t)(ab * MT * nt);
                for (int c = 0; c < NCH; c++) {
                    mbar_wait_u(bar_ar + 8u * sa, aph);
                    fence_after();
                    uint64_t a_tap = a_cur;
                    for (int j = 0; j < p.K; j++, a_tap += (uint64_t)(uint32_t)p.dil) {
                        mbar_wait_u(bar_wf + 8u * sw, wph);
                        fence_after();
                        for (int mt = 0; mt < MT; mt++)
                            umma_ksteps<F16>(d0 + (uint32_t)(mt * nt), a_tap + (uint64_t)(uint32_t)(mt * 128), b_cur, a_kstep, b_kstep, p.idesc, nk, 1u);
                        umma_commit_e(bar_we + 8u * sw);
                        b_cur += w_stage16;
                        if (++sw == nws_u) { sw = 0; wph ^= 1u; b_cur = b_desc_base; }
                    }
                    umma_commit_e(bar_ae + 8u * sa);
                    a_cur += a_stage16;
                    if (++sa == nas_u) { sa = 0; aph ^= 1u; a_cur = a_desc_base; }
                }
                umma_commit_e(BAR(B_ACC + ab));
            }
        }
    } else if (warp < 6) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
            uint8_t* st = sA + (size_t)sa * p.a_stage_bytes;
            if (F16) xform16_stage(reinterpret_cast<const float4*>(st), reinterpret_cast<uint4*>(st + p.a_op_off), p.KC / 8, R, r_lo, r_mask_hi, slope, tid2);
            else xform_stage(reinterpret_cast<float4*>(st), p.KC / 4, R, r_lo, r_mask_hi, slope, tid2);
            fence_async_smem();
            mbar_arrive(BAR(B_AREADY + sa));
        }
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        // ===== accumulator init (tile i+1) and tail (tile i), double-buffered TMEM
        const int q = warp & 3;
        auto init_tile = [&](int i) {
            int b, t0, ntile;
            decode(i, b, t0, ntile);
            const int n0 = ntile * nt;
            for (int mt = 0; mt < MT; mt++)
                acc_init_tile<8, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(((i & 1) * MT + mt) * nt), b, t0 + mt * 128 + q * 32 + lane, n0, nt);
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
            for (int mt = 0; mt < MT; mt++)
                acc_tail_tile<8, GEN>(p, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(((i & 1) * MT + mt) * nt), b, t0 + mt * 128 + q * 32 + lane, n0, nt, len);
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
// Persistent fused ResBlock pair (reference modules.py:296-309):  y = conv2(lrelu(conv1(lrelu(x)))) + x  [(+ y_old) * scale]
// conv1: K taps, dilation d;  conv2: K taps, dilation 1;  C <= 32 channels in and out.  One tile = TO = 128-(K-1) output
// rows: conv1 accumulates 128 rows in TMEM (D1), the epilogue warps turn D1 into the A-operand image XT[C/G][128+K-1][16 B]
// in SHARED memory (zero outside the sequence = conv2's padding), conv2 runs straight from XT through tap-shifted
// descriptors into a second accumulator (D2) that was pre-loaded with bias2 + residual.  The intermediate never touches
// HBM: 5 activation round trips per pair become ~2.5 and two launches become one.  Both weight tensors stay resident in
// shared memory; each CTA walks tiles g = blockIdx.x, +gridDim.x, ... and software-pipelines them across the roles:
//   warp 0      TMA producer: x tile i+2 while ...
//   warps 2-5   operand prologue (lrelu + operand conversion) of x tile i+1
//   warp 1      MMA: conv1(tile i+1) into D1[(i+1)&1], then conv2(tile i) from XT[i&1] into D2[i&1]
//   warps 6-9   D2 init (bias2 + residual [+ y_old]) of tile i+1, D1 -> XT of tile i+1, tail (D2 -> HBM) of tile i
// D1, D2 and XT are double-buffered, so conv2 of a tile overlaps conv1 of the next and both overlap the epilogues.
struct TcPairPParams {
    const float* x; float* y; const float* w1; const float* w2; const float* b1; const float* b2;
    int C, T, B, K, dil, R1, RT, TO, nas, tiles_per_b, total_tiles;
    uint32_t a_stage_bytes, a_op_off, w_bytes, xt_bytes, tmem_cols, idesc;
    float out_scale; int accumulate;
};

template <int F16>
__global__ void __launch_bounds__(320, 2) k_tc_pair_persist(TcPairPParams p) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = p.C, NAS = p.nas, R = p.R1, RT = p.RT;
    const int G = F16 ? 8 : 4, ncg = C / G, ncg_in = C / 4;
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
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // shfl: a warp-uniform value for ptxas (uniform registers in the MMA issuer)
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
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg_in);
            }
            __syncwarp();
            if (lane < ncg_in) {
                const float* src = p.x + (((size_t)b * ncg_in + lane) * p.T + (tx0 + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 1) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        {  // all 32 lanes run the issue loop convergently; uniform-register issue path: mbar_wait_u + in-asm election (see mbar_wait_u)
            const uint32_t b_lbo = (uint32_t)C * 16u;
            const uint64_t b_kstep = (uint64_t)(2u * (uint32_t)C);
            const int nk = C / (2 * G);
            const uint32_t tap_bytes = (uint32_t)C * (uint32_t)C * (F16 ? 2u : 4u);
            mbar_wait_u(BAR(B_W), 0);
            auto conv2 = [&](int i) {  // D2[i&1] += conv2(XT[i&1])
                const int buf = i & 1;
                mbar_wait_u(BAR(B_XTFULL + buf), (i >> 1) & 1);
                fence_after();
                const uint64_t xt0 = make_desc(smem_u32(sXT + (size_t)buf * p.xt_bytes), (uint32_t)RT * 16u, 128u);
                const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)RT);
                const uint32_t d2 = tmem + (uint32_t)(2 * C + buf * C);
                for (int j = 0; j < p.K; j++) {
                    uint64_t ad = xt0 + (uint64_t)(uint32_t)j;
                    uint64_t bd = make_desc(smem_u32(sW2) + (uint32_t)j * tap_bytes, b_lbo, 128u);
                    umma_ksteps<F16>(d2, ad, bd, a_kstep, b_kstep, p.idesc, nk, 1u);
                }
                umma_commit_e(BAR(B_D2FULL + buf));
            };
            for (int i = 0; i < ntl; i++) {
                const int sa = i % NAS, buf = i & 1;
                mbar_wait_u(BAR(B_AREADY + sa), (i / NAS) & 1);
                mbar_wait_u(BAR(B_D1EMPTY + buf), ((i >> 1) & 1) ^ 1);
                fence_after();
                const uint64_t a0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + p.a_op_off, (uint32_t)R * 16u, 128u);
                const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R);
                const uint32_t d1 = tmem + (uint32_t)(buf * C);
                for (int j = 0; j < p.K; j++) {
                    uint64_t ad = a0 + (uint64_t)(uint32_t)(j * p.dil);
                    uint64_t bd = make_desc(smem_u32(sW1) + (uint32_t)j * tap_bytes, b_lbo, 128u);
                    umma_ksteps<F16>(d1, ad, bd, a_kstep, b_kstep, p.idesc, nk, j ? 1u : 0u);
                }
                umma_commit_e(BAR(B_AEMPTY + sa));
                umma_commit_e(BAR(B_D1FULL + buf));
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
            uint8_t* st = sA + (size_t)sa * p.a_stage_bytes;
            if (F16) xform16_stage(reinterpret_cast<const float4*>(st), reinterpret_cast<uint4*>(st + p.a_op_off), ncg, R, r_lo, r_hi, 0.1f, tid2);
            else xform_stage(reinterpret_cast<float4*>(st), ncg, R, r_lo, r_hi, 0.1f, tid2);
            fence_async_smem();
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
                uint4* XT = reinterpret_cast<uint4*>(sXT + (size_t)buf * p.xt_bytes);
                for (int cg = 0; cg < ncg; cg++) XT[(size_t)cg * RT + 128 + m] = make_uint4(0u, 0u, 0u, 0u);
            }
        auto tail = [&](int i) {  // D2[i&1] -> scale -> HBM
            int b, t0; tile_bt(i, b, t0);
            const int buf = i & 1, t_out = t0 + m;
            const bool ok_out = m < p.TO && t_out < p.T;
            float4* yb = reinterpret_cast<float4*>(p.y) + (size_t)b * ncg_in * p.T;
            mbar_wait(BAR(B_D2FULL + buf), (i >> 1) & 1);
            fence_after();
            for (int col = 0; col < C; col += 16) {
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)(2 * C + buf * C + col), v);
                tmem_wait_ld();
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
            const float4* xb = reinterpret_cast<const float4*>(p.x) + (size_t)b * ncg_in * p.T;
            const float4* yb = reinterpret_cast<const float4*>(p.y) + (size_t)b * ncg_in * p.T;
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
            tmem_wait_st();
            // ---- D1[buf] -> + bias1 -> lrelu -> operand type -> XT[buf]   (XT[buf]'s previous reader, conv2(i-2), finished before tail(i-2) returned)
            const int t_xt = t0 - p2 + m;
            const bool xt_in = t_xt >= 0 && t_xt < p.T;
            mbar_wait(BAR(B_D1FULL + buf), (i >> 1) & 1);
            fence_after();
            for (int col = 0; col < C; col += 16) {
                uint32_t v[16];
                tmem_ld16(trow + (uint32_t)(buf * C + col), v);
                tmem_wait_ld();
                float f[16];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const float4 bb = *reinterpret_cast<const float4*>(p.b1 + col + 4 * g);
                    f[4 * g] = xt_in ? lrelu(__uint_as_float(v[4 * g]) + bb.x, 0.1f) : 0.f;
                    f[4 * g + 1] = xt_in ? lrelu(__uint_as_float(v[4 * g + 1]) + bb.y, 0.1f) : 0.f;
                    f[4 * g + 2] = xt_in ? lrelu(__uint_as_float(v[4 * g + 2]) + bb.z, 0.1f) : 0.f;
                    f[4 * g + 3] = xt_in ? lrelu(__uint_as_float(v[4 * g + 3]) + bb.w, 0.1f) : 0.f;
                }
                if (F16) {
                    uint4* XT = reinterpret_cast<uint4*>(sXT + (size_t)buf * p.xt_bytes);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        uint4 o;
                        o.x = pack_h2(f[8 * h], f[8 * h + 1]); o.y = pack_h2(f[8 * h + 2], f[8 * h + 3]);
                        o.z = pack_h2(f[8 * h + 4], f[8 * h + 5]); o.w = pack_h2(f[8 * h + 6], f[8 * h + 7]);
                        XT[(size_t)((col >> 3) + h) * RT + m] = o;
                    }
                } else {
                    float4* XT = reinterpret_cast<float4*>(sXT + (size_t)buf * p.xt_bytes);
#pragma unroll
                    for (int g = 0; g < 4; g++)
                        XT[(size_t)((col >> 2) + g) * RT + m] = make_float4(to_tf32(f[4 * g]), to_tf32(f[4 * g + 1]), to_tf32(f[4 * g + 2]), to_tf32(f[4 * g + 3]));
                }
            }
            fence_before();
            mbar_arrive(BAR(B_D1EMPTY + buf));
            fence_async_smem();
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

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
// The > 48 KB dynamic shared memory opt-in is a per-device (per-context) function attribute: call once per device before
// the first launch there (bv2_engine::finalize does; probes call it themselves).
inline void tc_clear_error() {
    const int z = 0;
    BV2_CUDA(cudaMemcpyToSymbol(g_tc_err_dev, &z, sizeof(z)));
}
// Returns the host view of this device's error flag (pinned, host-mapped; raised by a barrier timeout in any tcgen05 kernel).
inline int* tc_init_device() {
    const int mx = 227 * 1024;
#define BV2_SMEM_ATTR(k) BV2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))
    BV2_SMEM_ATTR((k_tc_conv1d<0, 0>)); BV2_SMEM_ATTR((k_tc_conv1d<1, 0>)); BV2_SMEM_ATTR((k_tc_conv1d<0, 1>)); BV2_SMEM_ATTR((k_tc_conv1d<1, 1>));
    BV2_SMEM_ATTR((k_tc_conv1d_persist<0, 2, 0>)); BV2_SMEM_ATTR((k_tc_conv1d_persist<1, 2, 0>));
    BV2_SMEM_ATTR((k_tc_conv1d_persist<0, 2, 1>)); BV2_SMEM_ATTR((k_tc_conv1d_persist<1, 2, 1>));
    BV2_SMEM_ATTR((k_tc_conv1d_pstream<0, 0>)); BV2_SMEM_ATTR((k_tc_conv1d_pstream<1, 0>));
    BV2_SMEM_ATTR((k_tc_conv1d_pstream<0, 1>)); BV2_SMEM_ATTR((k_tc_conv1d_pstream<1, 1>));
    BV2_SMEM_ATTR((k_tc_pair_persist<0>)); BV2_SMEM_ATTR((k_tc_pair_persist<1>));
#undef BV2_SMEM_ATTR
    int* h = nullptr;
    BV2_CUDA(cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped));
    *h = 0;
    int* d = nullptr;
    BV2_CUDA(cudaHostGetDevicePointer(&d, h, 0));
    BV2_CUDA(cudaMemcpyToSymbol(g_tc_err_flag, &d, sizeof(d)));
    tc_clear_error();
    return h;
}

// x: c4 input [B][x.C/4][T][4]; y: c4 output ([B][y.C/4][T*max(1,ups_u)][4]).  Channel windows via e.cin_off/e.cout_off.
// With e.in_f16 / e.out_f16 the tensor is the 16-bit c8 form [B][C/8][T][8] (Act.p reinterpreted).
inline void tc_conv1d(const TcConvW& w, const float* bias, const Act& x, const Act& y, const TcEpi& e, cudaStream_t st, int num_sms) {
    const int u = w.ups_u ? w.ups_u : 1;
    const int F16 = w.f16;
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
    const int halo = (w.K - 1) * e.dil;
    p.in_slope = e.in_slope; p.out_scale = e.out_scale; p.accumulate = e.accumulate; p.relu = e.relu; p.res_mode = e.res ? (e.res_mode ? e.res_mode : 1) : 0;
    p.in_mask = e.in_mask; p.out_mask = e.out_mask; p.ups_u = w.ups_u; p.ups_cout = w.ups_cout;
    p.out_tf32 = e.out_tf32; p.skip_xform = e.skip_xform; p.in_f16 = e.in_f16; p.out_f16 = e.out_f16;
    p.ln_gamma = e.ln_gamma; p.ln_beta = e.ln_beta; p.gate = e.gate; p.prof = e.prof;
    if (e.gate) BV2_CHECK(F16 && !w.ups_u && !e.res && !e.accumulate && !e.relu && !e.out_f16 && !e.ln_gamma && e.cout_off % 8 == 0 && y.C % 8 == 0 && 2 * y.C >= w.Cout, "gate epilogue");
    if (e.ln_gamma) BV2_CHECK(e.ln_beta && ntiles == 1 && nt % 32 == 0 && !w.ups_u && !e.out_f16 && !e.out_tf32 && !e.relu && e.out_scale == 1.f && e.res_mode != 2 && e.cout_off % 4 == 0, "LayerNorm tail needs one N tile holding every channel");
    if (e.skip_xform) BV2_CHECK(!F16 && w.K == 1 && e.in_slope == 1.f && !e.in_mask, "skip_xform needs a TF32 plain 1x1 conv input");
    if (e.in_f16) BV2_CHECK(F16 && e.in_slope == 1.f && !e.in_mask && e.cin_off % 8 == 0 && x.C % 8 == 0, "in_f16: the 16-bit tensor is the operand image (activation / mask applied by its producer)");
    if (e.out_f16) BV2_CHECK(F16 && !w.ups_u && e.cout_off % 8 == 0 && y.C % 8 == 0 && !e.res && !e.accumulate, "out_f16 epilogue");
    if (p.in_mask || p.out_mask) BV2_CHECK(e.lens != nullptr, "mask needs lens");
    BV2_CHECK(!(p.relu && (p.res_mode || p.accumulate)), "relu cannot be combined with residual/accumulate (accumulator-init fusion)");
    p.idesc = tc::make_idesc(F16, nt);
    const bool generic = w.ups_u || e.bias_b || e.relu || e.out_f16 || e.ln_gamma || e.gate;
    const uint32_t esz = F16 ? 2u : 4u;
    p.w_stage_bytes = (uint32_t)(p.KC * nt) * esz;
    auto set_rows = [&](int MT) {
        p.MT = MT; p.R = MT * 128 + halo;
        if (F16 && !e.in_f16) { p.a_op_off = (uint32_t)(p.KC * p.R * 4); p.a_stage_bytes = (uint32_t)(p.KC * p.R * 6); }
        else { p.a_op_off = 0; p.a_stage_bytes = (uint32_t)(p.KC * p.R) * esz; }
    };
    set_rows(1);
    const long long nctas = (long long)cdiv(p.T, 128) * ntiles * p.B;

    // ---- narrow layer with many tiles: persistent CTAs, resident weights, double-buffered TMEM
    const size_t w_all = (size_t)p.K * p.KC * nt * esz;
    if (tune_env("BV2_TC_PERSIST", 1) && !e.skip_xform && !e.in_f16 && !e.out_f16 && !e.ln_gamma && !e.gate && p.nchunks == 1 && ntiles == 1 && w_all <= 64 * 1024 && nctas >= 2 * num_sms) {
        p.nas = 3;
        const size_t wb = (w_all + 127) & ~(size_t)127;
        const size_t smem_p = wb + (size_t)p.nas * p.a_stage_bytes + (size_t)(3 * p.nas + 5) * 8 + 16;
        uint32_t pc = 32; while ((int)pc < 2 * nt) pc <<= 1;
        p.tmem_cols = pc;
        // measured (round 1): 2 CTAs/SM with ~100 registers (no spills) beat 3 with 68 for the plain epilogue
        const int per_sm = smem_p <= 110 * 1024 ? 2 : 1;
        const int mtiles = cdiv(p.T, 128);
        const int total = mtiles * p.B;
        const int grid_p = std::min(total, per_sm * num_sms);
        if (generic) launch_pdl(F16 ? k_tc_conv1d_persist<1, 2, 1> : k_tc_conv1d_persist<1, 2, 0>, dim3(grid_p), dim3(320), smem_p, st, p, mtiles, total);
        else launch_pdl(F16 ? k_tc_conv1d_persist<0, 2, 1> : k_tc_conv1d_persist<0, 2, 0>, dim3(grid_p), dim3(320), smem_p, st, p, mtiles, total);
        return;
    }
    // ---- wide layer with at least one tile per SM: persistent CTAs, continuously streamed weights, double-buffered TMEM
    if (tune_env("BV2_TC_PSTREAM", 1) && !e.skip_xform && !e.in_f16 && !e.ln_gamma && nctas >= num_sms && nt >= 64 && 2 * nt <= 512) {
        // MT = 2 (256-row tiles) when both accumulator pairs fit TMEM and there are enough 256-row tiles to fill the SMs
        int MT = (4 * nt <= 512 && (long long)cdiv(p.T, 256) * ntiles * p.B >= num_sms) ? 2 : 1;
        MT = tune_env("BV2_PSTREAM_MT", MT);
        if (4 * nt > 512) MT = 1;
        set_rows(MT);
        const uint32_t big = 220 * 1024;
        int nas2 = std::min(4, std::max(2, p.nchunks * 2));
        while (nas2 > 2 && (size_t)nas2 * p.a_stage_bytes + 3 * (size_t)p.w_stage_bytes + 2048 > big) nas2--;
        if ((size_t)nas2 * p.a_stage_bytes + 2 * (size_t)p.w_stage_bytes + 2048 > big && MT == 2) {  // does not fit: fall back to 128-row tiles
            MT = 1; set_rows(1);
            nas2 = std::min(4, std::max(2, p.nchunks * 2));
            while (nas2 > 2 && (size_t)nas2 * p.a_stage_bytes + 3 * (size_t)p.w_stage_bytes + 2048 > big) nas2--;
        }
        int nws2 = (int)((big - (size_t)nas2 * p.a_stage_bytes - 2048) / p.w_stage_bytes);
        nws2 = std::max(2, std::min(nws2, 8));
        p.nas = nas2; p.nws = nws2;
        uint32_t pc = 32; while ((int)pc < 2 * MT * nt) pc <<= 1;
        p.tmem_cols = pc;
        const size_t smem_s = (size_t)nas2 * p.a_stage_bytes + (size_t)nws2 * p.w_stage_bytes + (size_t)(3 * nas2 + 2 * nws2 + 4) * 8 + 16;
        BV2_CHECK(smem_s <= 227 * 1024, "tc_conv1d pstream shared memory");
        const int mtiles = cdiv(p.T, 128 * MT);
        const int total = mtiles * p.B * ntiles;
        const int grid_s = std::min(total, num_sms);
        if (generic) launch_pdl(F16 ? k_tc_conv1d_pstream<1, 1> : k_tc_conv1d_pstream<1, 0>, dim3(grid_s), dim3(352), smem_s, st, p, mtiles, ntiles, total);
        else launch_pdl(F16 ? k_tc_conv1d_pstream<0, 1> : k_tc_conv1d_pstream<0, 0>, dim3(grid_s), dim3(352), smem_s, st, p, mtiles, ntiles, total);
        return;
    }
    // ---- one tile per CTA.  Shared memory per CTA is capped (~48 KB) when there are more CTAs than SMs so that several
    // CTAs co-reside: one CTA's accumulator init / tail overlaps the other's MMA main loop
    uint32_t budget = (nctas > num_sms && nt <= 128) ? (uint32_t)tune_env("BV2_TC_SMEM_KB", 48) * 1024 : 200 * 1024;
    if (nt > 128 && 2 * nctas > num_sms && 2ull * p.a_stage_bytes + 2ull * p.w_stage_bytes + 2048 <= 112 * 1024)
        budget = 112 * 1024;  // wide layer launched on three streams at once (MRF resblock chains): let two CTAs share an SM
    // LayerNorm tail with a residual, at most one CTA per SM (small batches: the launch is a latency chain, not a throughput problem):
    // the residual tile is staged in shared memory by TMA (nt/4 channel groups x 128 rows x 16 B) instead of being pre-loaded into the
    // accumulator; the rings shrink to make room (the weight ring never needs more stages than the conv has)
    const bool res_smem = e.ln_gamma && p.res_mode == 1 && !p.accumulate && nctas <= num_sms && p.res_c_off % 4 == 0 && tune_env("BV2_LN_RES_SMEM", 1);
    const uint32_t res_bytes = res_smem ? (uint32_t)nt * 512u : 0u;
    if (res_smem) budget = 224 * 1024 - res_bytes;
    int nas = std::min(3, std::max(2, p.nchunks));
    while (nas > 2 && (size_t)nas * p.a_stage_bytes + 4 * (size_t)p.w_stage_bytes + 1024 > budget) nas--;
    p.nas = nas;
    int nws = ((int)budget - nas * (int)p.a_stage_bytes - 1024) / (int)p.w_stage_bytes;
    // (a ring as deep as the conv on small grids -- every weight stage of a 36-stage FFN conv_2 tile in flight ahead of the PDL wait -- measured
    //  no gain: 14.2 -> 14.1 us per launch, profiles/r02k_ab_ln_regs_weight_ring.jsonl)
    p.nws = std::max(2, std::min(nws, tune_env("BV2_TC_NWS_MAX", 8)));
    p.nws = std::max(2, std::min(p.nws, p.nchunks * p.K));
    uint32_t cols = 32; while ((int)cols < nt) cols <<= 1;
    p.tmem_cols = cols;
    size_t smem = (size_t)p.nas * p.a_stage_bytes + (size_t)p.nws * p.w_stage_bytes + (size_t)(3 * p.nas + 2 * p.nws + 3) * 8 + 16;
    if (res_smem) { smem = (smem + 15) & ~(size_t)15; p.res_soff = (uint32_t)smem; smem += res_bytes; }
    BV2_CHECK(smem <= 227 * 1024, "tc_conv1d shared memory");
    dim3 grid(cdiv(p.T, 128), ntiles, p.B);
    if (generic) launch_pdl(F16 ? k_tc_conv1d<1, 1> : k_tc_conv1d<1, 0>, grid, dim3(224), smem, st, p);
    else launch_pdl(F16 ? k_tc_conv1d<0, 1> : k_tc_conv1d<0, 0>, grid, dim3(224), smem, st, p);
}

// TF32 batched GEMMs on c4 operands (attention of the tf32 engine).
inline void tc_launch_simple(TcParams& p, int ntiles, int zdim, cudaStream_t st) {
    const int halo = (p.K - 1) * p.dil;
    p.MT = 1; p.R = 128 + halo; p.pad = (p.K - 1) / 2 * p.dil;
    p.a_stage_bytes = (uint32_t)(p.KC * p.R * 4); p.a_op_off = 0;
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
    p.idesc = tc::make_idesc(0, p.nt);
    const size_t smem = (size_t)p.nas * p.a_stage_bytes + (size_t)p.nws * p.w_stage_bytes + (size_t)(3 * p.nas + 2 * p.nws + 3) * 8 + 16;
    BV2_CHECK(smem <= 227 * 1024, "tc gemm shared memory");
    dim3 grid(cdiv(p.T, 128), ntiles, zdim);
    launch_pdl(k_tc_conv1d<0, 0>, grid, dim3(224), smem, st, p);
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

// Fused ResBlock pair launcher; returns false when the shapes do not fit (caller falls back to two launches).
inline bool tc_pair_persist(const TcConvW& w1, const TcConvW& w2, const float* b1, const float* b2, const Act& x, const Act& y, int dil, float out_scale,
                            int accumulate, cudaStream_t st, int num_sms) {
    const int C = w1.Cin, F16 = w1.f16;
    if (!(w1.Cout == C && w2.Cin == C && w2.Cout == C && w1.K == w2.K && w1.KC == C && w2.KC == C && w1.nt == C && w2.nt == C && w2.f16 == F16 && !w1.ups_u &&
          C <= 32 && C % 16 == 0 && (w1.K & 1)))
        return false;
    const uint32_t esz = F16 ? 2u : 4u;
    TcPairPParams p{};
    p.x = x.p; p.y = y.p; p.w1 = w1.w; p.w2 = w2.w; p.b1 = b1; p.b2 = b2;
    p.C = C; p.T = x.T; p.B = x.B; p.K = w1.K; p.dil = dil;
    p.R1 = 128 + (p.K - 1) * dil; p.RT = 128 + p.K - 1; p.TO = 128 - (p.K - 1);
    if (F16) { p.a_op_off = (uint32_t)(C * p.R1 * 4); p.a_stage_bytes = (uint32_t)(C * p.R1 * 6); }
    else { p.a_op_off = 0; p.a_stage_bytes = (uint32_t)(C * p.R1 * 4); }
    p.w_bytes = (uint32_t)(p.K * C * C) * esz; p.xt_bytes = (uint32_t)(C * p.RT) * esz;
    p.tiles_per_b = cdiv(p.T, p.TO); p.total_tiles = p.tiles_per_b * p.B;
    uint32_t cols = 32; while ((int)cols < 4 * C) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = tc::make_idesc(F16, C);
    p.out_scale = out_scale; p.accumulate = accumulate;
    static const int regs[2] = {[] { cudaFuncAttributes a{}; cudaFuncGetAttributes(&a, k_tc_pair_persist<0>); return a.numRegs; }(),
                                [] { cudaFuncAttributes a{}; cudaFuncGetAttributes(&a, k_tc_pair_persist<1>); return a.numRegs; }()};
    const int min_occ = tune_env("BV2_PPAIR_MINOCC", 2);
    int occ = 0;
    size_t smem = 0;
    for (int nas = 3; nas >= 2; nas--) {  // prefer the 3-deep activation ring, drop to 2 if that buys a second resident CTA
        p.nas = nas;
        smem = 2 * (size_t)p.w_bytes + (size_t)nas * p.a_stage_bytes + 2 * (size_t)p.xt_bytes + (size_t)(3 * nas + 9) * 8 + 16;
        if (smem > 227 * 1024) continue;
        // resident CTAs per SM: shared memory (228 KB/SM, 1 KB reserved per CTA), registers (64K/SM), TMEM columns (512/SM)
        occ = (int)((228 * 1024) / (smem + 1024));
        occ = std::min(occ, 65536 / (320 * std::max(regs[F16], 1)));
        occ = std::min(occ, (int)(512 / p.tmem_cols));
        occ = std::min(occ, 4);
        if (occ >= min_occ) break;
    }
    if (occ < min_occ) return false;  // one CTA per SM cannot hide the per-tile latency chain: the two-launch path is faster
    const int grid = std::min(p.total_tiles, num_sms * occ);
    launch_pdl(F16 ? k_tc_pair_persist<1> : k_tc_pair_persist<0>, dim3(grid), dim3(320), smem, st, p);
    return true;
}

}  // namespace bv2
