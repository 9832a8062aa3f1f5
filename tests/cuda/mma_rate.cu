// Microbenchmark: issue/execute rate of tcgen05.mma (kind::f16, M = 128, cta_group::1) as a function of N, of the shared-memory
// operand layout (K-major SWIZZLE_NONE "interleave" core matrices as used by tc_conv.cuh / tc_gen.cuh, vs K-major SWIZZLE_128B) and
// of how many accumulators the MMAs rotate over.  One CTA per SM, static operands (no TMA), one elected lane issues NMMA MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/mma_rate tests/cuda/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "../../bert_vits2_b200/csrc/tc_conv.cuh"

using namespace bv2;
using namespace bv2::tc;

// mode 0: no-swizzle K-major (LBO = rows*16 between 8-channel groups, SBO = 128); mode 1: SWIZZLE_128B K-major (SBO = 1024)
__global__ void __launch_bounds__(128, 1) k_rate(int N, int mode, int nacc, int nmma, int shift_rows, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // halves 1.0
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tslot;
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, N);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 24 * 1024);
        uint64_t ad0, bd0;
        uint32_t kstep_a, kstep_b;
        if (mode == 0) {
            const uint32_t R = 160;  // staged rows (128 + halo)
            ad0 = make_desc(a_base, R * 16u, 128u); bd0 = make_desc(b_base, (uint32_t)N * 16u, 128u);
            kstep_a = 2 * R; kstep_b = 2 * N;
        } else {
            // SWIZZLE_128B, K-major: rows of 128 B (64 halves), 8-row atoms of 1024 B; layout_type 2 at bits [61,64); LBO unused (1)
            auto d128 = [](uint32_t addr) {
                return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | ((uint64_t)((addr >> 7) & 7) << 49) | (2ull << 61);
            };
            ad0 = d128(a_base); bd0 = d128(b_base);
            kstep_a = 2; kstep_b = 2;  // 32 bytes per K = 16 step inside the 128-byte row
        }
        const uint64_t a_shift = mode == 0 ? (uint64_t)shift_rows : (uint64_t)(shift_rows * 8);  // rows -> 16-byte units
        long long t0 = clock64();
        for (int i = 0; i < nmma; i++) {
            const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
            const int kk = i & 3;
            uint64_t ad = ad0 + (uint64_t)(kk * kstep_a) + ((i >> 2) % 7) * a_shift;
            if (mode == 1 && shift_rows) ad = (ad & ~(7ull << 49)) | ((uint64_t)((((a_base >> 7) + ((i >> 2) % 7) * shift_rows)) & 7) << 49);
            const uint64_t bd = bd0 + (uint64_t)(kk * kstep_b);
            if (elect_one()) umma<1>(d, ad, bd, idesc, i >= nacc ? 1u : 0u);
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// issue-path variants on a lean loop (32-bit descriptor low words, 4 k-steps unrolled): 0 = inside an `if (lane == 0)` region,
// 1 = `if (elect_one())` per MMA, 2 = the election inside the asm block (convergent asm, operands can live in uniform registers)
__device__ __forceinline__ void umma_pred(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int V>
__global__ void __launch_bounds__(128, 1) k_issue(int N, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); fence_before(); __syncthreads(); fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, N), R = 160;
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t a0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (((R * 16u) >> 4) << 16), b0 = ((smem_u32(smem + 24 * 1024) & 0x3ffffu) >> 4) | ((((uint32_t)N * 16u) >> 4) << 16);
        long long t0 = clock64(), t1 = 0;
        if (V == 0) {
            if (lane == 0) {
                for (int i = 0; i < nmma; i += 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) umma<1>(tmem, ((uint64_t)hi << 32) | (a0 + kk * 2 * R), ((uint64_t)hi << 32) | (b0 + kk * 2 * N), idesc, 1u);
                }
                t1 = clock64();
                umma_commit(smem_u32(&bar));
            }
            __syncwarp();
        } else {
            for (int i = 0; i < nmma; i += 4) {
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    if (V == 1) { if (elect_one()) umma<1>(tmem, ((uint64_t)hi << 32) | (a0 + kk * 2 * R), ((uint64_t)hi << 32) | (b0 + kk * 2 * N), idesc, 1u); }
                    else umma_pred(tmem, ((uint64_t)hi << 32) | (a0 + kk * 2 * R), ((uint64_t)hi << 32) | (b0 + kk * 2 * N), idesc, 1u);
                }
            }
            t1 = clock64();
            if (elect_one()) umma_commit(smem_u32(&bar));
        }
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int V>
static void run_issue(long long* d, int nmma) {
    cudaFuncSetAttribute(k_issue<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int N : {16, 64, 128, 256}) {
        k_issue<V><<<148, 128, 64 * 1024>>>(N, nmma, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d N %d: CUDA error %s\n", V, N, cudaGetErrorString(e)); exit(1); }
        long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double a = 0, b = 0; for (int i = 0; i < 148; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
        printf("  issue variant %d (%s) N=%3d : issue %.1f  complete %.1f cycles per MMA (tensor floor %d)\n", V,
               V == 0 ? "if (lane == 0) region" : V == 1 ? "if (elect_one()) per MMA" : "election inside the asm", N, a / 148 / nmma, b / 148 / nmma, 128 * N / 256);
    }
}

// accumulator switching: MMAs rotate over `nacc` accumulators, switching every `period` MMAs (lean unrolled issue loop)
template <int PERIOD>
__global__ void __launch_bounds__(128, 1) k_accsw(int N, int nacc, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); fence_before(); __syncthreads(); fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, N), R = 160;
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
        const uint32_t a0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (((R * 16u) >> 4) << 16), b0 = ((smem_u32(smem + 24 * 1024) & 0x3ffffu) >> 4) | ((((uint32_t)N * 16u) >> 4) << 16);
        long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t d = tmem + (uint32_t)(acc * N);
                if (elect_one()) umma<1>(d, hi | (a0 + (u & 3) * 2 * R), hi | (b0 + (u & 3) * 2 * N), idesc, 1u);
                if (((u + 1) % PERIOD) == 0) acc = acc + 1 == nacc ? 0 : acc + 1;
            }
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int PERIOD>
static void run_accsw(long long* d, int nmma) {
    cudaFuncSetAttribute(k_accsw<PERIOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int N : {16, 128})
        for (int nacc : {1, 2, 4}) {
            if (nacc * N > 512) continue;
            k_accsw<PERIOD><<<148, 128, 64 * 1024>>>(N, nacc, nmma, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("accsw: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
            long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double a = 0, b = 0; for (int i = 0; i < 148; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
            printf("  accumulator switch every %d MMAs over %d accumulators, N=%3d : issue %.1f  complete %.1f cycles per MMA (tensor floor %d)\n", PERIOD, nacc, N, a / 148 / nmma,
                   b / 148 / nmma, 128 * N / 256);
        }
}

// streaming operands: every MMA reads a different 128-row window of a large staged tile (no operand re-use between consecutive MMAs)
// MODE 0: no-swizzle K-major [2 k-groups][R rows][16 B];  MODE 1: SWIZZLE_128B K-major rows of 128 B (uses k-slice u&3 of each row block)
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_stream(int N, int a_step_rows, int b_rot, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); fence_before(); __syncthreads(); fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, N);
        // A region: 128 KB at offset 0; B region: 64 KB at offset 128 KB
        const uint32_t RA = 4096;  // no-swizzle: 2 k-groups x 4096 rows x 16 B = 128 KB
        uint64_t a_hi, b_hi; uint32_t a0, b0, a_row_units, b_stage_units, ak, bk;
        if (MODE == 0) {
            a_hi = b_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
            a0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (((RA * 16u) >> 4) << 16);
            b0 = ((smem_u32(smem + 128 * 1024) & 0x3ffffu) >> 4) | ((((uint32_t)N * 16u) >> 4) << 16);
            a_row_units = 1; b_stage_units = (uint32_t)(2 * N); ak = 0; bk = 0;
        } else {
            a_hi = b_hi = ((uint64_t)((1024u >> 4) | (1u << 14)) << 32) | (2ull << 61);
            a0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (1u << 16);
            b0 = ((smem_u32(smem + 128 * 1024) & 0x3ffffu) >> 4) | (1u << 16);
            a_row_units = 8; b_stage_units = (uint32_t)(N * 8); ak = 2; bk = 2;  // 128-byte rows; k-slices of 32 B
        }
        const int a_windows = MODE == 0 ? (int)((RA - 128) / (a_step_rows ? a_step_rows : 1)) : (1024 - 128) / (a_step_rows ? a_step_rows : 1);
        long long t0 = clock64();
        int aw = 0, bw = 0;
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                uint32_t a_lo = a0 + (uint32_t)(aw * a_step_rows) * a_row_units + (MODE == 1 ? (u & 3) * ak : 0);
                uint32_t b_lo = b0 + (uint32_t)bw * b_stage_units + (MODE == 1 ? (u & 3) * bk : 0);
                uint64_t ad = a_hi | a_lo, bd = b_hi | b_lo;
                if (MODE == 1) {  // base_offset = (address >> 7) & 7 for windows that do not start on a 1024-byte atom
                    ad |= (uint64_t)(((a_lo & 0x3fffu) >> 3) & 7) << 49;
                }
                if (elect_one()) umma<1>(tmem, ad, bd, idesc, 1u);
                aw = aw + 1 >= a_windows ? 0 : aw + 1;
                if (b_rot) bw = bw + 1 >= b_rot ? 0 : bw + 1;
            }
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int MODE>
static void run_stream(long long* d, int nmma) {
    cudaFuncSetAttribute(k_stream<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int N : {16, 64, 128})
        for (int step : {0, 128, 1, 5})
            for (int brot : {0, 4}) {
                if (MODE == 1 && (size_t)brot * N * 128 > 64 * 1024) continue;
                if (MODE == 0 && (size_t)brot * N * 32 > 64 * 1024) continue;
                k_stream<MODE><<<148, 128, 200 * 1024>>>(N, step, brot, nmma, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("stream: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
                long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double a = 0, b = 0; for (int i = 0; i < 148; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
                printf("  streaming %s N=%3d A window step %3d rows, B rotating over %d stages : issue %.1f  complete %.1f cycles per MMA (tensor floor %d)\n",
                       MODE ? "SWIZZLE_128B" : "no-swizzle  ", N, step, brot, a / 148 / nmma, b / 148 / nmma, 128 * N / 256);
            }
}

// waiting warps: does mbarrier polling by the other warps of the CTA (epilogue / producer roles parked in mbar_wait) slow the issuer?
// WAITERS = number of extra warps parked in mbar_wait (all 32 lanes or lane 0 only) until the MMAs are done
__global__ void __launch_bounds__(512, 1) k_waiters(int N, int nwait, int lane0_only, int nmma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); fence_before(); __syncthreads(); fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, N), R = 160;
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
        const uint32_t a0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (((R * 16u) >> 4) << 16), b0 = ((smem_u32(smem + 24 * 1024) & 0x3ffffu) >> 4) | ((((uint32_t)N * 16u) >> 4) << 16);
        long long t0 = clock64();
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (elect_one()) umma<1>(tmem, hi | (a0 + (u & 3) * 2 * R), hi | (b0 + (u & 3) * 2 * N), idesc, 1u);
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; mbar_arrive(smem_u32(&bar2)); }
    } else if (warp >= 2 && warp < 2 + nwait) {
        if (!lane0_only || lane == 0) mbar_wait(smem_u32(&bar2), 0);
    }
    fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
static void run_waiters(long long* d, int nmma) {
    cudaFuncSetAttribute(k_waiters, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int N : {16, 128})
        for (int nwait : {0, 1, 3, 6, 14})
            for (int l0 : {0, 1}) {
                if (nwait == 0 && l0) continue;
                k_waiters<<<148, 512, 64 * 1024>>>(N, nwait, l0, nmma, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("waiters: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
                long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double a = 0, b = 0; for (int i = 0; i < 148; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
                printf("  %2d warps parked in mbar_wait (%s), N=%3d : issue %.1f  complete %.1f cycles per MMA (tensor floor %d)\n", nwait, l0 ? "lane 0 only" : "all 32 lanes", N,
                       a / 148 / nmma, b / 148 / nmma, 128 * N / 256);
            }
}


// Descriptor-update / tap-shift study on the uniform-register issue path (umma_el, no per-thread control flow): 8 unrolled MMAs per
// iteration; slot u uses A start = a_base + u * shift rows (the tap shift of a dilated conv) and B k-step u & 3.
// UPD = 0: descriptors are loop-invariant per slot; UPD = 1: both bases advance by a runtime step every iteration (UIADD3 updates of
// the uniform registers feeding UTCHMMA).  commit = 1: a tcgen05.commit + fence after every 4 MMAs (one weight stage of k_g2_conv).
template <int UPD>
__global__ void __launch_bounds__(128, 1) k_desc(int N, int nmma, int shift, int step, int commit, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); fence_before(); __syncthreads(); fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tslot, 0);
    if (warp == 1) {
        const uint32_t idesc = make_idesc(1, N), R = 320;
        const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
        const uint32_t a0 = ((smem_u32(smem) & 0x3ffffu) >> 4) | (((R * 16u) >> 4) << 16), b0 = ((smem_u32(smem + 40 * 1024) & 0x3ffffu) >> 4) | ((((uint32_t)N * 16u) >> 4) << 16);
        uint32_t aoff = 0, boff = 0;
        const uint32_t bar2a = smem_u32(&bar2);
        long long t0 = clock64();
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                umma_el<1>(tmem, hi | (a0 + aoff + (uint32_t)(u * shift)), hi | (b0 + boff + (uint32_t)((u & 3) * 2 * N)), idesc, 1u);
                if (commit && (u & 3) == 3) { umma_commit_el(bar2a); fence_after(); }
            }
            if (UPD) { aoff = (aoff + (uint32_t)step) & 63u; boff = (boff + (uint32_t)step) & 15u; }
        }
        long long t1 = clock64();
        umma_commit_el(smem_u32(&bar));
        mbar_wait_u(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int UPD>
static void run_desc(long long* d, int nmma) {
    cudaFuncSetAttribute(k_desc<UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int N : {16, 64, 128})
        for (int shift : {0, 1, 5, 8})
            for (int commit : {0, 1}) {
                k_desc<UPD><<<148, 128, 64 * 1024>>>(N, nmma, shift, 1, commit, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("desc: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
                long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double a = 0, b = 0; for (int i = 0; i < 148; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
                printf("  desc study: %s N=%3d tap-shift=%d rows commit/4=%d : issue %.1f  complete %.1f cycles per MMA\n", UPD ? "bases updated per iteration" : "loop-invariant descriptors ", N, shift, commit,
                       a / 148 / nmma, b / 148 / nmma);
            }
}

int main() {
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    int* flag = tc_init_device(); (void)flag;
    long long* d; cudaMalloc(&d, 148 * 2 * 8);
    const int nmma = 512;
    if (getenv("MMA_RATE_DESC")) { run_desc<0>(d, 2048); run_desc<1>(d, 2048); return 0; }
    run_issue<0>(d, 2048); run_issue<1>(d, 2048); run_issue<2>(d, 2048);
    if (getenv("MMA_RATE_WAITERS")) { run_waiters(d, 2048); return 0; }
    if (getenv("MMA_RATE_STREAM")) { run_stream<0>(d, 2048); run_stream<1>(d, 2048); return 0; }
    run_accsw<1>(d, 2048); run_accsw<2>(d, 2048); run_accsw<8>(d, 2048);
    if (getenv("MMA_RATE_ISSUE_ONLY")) return 0;
    printf("tcgen05.mma kind::f16 M=128, %d MMAs per CTA, 148 CTAs: cycles per MMA (issue loop | until commit completes)\n", nmma);
    for (int mode = 0; mode < 2; mode++)
        for (int N : {16, 32, 64, 128, 256})
            for (int nacc : {1, 2})
                for (int shift : {0, 1}) {
                    if (nacc * N > 512) continue;
                    k_rate<<<148, 128, 64 * 1024>>>(N, mode, nacc, nmma, shift, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d N %d: CUDA error %s\n", mode, N, cudaGetErrorString(e)); return 1; }
                    long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                    double a = 0, b = 0; for (int i = 0; i < 148; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
                    printf("  %s N=%3d accumulators=%d tap-shift=%d : issue %.1f  complete %.1f   (tensor floor %d)\n", mode ? "SWIZZLE_128B" : "no-swizzle  ", N, nacc, shift,
                           a / 148 / nmma, b / 148 / nmma, 128 * N / 256);
                }
    return 0;
}
