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

int main() {
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    int* flag = tc_init_device(); (void)flag;
    long long* d; cudaMalloc(&d, 148 * 2 * 8);
    const int nmma = 512;
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
