// EXPERIMENTAL -- round-2 candidate, NOT compiled into libbv2.so (only tests/cuda/splitk_probe.cu includes it) and not
// yet run on hardware.  Split-K form of the one-tile-per-CTA tcgen05 conv (k_tc_conv1d<0> in ../tc_conv.cuh) for the
// flow's short sequences (13 time tiles at F = 1573): instead of cutting N into 32..128-column tiles to fill the SMs
// (which makes every MMA re-fetch a 4 KB A slab for 8-32 tensor cycles), every CTA keeps the FULL N (<= 256 columns) and
// takes a slice of the reduction (input-channel chunks); partial tiles are reduced in HBM with red.global.add.v4.f32.
//   grid = (time tiles, nsplit, batch*heads).  Split 0 carries bias / residual in its accumulator (accumulator-init
//   fusion as in the base kernel); the output must hold zeros (or the term to accumulate onto, e.g. the relative-value
//   seed of P.V) before the launch -- the launcher memsets it unless `accumulate` is set.
// Intended users: FFN conv_2 (768->192, k5: 24 chunks -> 6 splits x N=192) and P.V (25 key chunks
// -> 4 splits x N=96).  Restrictions: plain epilogue (no relu / out_tf32 / polyphase), weights from the packed image.
#pragma once
#include "../tc_conv.cuh"

namespace bv2 {

__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(224, 4) k_tc_conv1d_splitk(TcParams p, int nsplit) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128, s = blockIdx.y, z = blockIdx.z;
    const int zs = p.zsplit > 1 ? p.zsplit : 1;
    const int b = z / zs, hz = z - b * zs;
    const int xb = p.x_batch_z ? z : b, yb = p.y_batch_z ? z : b;
    const int cin_off = p.cin_off + hz * p.x_c_zstride, cout_off = p.cout_off + hz * p.y_c_zstride;
    const int nt = p.nt;
    const int c_lo = (int)((long long)s * p.nchunks / nsplit), c_hi = (int)((long long)(s + 1) * p.nchunks / nsplit);
    uint8_t* sA = smem;
    const int NAS = p.nas;
    uint8_t* sW = smem + (size_t)NAS * p.a_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)p.nws * p.w_stage_bytes);
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
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const int R = p.R, ncg = p.KC / 4;
    const int len = p.lens ? p.lens[b] : p.T;
    const int r_lo = max(0, p.pad - t0);
    const int r_hi = min(R, p.T - (t0 - p.pad));
    const int r_mask_hi = p.in_mask ? min(r_hi, len - (t0 - p.pad)) : r_hi;

    if (warp == 0) {
        const uint32_t row_bytes = (uint32_t)(r_hi - r_lo) * 16u;
        for (int c = c_lo; c < c_hi; c++) {
            const int lc = c - c_lo, sa = lc % NAS;
            if (lane == 0) {
                mbar_wait(BAR(B_AEMPTY + sa), ((lc / NAS) & 1) ^ 1);
                mbar_expect_tx(BAR(B_AFULL + sa), row_bytes * ncg);
            }
            __syncwarp();
            if (lane < ncg) {
                const float* src = p.x + (((size_t)xb * (p.Cin_total / 4) + cin_off / 4 + (size_t)c * ncg + lane) * p.T + (t0 - p.pad + r_lo)) * 4;
                bulk_g2s(smem_u32(sA + (size_t)sa * p.a_stage_bytes) + ((uint32_t)lane * R + (uint32_t)r_lo) * 16u, src, row_bytes, BAR(B_AFULL + sa));
            }
        }
    } else if (warp == 6) {
        if (lane == 0) {
            int wi = 0;
            const float* wtile = p.w + (size_t)z * p.w_zstride;
            for (int c = c_lo; c < c_hi; c++)
                for (int j = 0; j < p.K; j++, wi++) {
                    const int sw = wi % p.nws;
                    mbar_wait(BAR(B_WEMPTY + sw), ((wi / p.nws) & 1) ^ 1);
                    mbar_expect_tx(BAR(B_WFULL + sw), p.w_stage_bytes);
                    bulk_g2s(smem_u32(sW + (size_t)sw * p.w_stage_bytes), wtile + ((size_t)c * p.K + j) * (p.w_stage_bytes / 4), p.w_stage_bytes, BAR(B_WFULL + sw));
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t a_lbo = (uint32_t)R * 16u, b_lbo = (uint32_t)nt * 16u;
            const uint64_t a_kstep = (uint64_t)(2u * (uint32_t)R), b_kstep = (uint64_t)(2u * (uint32_t)nt);
            const int nk = p.KC / 8;
            int wi = 0;
            mbar_wait(BAR(B_INIT), 0);
            fence_after();
            for (int c = c_lo; c < c_hi; c++) {
                const int lc = c - c_lo, sa = lc % NAS;
                mbar_wait(BAR(B_AREADY + sa), (lc / NAS) & 1);
                fence_after();
                const uint64_t a_desc0 = make_desc(smem_u32(sA + (size_t)sa * p.a_stage_bytes), a_lbo, 128u);
                for (int j = 0; j < p.K; j++, wi++) {
                    const int sw = wi % p.nws;
                    mbar_wait(BAR(B_WFULL + sw), (wi / p.nws) & 1);
                    fence_after();
                    uint64_t ad = a_desc0 + (uint64_t)(uint32_t)(j * p.dil);
                    uint64_t bd = make_desc(smem_u32(sW + (size_t)sw * p.w_stage_bytes), b_lbo, 128u);
                    // split 0 accumulates onto its pre-loaded bias/residual tile; the other splits start from the first product
                    for (int kk = 0; kk < nk; kk++, ad += a_kstep, bd += b_kstep) umma_tf32(tmem, ad, bd, p.idesc, (s == 0 || (wi | kk)) ? 1u : 0u);
                    umma_commit(BAR(B_WEMPTY + sw));
                }
                umma_commit(BAR(B_AEMPTY + sa));
            }
            umma_commit(BAR(B_ACC));
        }
    } else {
        const int tid2 = threadIdx.x - 64;
        const int q = warp & 3;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        const int t = t0 + q * 32 + lane;
        if (s == 0) acc_init_tile<4, 0>(p, trow, b, t, 0, nt, yb, cout_off);  // p.accumulate is 0 here (launcher): bias (+/- residual) only
        fence_before();
        mbar_arrive(BAR(B_INIT));
        for (int c = c_lo; c < c_hi; c++) {
            const int lc = c - c_lo, sa = lc % NAS;
            mbar_wait(BAR(B_AFULL + sa), (lc / NAS) & 1);
            if (!p.skip_xform) xform_stage(reinterpret_cast<float4*>(sA + (size_t)sa * p.a_stage_bytes), ncg, R, r_lo, r_mask_hi, p.in_slope, tid2);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(BAR(B_AREADY + sa));
        }
        // ===== tail: partial tile -> scale/mask -> red.add into the output
        mbar_wait(BAR(B_ACC), 0);
        fence_after();
        const bool ok = t < p.T;
        float4* ybp = reinterpret_cast<float4*>(p.y) + (size_t)yb * (p.Cout_total / 4) * p.T;
        const float sc = ((p.out_mask && t >= len) ? 0.f : p.out_scale) * (p.res_mode == 2 ? -1.f : 1.f);
        for (int col0 = 0; col0 < nt; col0 += 16) {
            uint32_t v[16];
            tmem_ld16(trow + (uint32_t)col0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (!ok) continue;
#pragma unroll
            for (int g = 0; g < 4; g++)
                red_add_v4(&ybp[(size_t)((cout_off + col0 + 4 * g) / 4) * p.T + t],
                           make_float4(__uint_as_float(v[4 * g]) * sc, __uint_as_float(v[4 * g + 1]) * sc, __uint_as_float(v[4 * g + 2]) * sc,
                                       __uint_as_float(v[4 * g + 3]) * sc));
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// y (+)= conv(x) with the reduction split `nsplit` ways.  e.accumulate: y already holds the term to add onto; otherwise
// y (the whole tensor: no channel window) is zeroed first.  The weights must be packed with nt == Cout (one N tile).
inline void tc_conv1d_splitk(const TcConvW& w, const float* bias, const Act& x, const Act& y, const TcEpi& e, int nsplit, cudaStream_t st) {
    BV2_CHECK(w.w && x.B == y.B && y.T == x.T && !w.ups_u && !w.x3 && w.nt == w.Cout && w.Cout <= 256, "splitk: one N tile, plain conv");
    BV2_CHECK(!e.relu && !e.out_tf32 && !e.bias_b, "splitk: linear epilogue only");
    BV2_CHECK(e.cin_off % 4 == 0 && e.cout_off % 4 == 0 && e.cin_off + w.Cin <= x.C, "splitk channel window");
    nsplit = std::max(1, std::min(nsplit, w.nchunks));
    TcParams p{};
    p.x = x.p; p.y = y.p; p.w = w.w; p.bias = bias; p.res = e.res; p.lens = e.lens;
    p.Cin_total = x.C; p.cin_off = e.cin_off; p.Cout_total = y.C; p.cout_off = e.cout_off;
    p.res_C_total = e.res_C_total ? e.res_C_total : y.C; p.res_c_off = e.res_c_off;
    p.T = x.T; p.B = x.B; p.K = w.K; p.dil = e.dil; p.pad = (w.K - 1) / 2 * e.dil;
    p.KC = w.KC; p.nchunks = w.nchunks; p.nt = w.nt; p.MT = 1;
    p.R = 128 + (w.K - 1) * e.dil;
    p.a_stage_bytes = (uint32_t)(p.KC * p.R * 4);
    p.w_stage_bytes = (uint32_t)(p.KC * p.nt * 4);
    const uint32_t budget = 100 * 1024;  // two CTAs per SM
    int nas = std::min(3, std::max(2, (p.nchunks + nsplit - 1) / nsplit));
    while (nas > 2 && (size_t)nas * p.a_stage_bytes + 3 * (size_t)p.w_stage_bytes + 1024 > budget) nas--;
    p.nas = nas;
    int nws = ((int)budget - nas * (int)p.a_stage_bytes - 1024) / (int)p.w_stage_bytes;
    p.nws = std::max(2, std::min(nws, 8));
    uint32_t cols = 32; while ((int)cols < p.nt) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.nt >> 3) << 17) | ((128u >> 4) << 24);
    p.in_slope = e.in_slope; p.out_scale = e.out_scale; p.accumulate = 0; p.res_mode = e.res ? (e.res_mode ? e.res_mode : 1) : 0;
    p.in_mask = e.in_mask; p.out_mask = e.out_mask; p.skip_xform = e.skip_xform;
    if (p.in_mask || p.out_mask) BV2_CHECK(e.lens != nullptr, "mask needs lens");
    const size_t smem = (size_t)p.nas * p.a_stage_bytes + (size_t)p.nws * p.w_stage_bytes + (size_t)(3 * p.nas + 2 * p.nws + 2) * 8 + 16;
    BV2_CHECK(smem <= 227 * 1024, "splitk shared memory");
    static bool attr = false;
    if (!attr) { BV2_CUDA(cudaFuncSetAttribute(k_tc_conv1d_splitk, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
    if (!e.accumulate) {
        BV2_CHECK(e.cout_off == 0 && y.C == w.Cout, "splitk without accumulate zeroes the whole output tensor");
        BV2_CUDA(cudaMemsetAsync(y.p, 0, y.elems() * sizeof(float), st));
    }
    launch_pdl(k_tc_conv1d_splitk, dim3(cdiv(p.T, 128), nsplit, p.B), dim3(224), smem, st, p, nsplit);
}

}  // namespace bv2
