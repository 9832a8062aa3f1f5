// Standalone hardware probe for the tcgen05 implicit-GEMM conv (bert_vits2_b200/csrc/tc_conv.cuh):
// validates the smem-descriptor scheme (tap = start-address shift) against a CPU conv on TF32-rounded operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/tc_probe tests/cuda/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../bert_vits2_b200/csrc/tc_conv.cuh"

using namespace bv2;

static std::vector<void*> g_allocs;
static float* up(const std::vector<float>& v) {
    void* p; cudaMalloc(&p, v.size() * 4); cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice); g_allocs.push_back(p);
    return (float*)p;
}

static int run_case(int Cin, int Cout, int K, int dil, int T, int B, float slope, bool res, bool acc, float scale, int iters) {
    std::mt19937 rng(Cin * 131 + Cout * 17 + K * 7 + dil + T);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> x((size_t)B * Cin * T), w((size_t)Cout * Cin * K), bias(Cout), r((size_t)B * Cout * T), y0((size_t)B * Cout * T);
    for (auto& v : x) v = nd(rng);
    for (auto& v : w) v = nd(rng) / std::sqrt((float)(Cin * K));
    for (auto& v : bias) v = nd(rng);
    for (auto& v : r) v = nd(rng);
    for (auto& v : y0) v = nd(rng);
    // c4 layouts
    auto to_c4 = [&](const std::vector<float>& s, int C) {
        std::vector<float> d(s.size());
        for (int b = 0; b < B; b++) for (int c = 0; c < C; c++) for (int t = 0; t < T; t++)
            d[(((size_t)b * (C / 4) + c / 4) * T + t) * 4 + (c & 3)] = s[((size_t)b * C + c) * T + t];
        return d;
    };
    std::function<float*(const std::vector<float>&)> upf = up;
    TcConvW tw = tc_pack_weights(upf, w, Cout, Cin, K);
    Act ax; ax.B = B; ax.C = Cin; ax.T = T; ax.p = up(to_c4(x, Cin));
    Act ay; ay.B = B; ay.C = Cout; ay.T = T; ay.p = up(to_c4(y0, Cout));
    float* dres = up(to_c4(r, Cout));
    float* dbias = up(bias);
    TcEpi e; e.in_slope = slope; e.res = res ? dres : nullptr; e.res_mode = 1; e.accumulate = acc; e.out_scale = scale; e.dil = dil;
    tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
    cudaError_t er = cudaDeviceSynchronize();
    if (er != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(er)); return 1; }
    std::vector<float> got((size_t)B * Cout * T);
    cudaMemcpy(got.data(), ay.p, got.size() * 4, cudaMemcpyDeviceToHost);
    // CPU reference on tf32-rounded operands
    const int pad = (K - 1) / 2 * dil;
    double maxerr = 0, maxref = 0;
    std::vector<float> xa(x.size());
    for (size_t i = 0; i < x.size(); i++) { float v = x[i]; v = v > 0 ? v : v * slope; xa[i] = tf32_rn_host(v); }
    std::vector<float> wr(w.size());
    for (size_t i = 0; i < w.size(); i++) wr[i] = tf32_rn_host(w[i]);
    for (int b = 0; b < B; b++)
        for (int co = 0; co < Cout; co++)
            for (int t = 0; t < T; t += std::max(1, T / 300 - 1)) {
                double s = bias[co];
                for (int ci = 0; ci < Cin; ci++)
                    for (int j = 0; j < K; j++) {
                        int tt = t + j * dil - pad;
                        if (tt >= 0 && tt < T) s += (double)xa[((size_t)b * Cin + ci) * T + tt] * wr[((size_t)co * Cin + ci) * K + j];
                    }
                if (res) s += r[((size_t)b * Cout + co) * T + t];
                if (acc) s += y0[((size_t)b * Cout + co) * T + t];
                s *= scale;
                double g = got[(((size_t)b * (Cout / 4) + co / 4) * T + t) * 4 + (co & 3)];
                maxerr = std::max(maxerr, std::fabs(g - s)); maxref = std::max(maxref, std::fabs(s));
            }
    float ms = 0;
    if (iters > 0) {
        cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
        e.accumulate = 0;
        for (int i = 0; i < 3; i++) tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
        cudaEventRecord(c); cudaEventSynchronize(c); cudaEventElapsedTime(&ms, a, c); ms /= iters;
    }
    double flop = 2.0 * B * T * (double)Cin * Cout * K;
    double bytes = 4.0 * B * T * (Cin + Cout * (1 + (res ? 1 : 0)));
    bool ok = maxerr < 2e-3 * std::max(1.0, maxref);
    printf("%s Cin=%3d Cout=%3d K=%2d dil=%d T=%6d B=%d slope=%.2f res=%d acc=%d : maxerr %.3e (ref max %.2f)", ok ? "PASS" : "FAIL", Cin, Cout, K,
           dil, T, B, slope, (int)res, (int)acc, maxerr, maxref);
    if (iters > 0) printf("  | %.3f ms  %.1f TFLOP/s  %.0f GB/s", ms, flop / ms * 1e-9, bytes / ms * 1e-6);
    printf("\n");
    fflush(stdout);
    return ok ? 0 : 1;
}

static int run_ups(int Cin, int Cout, int K, int u, int T, int B, int iters) {
    std::mt19937 rng(Cin * 13 + Cout + K + u + T);
    std::normal_distribution<float> nd(0.f, 1.f);
    const int To = T * u, p = (K - u) / 2;
    std::vector<float> x((size_t)B * Cin * T), w((size_t)Cin * Cout * K), bias(Cout);
    for (auto& v : x) v = nd(rng);
    for (auto& v : w) v = nd(rng) / std::sqrt((float)(Cin * K / u));
    for (auto& v : bias) v = nd(rng);
    std::vector<float> xc(x.size());
    for (int b = 0; b < B; b++) for (int c = 0; c < Cin; c++) for (int t = 0; t < T; t++)
        xc[(((size_t)b * (Cin / 4) + c / 4) * T + t) * 4 + (c & 3)] = x[((size_t)b * Cin + c) * T + t];
    std::function<float*(const std::vector<float>&)> upf = up;
    TcConvW tw = tc_pack_upsample(upf, w, Cin, Cout, K, u);
    Act ax; ax.B = B; ax.C = Cin; ax.T = T; ax.p = up(xc);
    std::vector<float> y0((size_t)B * Cout * To, 0.f);
    Act ay; ay.B = B; ay.C = Cout; ay.T = To; ay.p = up(y0);
    float* dbias = up(bias);
    TcEpi e; e.in_slope = 0.1f;
    tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
    cudaError_t er = cudaDeviceSynchronize();
    if (er != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(er)); return 1; }
    std::vector<float> got(y0.size());
    cudaMemcpy(got.data(), ay.p, got.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int b = 0; b < B; b++)
        for (int co = 0; co < Cout; co += 3)
            for (int n = 0; n < To; n += std::max(1, To / 300)) {
                double s = bias[co];
                for (int i = 0; i < T; i++) {
                    int j = n + p - i * u;
                    if (j < 0 || j >= K) continue;
                    for (int ci = 0; ci < Cin; ci++) {
                        float xv = x[((size_t)b * Cin + ci) * T + i]; xv = xv > 0 ? xv : 0.1f * xv;
                        s += (double)tf32_rn_host(xv) * tf32_rn_host(w[((size_t)ci * Cout + co) * K + j]);
                    }
                }
                double g = got[(((size_t)b * (Cout / 4) + co / 4) * To + n) * 4 + (co & 3)];
                maxerr = std::max(maxerr, std::fabs(g - s)); maxref = std::max(maxref, std::fabs(s));
            }
    float ms = 0;
    if (iters > 0) {
        cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
        for (int i = 0; i < 3; i++) tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
        cudaEventRecord(c); cudaEventSynchronize(c); cudaEventElapsedTime(&ms, a, c); ms /= iters;
    }
    bool ok = maxerr < 2e-3 * std::max(1.0, maxref);
    printf("%s UPS Cin=%3d Cout=%3d K=%2d u=%d T=%6d B=%d (Kp=%d) : maxerr %.3e (ref max %.2f)", ok ? "PASS" : "FAIL", Cin, Cout, K, u, T, B, tw.K, maxerr, maxref);
    if (iters > 0) printf("  | %.3f ms", ms);
    printf("\n"); fflush(stdout);
    return ok ? 0 : 1;
}

// 3xTF32 accuracy: compare against a double-precision conv on the UNROUNDED operands
static int run_x3(int Cin, int Cout, int K, int T, int nt) {
    std::mt19937 rng(Cin + Cout * 3 + K);
    std::normal_distribution<float> nd(0.f, 1.f);
    const int B = 1;
    std::vector<float> x((size_t)Cin * T), w((size_t)Cout * Cin * K), bias(Cout);
    for (auto& v : x) v = nd(rng);
    for (auto& v : w) v = nd(rng) / std::sqrt((float)(Cin * K));
    for (auto& v : bias) v = nd(rng);
    std::vector<float> xc(x.size());
    for (int c = 0; c < Cin; c++) for (int t = 0; t < T; t++) xc[((size_t)(c / 4) * T + t) * 4 + (c & 3)] = x[(size_t)c * T + t];
    std::function<float*(const std::vector<float>&)> upf = up;
    double errs[2];
    for (int mode = 0; mode < 2; mode++) {
        TcConvW tw = tc_pack_weights(upf, w, Cout, Cin, K, nt, mode);
        Act ax; ax.B = B; ax.C = Cin; ax.T = T; ax.p = up(xc);
        std::vector<float> y0((size_t)Cout * T, 0.f);
        Act ay; ay.B = B; ay.C = Cout; ay.T = T; ay.p = up(y0);
        float* dbias = up(bias);
        TcEpi e;
        tc_conv1d(tw, dbias, ax, ay, e, 0, 148);
        cudaError_t er = cudaDeviceSynchronize();
        if (er != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(er)); return 1; }
        std::vector<float> got(y0.size());
        cudaMemcpy(got.data(), ay.p, got.size() * 4, cudaMemcpyDeviceToHost);
        const int pad = (K - 1) / 2;
        double maxerr = 0;
        for (int co = 0; co < Cout; co++)
            for (int t = 0; t < T; t++) {
                double s = bias[co];
                for (int ci = 0; ci < Cin; ci++)
                    for (int j = 0; j < K; j++) { int tt = t + j - pad; if (tt >= 0 && tt < T) s += (double)x[(size_t)ci * T + tt] * w[((size_t)co * Cin + ci) * K + j]; }
                maxerr = std::max(maxerr, std::fabs((double)got[((size_t)(co / 4) * T + t) * 4 + (co & 3)] - s));
            }
        errs[mode] = maxerr;
    }
    // informational: documents that the tcgen05 FP32 accumulator truncates (error grows ~6.6e-8 per accumulated product),
    // which is why the stages feeding ceil(durations) stay on FP32 FMA
    bool ok = errs[1] < errs[0];
    printf("%s X3 Cin=%4d Cout=%3d K=%d T=%d nt=%d : max err tf32 %.3e, 3xtf32 %.3e\n", ok ? "PASS" : "FAIL", Cin, Cout, K, T, nt, errs[0], errs[1]);
    fflush(stdout);
    return ok ? 0 : 1;
}

// Fused ResBlock pair: y = (conv2(lrelu(conv1(lrelu(x)))) + x [+ y0]) * scale, checked at sampled time steps.
static int run_pair(int C, int K, int dil, int T, int B, bool acc, float scale, int iters, int kc = 0, int persist = 0) {
    auto launch = [&](const TcConvW& a, const TcConvW& b, const float* b1, const float* b2, const Act& x, const Act& y, int acc_) {
        return persist ? tc_pair_persist(a, b, b1, b2, x, y, dil, scale, acc_, 0, 148) : tc_pair(a, b, b1, b2, x, y, dil, scale, acc_, 0);
    };
    std::mt19937 rng(C * 31 + K * 7 + dil + T);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> x((size_t)B * C * T), w1((size_t)C * C * K), w2((size_t)C * C * K), b1(C), b2(C), y0((size_t)B * C * T);
    for (auto& v : x) v = nd(rng);
    for (auto& v : w1) v = nd(rng) / std::sqrt((float)(C * K));
    for (auto& v : w2) v = nd(rng) / std::sqrt((float)(C * K));
    for (auto& v : b1) v = nd(rng);
    for (auto& v : b2) v = nd(rng);
    for (auto& v : y0) v = nd(rng);
    auto to_c4 = [&](const std::vector<float>& s) {
        std::vector<float> d(s.size());
        for (int b = 0; b < B; b++) for (int c = 0; c < C; c++) for (int t = 0; t < T; t++)
            d[(((size_t)b * (C / 4) + c / 4) * T + t) * 4 + (c & 3)] = s[((size_t)b * C + c) * T + t];
        return d;
    };
    std::function<float*(const std::vector<float>&)> upf = up;
    TcConvW tw1 = tc_pack_weights(upf, w1, C, C, K, C, 0, kc), tw2 = tc_pack_weights(upf, w2, C, C, K, C, 0, kc);
    Act ax; ax.B = B; ax.C = C; ax.T = T; ax.p = up(to_c4(x));
    Act ay; ay.B = B; ay.C = C; ay.T = T; ay.p = up(to_c4(y0));
    float* db1 = up(b1); float* db2 = up(b2);
    if (!launch(tw1, tw2, db1, db2, ax, ay, acc ? 1 : 0)) { printf("SKIP pair C=%d K=%d (does not fit)\n", C, K); return 0; }
    cudaError_t er = cudaDeviceSynchronize();
    if (er != cudaSuccess) { printf("CUDA error (pair C=%d K=%d d=%d): %s\n", C, K, dil, cudaGetErrorString(er)); return 1; }
    std::vector<float> got((size_t)B * C * T);
    cudaMemcpy(got.data(), ay.p, got.size() * 4, cudaMemcpyDeviceToHost);
    const int p2 = (K - 1) / 2, p1 = p2 * dil;
    std::vector<float> xa(x.size()), w1r(w1.size()), w2r(w2.size());
    for (size_t i = 0; i < x.size(); i++) { float v = x[i]; v = v > 0 ? v : v * 0.1f; xa[i] = tf32_rn_host(v); }
    for (size_t i = 0; i < w1.size(); i++) { w1r[i] = tf32_rn_host(w1[i]); w2r[i] = tf32_rn_host(w2[i]); }
    double maxerr = 0, maxref = 0;
    std::vector<float> xt((size_t)C * K);
    std::vector<int> ts;
    for (int t = 0; t < T; t += std::max(1, T / 150 - 1)) ts.push_back(t);
    for (int t : {1, 2, T - 2, T - 1, 117, 118, 119, 127, 128, 129}) if (t >= 0 && t < T) ts.push_back(t);
    for (int b = 0; b < B; b++)
        for (int t : ts) {
            for (int c = 0; c < C; c++)
                for (int j2 = 0; j2 < K; j2++) {
                    const int tt = t + j2 - p2;
                    float o = 0.f;
                    if (tt >= 0 && tt < T) {
                        double s = b1[c];
                        for (int ci = 0; ci < C; ci++)
                            for (int j = 0; j < K; j++) {
                                const int tx = tt + j * dil - p1;
                                if (tx >= 0 && tx < T) s += (double)xa[((size_t)b * C + ci) * T + tx] * w1r[((size_t)c * C + ci) * K + j];
                            }
                        float v = (float)s; v = v > 0 ? v : v * 0.1f; o = tf32_rn_host(v);
                    }
                    xt[(size_t)c * K + j2] = o;
                }
            for (int co = 0; co < C; co++) {
                double s = b2[co];
                for (int c = 0; c < C; c++)
                    for (int j2 = 0; j2 < K; j2++) s += (double)xt[(size_t)c * K + j2] * w2r[((size_t)co * C + c) * K + j2];
                s += x[((size_t)b * C + co) * T + t];
                if (acc) s += y0[((size_t)b * C + co) * T + t];
                s *= scale;
                const double g = got[(((size_t)b * (C / 4) + co / 4) * T + t) * 4 + (co & 3)];
                maxerr = std::max(maxerr, std::fabs(g - s)); maxref = std::max(maxref, std::fabs(s));
            }
        }
    float ms = 0, ms2 = 0;
    if (iters > 0) {
        cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
        for (int i = 0; i < 3; i++) launch(tw1, tw2, db1, db2, ax, ay, 0);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) launch(tw1, tw2, db1, db2, ax, ay, 0);
        cudaEventRecord(c); cudaEventSynchronize(c); cudaEventElapsedTime(&ms, a, c); ms /= iters;
        // the two-launch path it replaces
        Act am; am.B = B; am.C = C; am.T = T; am.p = up(std::vector<float>((size_t)B * C * T));
        TcEpi e1; e1.in_slope = 0.1f; e1.dil = dil;
        TcEpi e2; e2.in_slope = 0.1f; e2.dil = 1; e2.res = ax.p; e2.res_mode = 1; e2.out_scale = scale;
        for (int i = 0; i < 3; i++) { tc_conv1d(tw1, db1, ax, am, e1, 0, 148); tc_conv1d(tw2, db2, am, ay, e2, 0, 148); }
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) { tc_conv1d(tw1, db1, ax, am, e1, 0, 148); tc_conv1d(tw2, db2, am, ay, e2, 0, 148); }
        cudaEventRecord(c); cudaEventSynchronize(c); cudaEventElapsedTime(&ms2, a, c); ms2 /= iters;
    }
    bool ok = maxerr < 2e-3 * std::max(1.0, maxref);
    printf("%s %s C=%3d K=%2d dil=%d T=%6d B=%d acc=%d kc=%d : maxerr %.3e (ref max %.2f)", ok ? "PASS" : "FAIL", persist ? "PPAIR" : "PAIR", C, K, dil, T, B, (int)acc, tw1.KC, maxerr, maxref);
    if (iters > 0) printf("  | fused %.3f ms  vs two launches %.3f ms  (%.2fx)", ms, ms2, ms2 / ms);
    printf("\n");
    fflush(stdout);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    int fails = 0;
    bool perf = argc > 1;
    try {
        // functional: K=1 first (no tap shift), then taps with shifts not multiple of 8 rows
        fails += run_case(16, 16, 1, 1, 300, 1, 1.f, false, false, 1.f, 0);
        fails += run_case(64, 64, 1, 1, 300, 2, 1.f, false, false, 1.f, 0);
        fails += run_case(16, 16, 3, 1, 300, 1, 0.1f, false, false, 1.f, 0);
        fails += run_case(32, 32, 7, 3, 1000, 2, 0.1f, true, false, 1.f, 0);
        fails += run_case(64, 64, 11, 5, 700, 1, 0.1f, true, true, 1.f / 3, 0);
        fails += run_case(128, 128, 3, 1, 256, 1, 0.1f, false, false, 1.f, 0);
        fails += run_case(256, 256, 11, 5, 500, 1, 0.1f, true, false, 1.f, 0);
        fails += run_case(256, 256, 7, 1, 128, 3, 0.1f, false, true, 1.f, 0);
        fails += run_case(192, 192, 5, 1, 333, 1, 1.f, false, false, 1.f, 0);
        fails += run_case(192, 768, 5, 1, 1573, 1, 1.f, false, false, 1.f, perf ? 10 : 0);   // flow FFN conv_1 (N tiles)
        fails += run_case(768, 192, 5, 1, 1573, 1, 1.f, false, false, 1.f, perf ? 10 : 0);   // flow FFN conv_2
        fails += run_case(192, 576, 1, 1, 1573, 2, 1.f, false, false, 1.f, perf ? 10 : 0);   // fused QKV
        fails += run_case(96, 192, 1, 1, 700, 1, 1.f, false, false, 1.f, 0);
        fails += run_case(192, 512, 7, 1, 1573, 1, 1.f, false, false, 1.f, perf ? 10 : 0);   // conv_pre
        fails += run_case(32, 32, 7, 3, 20000, 3, 0.1f, true, false, 1.f, 0);        // persistent kernel (narrow, many tiles), batched
        fails += run_case(16, 16, 11, 5, 40001, 1, 0.1f, true, true, 1.f / 3, 0);    // persistent kernel, residual + accumulate + scale
        fails += run_case(32, 32, 3, 1, 38000, 2, 0.1f, false, false, 1.f, 0);
        fails += run_pair(16, 3, 1, 300, 1, false, 1.f, 0);
        fails += run_pair(32, 7, 3, 1000, 2, false, 1.f, 0);
        fails += run_pair(64, 11, 5, 700, 1, true, 1.f / 3, 0);
        fails += run_pair(128, 3, 5, 517, 1, false, 1.f, 0);
        fails += run_pair(128, 11, 5, 1300, 2, true, 1.f / 3, 0);
        fails += run_pair(64, 7, 1, 118, 1, false, 1.f, 0);
        if (getenv("PROBE_PPAIR")) {  // persistent pair kernel (experimental)
            fails += run_pair(16, 3, 1, 300, 1, false, 1.f, 0, 0, 1);
            fails += run_pair(32, 7, 3, 1000, 2, false, 1.f, 0, 0, 1);
            fails += run_pair(16, 11, 5, 40001, 1, true, 1.f / 3, 0, 0, 1);
            fails += run_pair(32, 11, 5, 5000, 3, true, 1.f / 3, 0, 0, 1);
            fails += run_pair(32, 3, 1, 117, 1, false, 1.f, 0, 0, 1);
            if (perf) {
                int F = 1573;
                for (int C : {32, 16})
                    for (int k : {3, 7, 11})
                        for (int d : {1, 5}) fails += run_pair(C, k, d, (C == 32 ? 256 : 512) * F, 1, false, 1.f, 10, 0, 1);
            }
            printf("%s (%d failing)\n", fails ? "PROBE FAILED" : "PROBE OK", fails);
            return fails ? 1 : 0;
        }
        fails += run_x3(32, 32, 1, 100, 32);
        fails += run_x3(192, 192, 3, 256, 32);
        fails += run_x3(768, 192, 3, 256, 32);
        fails += run_ups(512, 256, 16, 8, 300, 1, 0);
        fails += run_ups(256, 128, 16, 8, 257, 2, 0);
        fails += run_ups(128, 64, 8, 2, 1000, 1, 0);
        fails += run_ups(64, 32, 2, 2, 900, 1, 0);
        fails += run_ups(32, 16, 2, 2, 1111, 1, 0);
        if (perf) {
            fails += run_ups(512, 256, 16, 8, 1573, 1, 10);
            fails += run_ups(256, 128, 16, 8, 1573 * 8, 1, 10);
            fails += run_ups(128, 64, 8, 2, 1573 * 64, 1, 10);
            fails += run_ups(64, 32, 2, 2, 1573 * 128, 1, 10);
            fails += run_ups(32, 16, 2, 2, 1573 * 256, 1, 10);
        }
        if (perf) {
            int F = 1573;
            int Cs[4] = {128, 64, 32, 16}; int Ls[4] = {64, 128, 256, 512};
            for (int s = 0; s < 4; s++)
                for (int k : {3, 7, 11})
                    for (int d : {1, 5}) fails += run_pair(Cs[s], k, d, Ls[s] * F, 1, false, 1.f, 10);
            for (int k : {3, 11}) fails += run_pair(128, k, 3, 64 * F, 1, false, 1.f, 10, 16);
        }
        if (perf) {
            // Generator MRF shapes at F=1024 frames
            int F = 1024;
            int Cs[5] = {256, 128, 64, 32, 16}; int Ls[5] = {8, 64, 128, 256, 512};
            for (int s = 0; s < 5; s++)
                for (int k : {3, 7, 11})
                    for (int d : {1, 5}) fails += run_case(Cs[s], Cs[s], k, d, Ls[s] * F, 1, 0.1f, true, false, 1.f, 10);
        }
    } catch (const std::exception& ex) { printf("exception: %s\n", ex.what()); return 2; }
    printf("%s (%d failing)\n", fails ? "PROBE FAILED" : "PROBE OK", fails);
    return fails ? 1 : 0;
}
