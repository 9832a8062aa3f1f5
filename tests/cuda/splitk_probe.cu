// Hardware probe for the EXPERIMENTAL split-K conv (bert_vits2_b200/csrc/experimental/tc_splitk.cuh): correctness against
// a CPU conv on TF32-rounded operands and time against the split-N launch the engine uses today (tc_conv1d).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tests/cuda/splitk_probe tests/cuda/splitk_probe.cu
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../bert_vits2_b200/csrc/experimental/tc_splitk.cuh"

using namespace bv2;

static float* up(const std::vector<float>& v) {
    void* p; cudaMalloc(&p, v.size() * 4); cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    return (float*)p;
}

static int run(int Cin, int Cout, int K, int T, int B, int nsplit, int nt_base, bool res, bool acc, int iters) {
    std::mt19937 rng(Cin * 31 + Cout * 7 + K + T + nsplit);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> x((size_t)B * Cin * T), w((size_t)Cout * Cin * K), bias(Cout), r((size_t)B * Cout * T), y0((size_t)B * Cout * T);
    for (auto& v : x) v = nd(rng);
    for (auto& v : w) v = nd(rng) / std::sqrt((float)(Cin * K));
    for (auto& v : bias) v = nd(rng);
    for (auto& v : r) v = nd(rng);
    for (auto& v : y0) v = nd(rng);
    auto to_c4 = [&](const std::vector<float>& s, int C) {
        std::vector<float> d(s.size());
        for (int b = 0; b < B; b++) for (int c = 0; c < C; c++) for (int t = 0; t < T; t++)
            d[(((size_t)b * (C / 4) + c / 4) * T + t) * 4 + (c & 3)] = s[((size_t)b * C + c) * T + t];
        return d;
    };
    std::function<float*(const std::vector<float>&)> upf = up;
    TcConvW wk = tc_pack_weights(upf, w, Cout, Cin, K, Cout, 0, 32);        // one N tile (split-K)
    TcConvW wn = tc_pack_weights(upf, w, Cout, Cin, K, nt_base, 0, 64);     // the engine's split-N packing
    Act ax; ax.B = B; ax.C = Cin; ax.T = T; ax.p = up(to_c4(x, Cin));
    Act ay; ay.B = B; ay.C = Cout; ay.T = T; ay.p = up(to_c4(y0, Cout));
    float* dres = up(to_c4(r, Cout));
    float* dbias = up(bias);
    TcEpi e; e.in_slope = 1.f; e.res = res ? dres : nullptr; e.res_mode = 1; e.accumulate = acc; e.dil = 1;
    tc_conv1d_splitk(wk, dbias, ax, ay, e, nsplit, 0);
    cudaError_t er = cudaDeviceSynchronize();
    if (er != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(er)); return 1; }
    std::vector<float> got((size_t)B * Cout * T);
    cudaMemcpy(got.data(), ay.p, got.size() * 4, cudaMemcpyDeviceToHost);
    const int pad = (K - 1) / 2;
    std::vector<float> xa(x.size()), wr(w.size());
    for (size_t i = 0; i < x.size(); i++) xa[i] = tf32_rn_host(x[i]);
    for (size_t i = 0; i < w.size(); i++) wr[i] = tf32_rn_host(w[i]);
    double maxerr = 0, maxref = 0;
    for (int b = 0; b < B; b++)
        for (int co = 0; co < Cout; co++)
            for (int t = 0; t < T; t += std::max(1, T / 200 - 1)) {
                double s = bias[co];
                for (int ci = 0; ci < Cin; ci++)
                    for (int j = 0; j < K; j++) {
                        const int tt = t + j - pad;
                        if (tt >= 0 && tt < T) s += (double)xa[((size_t)b * Cin + ci) * T + tt] * wr[((size_t)co * Cin + ci) * K + j];
                    }
                if (res) s += r[((size_t)b * Cout + co) * T + t];
                if (acc) s += y0[((size_t)b * Cout + co) * T + t];
                const double g = got[(((size_t)b * (Cout / 4) + co / 4) * T + t) * 4 + (co & 3)];
                maxerr = std::max(maxerr, std::fabs(g - s)); maxref = std::max(maxref, std::fabs(s));
            }
    float ms = 0, ms2 = 0;
    if (iters > 0) {
        cudaEvent_t a, c; cudaEventCreate(&a); cudaEventCreate(&c);
        e.accumulate = 0;
        for (int i = 0; i < 3; i++) tc_conv1d_splitk(wk, dbias, ax, ay, e, nsplit, 0);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) tc_conv1d_splitk(wk, dbias, ax, ay, e, nsplit, 0);
        cudaEventRecord(c); cudaEventSynchronize(c); cudaEventElapsedTime(&ms, a, c); ms /= iters;
        for (int i = 0; i < 3; i++) tc_conv1d(wn, dbias, ax, ay, e, 0, 148);
        cudaEventRecord(a);
        for (int i = 0; i < iters; i++) tc_conv1d(wn, dbias, ax, ay, e, 0, 148);
        cudaEventRecord(c); cudaEventSynchronize(c); cudaEventElapsedTime(&ms2, a, c); ms2 /= iters;
    }
    const bool ok = maxerr < 2e-3 * std::max(1.0, maxref);
    printf("%s SPLITK Cin=%3d Cout=%3d K=%d T=%5d B=%d nsplit=%d res=%d acc=%d : maxerr %.3e (ref max %.2f)", ok ? "PASS" : "FAIL", Cin, Cout, K, T, B,
           nsplit, (int)res, (int)acc, maxerr, maxref);
    if (iters > 0) printf("  | split-K %.1f us (incl. memset)  vs split-N nt=%d %.1f us  (%.2fx)", ms * 1e3, nt_base, ms2 * 1e3, ms2 / ms);
    printf("\n");
    fflush(stdout);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    int fails = 0;
    const bool perf = argc > 1;
    try {
        fails += run(64, 32, 1, 300, 1, 2, 32, false, false, 0);
        fails += run(192, 192, 3, 700, 2, 3, 48, true, false, 0);
        fails += run(768, 192, 5, 1573, 1, 6, 32, false, false, perf ? 20 : 0);   // flow FFN conv_2
        fails += run(768, 192, 5, 1573, 1, 4, 32, false, true, perf ? 20 : 0);
        fails += run(768, 192, 5, 1573, 1, 8, 32, false, false, perf ? 20 : 0);
        fails += run(192, 192, 1, 1573, 1, 3, 48, true, false, perf ? 20 : 0);    // conv_o-like
        fails += run(768, 192, 5, 825, 32, 2, 32, false, false, perf ? 5 : 0);    // config 3 (B=32): fewer splits needed
    } catch (const std::exception& ex) { printf("exception: %s\n", ex.what()); return 2; }
    printf("%s (%d failing)\n", fails ? "SPLITK PROBE FAILED" : "SPLITK PROBE OK", fails);
    return fails ? 1 : 0;
}
