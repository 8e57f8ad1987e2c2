// Standalone GPU self-test + micro-benchmark of the tcgen05 implicit-GEMM conv (no Python, no torch).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo conv_tc.cu test_conv_tc.cu -o test_conv_tc
// Run:    ./test_conv_tc [check|perf|all]
// Correctness reference: direct convolution in double over the same fp16-rounded operands (exact up to fp32
// accumulation order), so the tolerance can be tight.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "conv_tc.cuh"

using namespace mi;
namespace mi { bool pdl_enabled() { return getenv("MI_PDL") != nullptr; } }   // launch.cuh hook (capi.cu in the library)

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #x); \
            fflush(stdout);                                                                     \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

static uint32_t rng_state = 12345u;
static float frand() {   // uniform in [-1, 1)
    rng_state = rng_state * 1664525u + 1013904223u;
    return ((rng_state >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

struct Case {
    const char* name;
    int B, H, W, Cin, Cout, k, stride;   // stride 1 (k=1|3) or 2 (k=4, pad 1)
    bool bias, residual, f16out;
    int hint;
    int pair;   // 0 auto, 1 force 1-CTA kernel, 2 force CTA-pair kernel (stream-K auto), 3 halo kernel, 4 pair with one
                // k-chunk per stage, 5 pair + stream-K forced, 6 pair without stream-K, 11..13 profiling switches
};

static int* g_err = nullptr;   // host-mapped

// returns max abs error (or -1 on failure)
static double run_case(const Case& c, bool check, int reps, double* ms_out) {
    const int Ho = c.H / c.stride, Wo = c.W / c.stride;
    const bool vert = c.k == 15;                         // 15 x 1 vertical conv (the stem geometry), pad 7
    const int pad = (c.k == 3) ? 1 : (c.k == 4 ? 1 : (vert ? 7 : 0));
    const int taps = vert ? 15 : c.k * c.k;
    const int kw_ = vert ? 1 : c.k;
    const size_t n_in = (size_t)c.B * c.H * c.W * c.Cin;
    const size_t n_w = (size_t)c.Cout * taps * c.Cin;
    const size_t n_out = (size_t)c.B * Ho * Wo * c.Cout;
    std::vector<__half> h_in(n_in), h_w(n_w);
    std::vector<float> f_in, f_w;
    std::vector<float> h_bias(c.Cout), h_res;
    if (check) { f_in.resize(n_in); f_w.resize(n_w); }
    for (size_t i = 0; i < n_in; ++i) { __half v = __float2half_rn(frand()); h_in[i] = v; if (check) f_in[i] = __half2float(v); }
    const float wscale = 1.0f / sqrtf((float)(taps * c.Cin));
    for (size_t i = 0; i < n_w; ++i) { __half v = __float2half_rn(frand() * wscale * 4.f); h_w[i] = v; if (check) f_w[i] = __half2float(v); }
    for (int i = 0; i < c.Cout; ++i) h_bias[i] = frand();
    if (c.residual) { h_res.resize(n_out); for (size_t i = 0; i < n_out; ++i) h_res[i] = frand(); }

    // device input layout: stride 1 -> [B][1][H][W][C]; stride 2 -> phase split [B][4][Ho][Wo][C], phase = (h&1)*2 + (w&1)
    std::vector<__half> h_dev_in(n_in);
    const bool direct_s2 = (c.stride == 2 && c.pair == 9);   // stride-2 conv read in place (TMA element strides)
    if (c.stride == 1 || direct_s2) {
        h_dev_in = h_in;
    } else {
        for (int b = 0; b < c.B; ++b)
            for (int h = 0; h < c.H; ++h)
                for (int w = 0; w < c.W; ++w) {
                    const int p = (h & 1) * 2 + (w & 1);
                    const size_t dst = ((((size_t)b * 4 + p) * Ho + (h >> 1)) * Wo + (w >> 1)) * c.Cin;
                    const size_t src = (((size_t)b * c.H + h) * c.W + w) * c.Cin;
                    memcpy(&h_dev_in[dst], &h_in[src], c.Cin * sizeof(__half));
                }
    }
    // weights are generated directly in packed order [Cout][tap][Cin], tap = r*k + s

    __half *d_in, *d_w, *d_o16 = nullptr;
    float *d_o32, *d_bias = nullptr, *d_res = nullptr;
    CK(cudaMalloc(&d_in, n_in * 2));
    CK(cudaMalloc(&d_w, n_w * 2));
    CK(cudaMalloc(&d_o32, n_out * 4));
    CK(cudaMemset(d_o32, 0xFF, n_out * 4));
    CK(cudaMemcpy(d_in, h_dev_in.data(), n_in * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, h_w.data(), n_w * 2, cudaMemcpyHostToDevice));
    if (c.bias) { CK(cudaMalloc(&d_bias, c.Cout * 4)); CK(cudaMemcpy(d_bias, h_bias.data(), c.Cout * 4, cudaMemcpyHostToDevice)); }
    if (c.residual) { CK(cudaMalloc(&d_res, n_out * 4)); CK(cudaMemcpy(d_res, h_res.data(), n_out * 4, cudaMemcpyHostToDevice)); }
    if (c.f16out) { CK(cudaMalloc(&d_o16, n_out * 2)); CK(cudaMemset(d_o16, 0xFF, n_out * 2)); }

    ConvTcProblem p{};
    p.act = d_in; p.B = c.B; p.H = Ho; p.W = Wo; p.phases = (c.stride == 2 && c.pair != 9) ? 4 : 1; p.in_stride = (c.stride == 2 && c.pair == 9) ? 2 : 1;
    p.lda = c.Cin; p.a_channels = c.Cin; p.a_chan_off = 0; p.Cin = c.Cin;
    p.wpacked = d_w; p.Cout = c.Cout; p.num_taps = taps;
    for (int r = 0; r < c.k; ++r)
        for (int s = 0; s < kw_; ++s) {
            const int t = r * kw_ + s;
            if (vert) { p.dh[t] = r - pad; p.dw[t] = 0; p.ph[t] = 0; }
            else if (c.stride == 1 || direct_s2) { p.dh[t] = r - pad; p.dw[t] = s - pad; p.ph[t] = 0; }
            else {
                // input row = 2*ho + r - 1  ->  phase row (r-1)&1, block shift floor((r-1)/2)
                const int rr = r - 1, ss = s - 1;
                p.dh[t] = (rr < 0) ? -1 : (rr >> 1); p.dw[t] = (ss < 0) ? -1 : (ss >> 1);
                p.ph[t] = (rr & 1) * 2 + (ss & 1);
            }
        }
    p.out_f32 = d_o32; p.out_f16 = d_o16; p.bias = d_bias; p.residual = d_res;
    p.out_sw = c.Cout; p.out_sh = (long long)Wo * c.Cout; p.out_sb = (long long)Ho * Wo * c.Cout;
    p.block_n_hint = c.hint; p.cta_pair = (c.pair == 3) ? 1 : c.pair; p.halo = (c.pair == 3) ? 2 : (c.pair == 8 ? (c.k == 15 ? 3 : 1) : 0); if (c.pair == 8) p.cta_pair = 1; p.dbg = (c.pair >= 10) ? c.pair - 10 : 0; if (c.pair == 7) p.cta_pair = 2; if (c.pair >= 10) p.cta_pair = 1; if (c.pair == 4) { p.cta_pair = 2; p.kmerge = 1; } p.err_flag = g_err;
    if (c.pair == 5 || c.pair == 6 || c.pair == 7) p.cta_pair = 2;      // (former stream-K cases: plain CTA-pair kernel)
    if (c.pair == 9) p.cta_pair = 0;

    *g_err = 0;
    int rc = conv_tc_launch(p, 0);
    if (rc != 0) { printf("[%s] launch rc=%d (%s)\n", c.name, rc, conv_tc_strerror(rc)); return -1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("[%s] KERNEL FAILED: %s, err_flag=%d\n", c.name, cudaGetErrorString(e), *g_err);
        fflush(stdout);
        exit(3);
    }
    if (reps > 0) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int i = 0; i < 3; ++i) conv_tc_launch(p, 0);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) conv_tc_launch(p, 0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        *ms_out = ms / reps;
    }

    double maxerr = 0, maxref = 0;
    if (check) {
        std::vector<float> h_out(n_out);
        std::vector<__half> h_out16;
        CK(cudaMemcpy(h_out.data(), d_o32, n_out * 4, cudaMemcpyDeviceToHost));
        if (c.f16out) { h_out16.resize(n_out); CK(cudaMemcpy(h_out16.data(), d_o16, n_out * 2, cudaMemcpyDeviceToHost)); }
        double max16 = 0;
        size_t bad_i = 0;
        for (int b = 0; b < c.B; ++b)
            for (int ho = 0; ho < Ho; ++ho)
                for (int wo = 0; wo < Wo; ++wo)
                    for (int n = 0; n < c.Cout; ++n) {
                        double acc = 0;
                        for (int r = 0; r < c.k; ++r) {
                            const int hi = ho * c.stride + r - pad;
                            if (hi < 0 || hi >= c.H) continue;
                            for (int s = 0; s < kw_; ++s) {
                                const int wi = wo * c.stride + s - (vert ? 0 : pad);
                                if (wi < 0 || wi >= c.W) continue;
                                const float* a = &f_in[(((size_t)b * c.H + hi) * c.W + wi) * c.Cin];
                                const float* wv = &f_w[((size_t)n * taps + r * kw_ + s) * c.Cin];
                                double d = 0;
                                for (int ci = 0; ci < c.Cin; ++ci) d += (double)a[ci] * wv[ci];
                                acc += d;
                            }
                        }
                        const size_t oi = (((size_t)b * Ho + ho) * Wo + wo) * c.Cout + n;
                        if (c.bias) acc += h_bias[n];
                        if (c.residual) acc += h_res[oi];
                        const double err = fabs(acc - (double)h_out[oi]);
                        if (!(err <= maxerr)) { maxerr = err; bad_i = oi; }   // also catches NaN
                        if (fabs(acc) > maxref) maxref = fabs(acc);
                        if (c.f16out) {
                            const double e16 = fabs(acc - (double)__half2float(h_out16[oi]));
                            if (!(e16 <= max16)) max16 = e16;
                        }
                    }
        printf("[%s] B=%d %dx%d Cin=%d Cout=%d k=%d s=%d hint=%d : max_abs_err=%.3e (max |ref|=%.3f, worst idx %zu)%s",
               c.name, c.B, c.H, c.W, c.Cin, c.Cout, c.k, c.stride, c.hint, maxerr, maxref, bad_i,
               (maxerr <= 2e-3 * (1 + maxref)) ? "  OK" : "  **MISMATCH**");
        if (c.f16out) printf("  f16out_err=%.3e%s", max16, (max16 <= 4e-3 * (1 + maxref)) ? " OK" : " **MISMATCH**");
        printf("\n");
        if (!(maxerr <= 2e-3 * (1 + maxref))) {
            // dump a few values to help diagnose layout bugs
            for (int i = 0; i < 8; ++i) printf("   out[%d]=%f\n", i, h_out[i]);
        }
    }
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_o32);
    if (d_bias) cudaFree(d_bias);
    if (d_res) cudaFree(d_res);
    if (d_o16) cudaFree(d_o16);
    fflush(stdout);
    return maxerr;
}

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "all";
    CK(cudaSetDeviceFlags(cudaDeviceMapHost));
    CK(cudaHostAlloc(&g_err, sizeof(int), cudaHostAllocMapped));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);

    int failures = 0;
    if (!strcmp(mode, "check") || !strcmp(mode, "all")) {
        const Case cases[] = {
            {"gemm1x1_n64", 1, 16, 16, 64, 64, 1, 1, false, false, false, 0, 1},
            {"gemm1x1_k128", 1, 16, 16, 128, 64, 1, 1, false, false, false, 0, 1},
            {"c3_16x16", 2, 16, 16, 64, 128, 3, 1, false, false, false, 0, 1},
            {"c3_32x32", 1, 32, 32, 128, 128, 3, 1, true, false, false, 0, 1},
            {"c3_8x8_bb2", 3, 8, 8, 64, 64, 3, 1, true, true, true, 0, 1},
            {"c3_w256", 1, 4, 256, 64, 64, 3, 1, false, false, false, 0, 1},
            {"c3_n256", 2, 16, 16, 128, 256, 3, 1, true, true, true, 256, 1},
            {"c3_n512", 2, 16, 16, 64, 512, 3, 1, true, false, false, 256, 1},
            {"c3_n16", 1, 16, 16, 64, 16, 3, 1, true, false, false, 0, 1},
            {"c3_n32", 1, 16, 16, 64, 32, 3, 1, true, false, false, 0, 1},
            {"c4_s2", 2, 32, 32, 64, 128, 4, 2, true, false, true, 0, 1},
            {"c3_persist", 8, 64, 64, 64, 64, 3, 1, true, true, false, 0, 1},
            {"c3_deepk", 1, 16, 16, 1024, 128, 3, 1, false, false, false, 0, 1},
            // ---- 3x3 halo-tile kernel (pair == 3)
            {"h_c3_16x16_n128", 2, 16, 16, 64, 128, 3, 1, true, false, false, 128, 3},
            {"h_c3_16x16_n256", 2, 16, 16, 128, 256, 3, 1, true, true, true, 256, 3},
            {"h_c3_32x32", 1, 32, 32, 128, 128, 3, 1, true, false, false, 128, 3},
            {"h_c3_persist", 8, 64, 64, 64, 128, 3, 1, true, true, false, 128, 3},
            {"h_c3_deepk", 2, 16, 16, 1024, 256, 3, 1, false, true, false, 256, 3},
            {"h_c3_w256", 1, 16, 256, 64, 128, 3, 1, false, false, false, 128, 3},
            {"h_c3_n16", 2, 32, 32, 128, 16, 3, 1, true, false, false, 0, 3},
            {"h_c3_n16_k256", 3, 32, 64, 256, 16, 3, 1, true, false, false, 0, 3},
            {"h_c3_n16_persist", 4, 128, 128, 64, 16, 3, 1, false, false, false, 0, 3},
            // ---- 4x4 stride-2 conv read in place through TMA element strides (pair == 9; kernel chosen automatically)
            {"s2_direct_small", 2, 32, 32, 64, 128, 4, 2, true, false, true, 0, 9},
            {"s2_direct_w256", 1, 8, 256, 64, 64, 4, 2, false, false, false, 0, 9},
            {"s2_direct_pair", 8, 64, 64, 64, 256, 4, 2, true, false, false, 0, 9},
            {"s2_direct_8x8", 3, 16, 16, 128, 128, 4, 2, true, false, false, 0, 9},
            // ---- 3x3 halo kernel with swapped operands (pair == 8)
            {"t_c3_32x32", 1, 32, 32, 128, 128, 3, 1, true, false, false, 0, 8},
            {"t_c3_32x8_k64", 3, 32, 8, 64, 128, 3, 1, false, false, true, 0, 8},
            {"t_c3_persist", 8, 64, 64, 64, 128, 3, 1, true, true, true, 0, 8},
            {"t_c3_n256", 2, 32, 32, 128, 256, 3, 1, true, true, false, 0, 8},
            {"t_c3_w256", 1, 32, 256, 64, 128, 3, 1, false, false, false, 0, 8},
            {"t16_c3", 3, 16, 16, 128, 128, 3, 1, true, true, true, 0, 8},
            {"lin_1x1", 2, 16, 16, 256, 256, 1, 1, true, true, true, 0, 8},
            {"lin_1x1_ragged", 5, 8, 8, 128, 128, 1, 1, true, false, true, 0, 8},
            {"lin_1x1_persist", 40, 32, 32, 64, 256, 1, 1, false, true, false, 0, 8},
            {"v15_stem_t", 2, 64, 32, 128, 128, 15, 1, true, false, true, 0, 8},
            {"v15_stem_pair", 2, 64, 32, 128, 128, 15, 1, true, false, true, 0, 2},
            {"t16_c3_h32_n256", 2, 32, 16, 64, 256, 3, 1, true, false, false, 0, 8},
            {"t16_c3_persist", 160, 16, 16, 64, 128, 3, 1, false, true, false, 0, 8},
            // ---- CTA-pair (cta_group::2) kernel
            {"p_c3_16x16_n128", 2, 16, 16, 64, 128, 3, 1, true, false, false, 128, 2},
            {"p_c3_16x16_n256", 2, 16, 16, 128, 256, 3, 1, true, true, true, 256, 2},
            {"p_c3_odd_tiles", 5, 8, 8, 64, 128, 3, 1, true, true, false, 128, 2},
            {"p_c3_n512", 2, 32, 32, 64, 512, 3, 1, true, false, false, 256, 2},
            {"p_gemm1x1", 1, 16, 16, 256, 256, 1, 1, false, false, false, 256, 2},
            {"p_c4_s2", 2, 32, 32, 64, 128, 4, 2, true, false, true, 128, 2},
            {"p_c3_persist", 8, 64, 64, 64, 128, 3, 1, true, true, false, 128, 2},
            {"p_c3_deepk", 2, 16, 16, 1024, 256, 3, 1, false, true, false, 256, 2},
            {"p_c3_w256", 1, 4, 256, 64, 128, 3, 1, false, false, false, 128, 2},
            // ---- CTA-pair kernel, stream-K forced (tiles > clusters, not a multiple)
            {"sk_c3_n256_kc1", 5, 64, 64, 64, 256, 3, 1, true, true, true, 256, 5},
            {"sk_c3_n256_kc2", 5, 64, 64, 128, 256, 3, 1, true, true, false, 256, 5},
            {"sk_c3_n128", 5, 64, 64, 128, 128, 3, 1, true, false, true, 128, 5},
            {"sk_c3_n512", 3, 64, 64, 64, 512, 3, 1, false, true, false, 256, 5},
            {"sk_gemm1x1", 10, 64, 64, 256, 256, 1, 1, true, false, false, 256, 5},
            {"sk_c4_s2", 5, 128, 128, 64, 128, 4, 2, true, false, true, 128, 5},
        };
        for (const Case& c : cases) {
            double ms = 0;
            double err = run_case(c, true, 0, &ms);
            if (err < 0) ++failures;
        }
    }
    if (!strcmp(mode, "perf") || !strcmp(mode, "all")) {
        const Case cases[] = {
            {"D1(noTMA) sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 11},
            {"D2(noEpi) sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 12},
            {"D3(neither) sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 13},
            {"D1(noTMA) sr_16_1024 n128", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 128, 11},
            {"D3(neither) sr_16_1024 n128", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 128, 13},
            {"D3alt(2 chains) sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 21},
            {"D3alt(2 chains) sr_16_1024 n128", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 128, 21},
            {"D3alt(2 chains) sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 21},
            {"D3alt(2 chains) sr_16_1024 n64", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 64, 21},
            {"D3(neither) sr_16_1024 n64", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 64, 13},
            {"D1(noTMA) sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 11},
            {"D2(noEpi) sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 12},
            {"D3(neither) sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 13},
            {"S2 direct 256->128 128ch", 32, 256, 256, 128, 128, 4, 2, true, false, false, 0, 9},
            {"S2 split  256->128 128ch", 32, 256, 256, 128, 128, 4, 2, true, false, false, 0, 0},
            {"S2 direct 64->32 256->512", 32, 64, 64, 256, 512, 4, 2, true, false, false, 0, 9},
            {"S2 split  64->32 256->512", 32, 64, 64, 256, 512, 4, 2, true, false, false, 0, 0},
            {"T pw_16_2048", 32, 16, 16, 2048, 1024, 1, 1, true, false, false, 256, 8},
            {"T pw_32_1024", 32, 32, 32, 1024, 512, 1, 1, true, false, false, 256, 8},
            {"T pw_64_512", 32, 64, 64, 512, 256, 1, 1, true, false, false, 256, 8},
            {"T pw_128_256", 32, 128, 128, 256, 128, 1, 1, true, false, false, 128, 8},
            {"pw_32_1024", 32, 32, 32, 1024, 512, 1, 1, true, false, false, 256, 1},
            {"pw_64_512", 32, 64, 64, 512, 256, 1, 1, true, false, false, 256, 1},
            {"pw_128_256", 32, 128, 128, 256, 128, 1, 1, true, false, false, 128, 1},
            {"T stem_256 15x1", 32, 256, 256, 128, 128, 15, 1, true, false, true, 128, 8},
            {"P stem_256 15x1", 32, 256, 256, 128, 128, 15, 1, true, false, true, 128, 2},
            {"T sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 8},
            {"T sr_128_128_f16", 32, 128, 128, 128, 128, 3, 1, true, false, true, 128, 8},
            {"T sr_128_256to128", 32, 128, 128, 256, 128, 3, 1, true, true, false, 128, 8},
            {"T sr_256_128", 16, 256, 256, 128, 128, 3, 1, true, true, false, 128, 8},
            {"T sr_64_256", 32, 64, 64, 256, 256, 3, 1, true, true, false, 256, 8},
            {"T16 sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 8},
            {"T16 sr_16_2048", 32, 16, 16, 2048, 1024, 3, 1, true, false, false, 256, 8},
            {"T sr_64_512to256", 32, 64, 64, 512, 256, 3, 1, true, false, false, 256, 8},
            {"T sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 8},
            {"T sr_32_1024", 32, 32, 32, 1024, 512, 3, 1, true, false, false, 256, 8},
            {"P sr_64_512to256", 32, 64, 64, 512, 256, 3, 1, true, false, false, 256, 2},
            {"H final_256_128_16", 32, 256, 256, 128, 16, 3, 1, true, false, false, 0, 3},
            {"H sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 3},
            {"H sr_16_2048", 32, 16, 16, 2048, 1024, 3, 1, true, true, false, 256, 3},
            {"H sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 3},
            {"H sr_64_256", 32, 64, 64, 256, 256, 3, 1, true, true, false, 256, 3},
            {"H sr_64_256 n128", 32, 64, 64, 256, 256, 3, 1, true, true, false, 128, 3},
            {"H sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 3},
            {"H sr_128_128_f16", 32, 128, 128, 128, 128, 3, 1, true, false, true, 128, 3},
            {"H sr_256_128", 16, 256, 256, 128, 128, 3, 1, true, true, false, 128, 3},
            {"P(kc1) sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 4},
            {"P(kc1) sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 4},
            {"P(kc1) sr_64_256", 32, 64, 64, 256, 256, 3, 1, true, true, false, 256, 4},
            {"P(kc1) sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 4},
            {"Pnosk 148tiles_16_1024", 37, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 6},
            {"Pnosk 74tiles_16_1024", 37, 16, 16, 1024, 512, 3, 1, true, true, false, 256, 6},
            {"Pnosk 64tiles_16_1024", 32, 16, 16, 1024, 512, 3, 1, true, true, false, 256, 6},
            {"SKnoexch sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 7},
            {"SKnoexch sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 7},
            {"Pnosk sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 6},
            {"Pnosk sr_16_2048", 32, 16, 16, 2048, 1024, 3, 1, true, true, false, 256, 6},
            {"Pnosk sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 6},
            {"Pnosk sr_32_1024", 32, 32, 32, 1024, 512, 3, 1, true, true, false, 256, 6},
            {"Pnosk pw_16_2048", 32, 16, 16, 2048, 1024, 1, 1, true, false, false, 256, 6},
            {"P sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 2},
            {"P sr_16_2048", 32, 16, 16, 2048, 1024, 3, 1, true, true, false, 256, 2},
            {"P sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 2},
            {"P sr_32_1024", 32, 32, 32, 1024, 512, 3, 1, true, true, false, 256, 2},
            {"P sr_64_256", 32, 64, 64, 256, 256, 3, 1, true, true, false, 256, 2},
            {"P sr_64_256 n128", 32, 64, 64, 256, 256, 3, 1, true, true, false, 128, 2},
            {"P sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 2},
            {"P sr_128_128_f16", 32, 128, 128, 128, 128, 3, 1, true, false, true, 128, 2},
            {"P sr_256_128", 16, 256, 256, 128, 128, 3, 1, true, true, false, 128, 2},
            {"P pw_16_2048", 32, 16, 16, 2048, 1024, 1, 1, true, false, false, 256, 2},
            {"sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 256, 1},
            {"sr_16_1024", 32, 16, 16, 1024, 1024, 3, 1, true, true, false, 128, 1},
            {"sr_16_2048", 32, 16, 16, 2048, 1024, 3, 1, true, true, false, 256, 1},
            {"sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 256, 1},
            {"sr_32_512", 32, 32, 32, 512, 512, 3, 1, true, true, false, 128, 1},
            {"sr_64_256", 32, 64, 64, 256, 256, 3, 1, true, true, false, 256, 1},
            {"sr_64_256", 32, 64, 64, 256, 256, 3, 1, true, true, false, 128, 1},
            {"sr_128_128", 32, 128, 128, 128, 128, 3, 1, true, true, false, 128, 1},
            {"sr_128_128_f16", 32, 128, 128, 128, 128, 3, 1, true, false, true, 128, 1},
            {"sr_256_128", 16, 256, 256, 128, 128, 3, 1, true, true, false, 128, 1},
            {"pw_16_2048", 32, 16, 16, 2048, 1024, 1, 1, true, false, false, 256, 1},
        };
        for (const Case& c : cases) {
            double ms = 0;
            run_case(c, false, 10, &ms);
            const double flops = 2.0 * c.B * (c.H / c.stride) * (c.W / c.stride) * (double)c.Cout * (c.k == 15 ? 15 : c.k * c.k) * c.Cin;
            printf("[perf %s hint=%d pair=%d] B=%d %dx%d %d->%d k=%d : %.3f ms  %.1f TFLOP/s\n", c.name, c.hint, c.pair, c.B, c.H, c.W,
                   c.Cin, c.Cout, c.k, ms, flops / ms * 1e-9);
            fflush(stdout);
        }
    }
    printf("done, launch failures=%d\n", failures);
    return failures ? 1 : 0;
}
