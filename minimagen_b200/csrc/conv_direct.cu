// Direct fp32 convolution on CUDA cores, for the shapes the tensor-core implicit GEMM does not take:
//   * CrossEmbedLayer stem convs on 3/6 input channels, k = 3/7/15   (minimagen/layers.py:300, Unet.py:169-172)
//   * final_conv  dim -> 3 channels                                    (minimagen/Unet.py:327)
//   * every conv of the tiny test config (channels 8/16/24/32)        (cfg 1)
// fp32 NHWC input (channel-contiguous, row pitch `ldi`), OIHW fp32 weights exactly as the reference stores them,
// fp32 accumulate, arbitrary output strides (so the result can land in a channel slice of an NHWC buffer or in NCHW).
// One thread = one output pixel x 8 output channels; weights for the CTA's 8 channels are staged in shared memory
// (input-channel chunks), inputs come through L1 (neighbouring pixels overlap heavily).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "launch.cuh"

namespace mi {

namespace {

constexpr int kCoT = 8;

__global__ void __launch_bounds__(256)
conv_direct_kernel(const float* __restrict__ in, int B, int Hin, int Win, int Cin, int ldi,
                   const float* __restrict__ w, int Cout, int KH, int KW, int stride, int pad,
                   const float* __restrict__ bias, const float* __restrict__ residual, float* __restrict__ out,
                   int Hout, int Wout, long long osb, long long osh, long long osw, long long osc, int ci_chunk) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float w_s[];   // [KH*KW][ci_chunk][kCoT]
    const int co0 = blockIdx.y * kCoT;
    const int taps = KH * KW;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = pix < (long long)B * Hout * Wout;
    const int wo = active ? (int)(pix % Wout) : 0;
    const int ho = active ? (int)((pix / Wout) % Hout) : 0;
    const int b = active ? (int)(pix / ((long long)Wout * Hout)) : 0;

    float acc[kCoT];
#pragma unroll
    for (int i = 0; i < kCoT; ++i) acc[i] = 0.f;

    const int Cin4 = (Cin + 3) & ~3;
    for (int c0 = 0; c0 < Cin4; c0 += ci_chunk) {
        const int cc = min(ci_chunk, Cin4 - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < taps * cc * kCoT; i += blockDim.x) {
            const int co = i % kCoT;
            const int ci = (i / kCoT) % cc;
            const int t = i / (kCoT * cc);
            float v = 0.f;
            if (co0 + co < Cout && c0 + ci < Cin) v = w[((long long)(co0 + co) * Cin + (c0 + ci)) * taps + t];
            w_s[i] = v;
        }
        __syncthreads();
        if (active) {
            for (int r = 0; r < KH; ++r) {
                const int hi = ho * stride + r - pad;
                if (hi < 0 || hi >= Hin) continue;
                for (int s = 0; s < KW; ++s) {
                    const int wi = wo * stride + s - pad;
                    if (wi < 0 || wi >= Win) continue;
                    const float* ip = in + (((long long)b * Hin + hi) * Win + wi) * ldi + c0;
                    const float* wp = w_s + (r * KW + s) * cc * kCoT;
                    for (int ci = 0; ci < cc; ci += 4) {
                        const float4 x = *reinterpret_cast<const float4*>(ip + ci);
                        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 w0 = *reinterpret_cast<const float4*>(wp + (ci + e) * kCoT);
                            const float4 w1 = *reinterpret_cast<const float4*>(wp + (ci + e) * kCoT + 4);
                            acc[0] += xv[e] * w0.x; acc[1] += xv[e] * w0.y; acc[2] += xv[e] * w0.z; acc[3] += xv[e] * w0.w;
                            acc[4] += xv[e] * w1.x; acc[5] += xv[e] * w1.y; acc[6] += xv[e] * w1.z; acc[7] += xv[e] * w1.w;
                        }
                    }
                }
            }
        }
    }
    if (!active) return;
    const long long base = (long long)b * osb + (long long)ho * osh + (long long)wo * osw;
#pragma unroll
    for (int co = 0; co < kCoT; ++co) {
        if (co0 + co < Cout) {
            const long long off = base + (long long)(co0 + co) * osc;
            float v = acc[co];
            if (bias) v += bias[co0 + co];
            if (residual) v += residual[off];
            out[off] = v;
        }
    }
}

}  // namespace

int conv_direct_f32(const float* in, int B, int Hin, int Win, int Cin, int ldi, const float* w, int Cout, int KH,
                    int KW, int stride, int pad, const float* bias, const float* residual, float* out, int Hout,
                    int Wout, long long osb, long long osh, long long osw, long long osc, cudaStream_t st) {
    if (ldi % 4 || ((Cin + 3) & ~3) > ldi) return -1;
    if (reinterpret_cast<uintptr_t>(in) & 15) return -1;
    const int taps = KH * KW;
    const int Cin4 = (Cin + 3) & ~3;
    // stage as many input channels per pass as fit in ~96 KiB of shared memory
    int ci_chunk = (96 * 1024) / (taps * kCoT * 4);
    ci_chunk &= ~3;
    if (ci_chunk < 4) return -3;
    if (ci_chunk > Cin4) ci_chunk = Cin4;
    const size_t smem = (size_t)taps * ci_chunk * kCoT * 4;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
            return -4;
        attr_set = true;
    }
    const long long npix = (long long)B * Hout * Wout;
    dim3 grid((unsigned)((npix + 255) / 256), (Cout + kCoT - 1) / kCoT);
    launch_k(conv_direct_kernel, grid, 256, smem, st, in, B, Hin, Win, Cin, ldi, w, Cout, KH, KW, stride, pad, bias, residual,
                                                out, Hout, Wout, osb, osh, osw, osc, ci_chunk);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace mi
