// Shared epilogue of the tcgen05 conv kernels (conv_tc.cu, conv_gn.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "conv_tc.cuh"
#include "ptx.cuh"
#include "sat_half.cuh"

namespace mi {
namespace {

constexpr int kEpiWarps = 8;                                   // two warps per TMEM lane quarter, each owning half the columns
constexpr uint32_t kEpiBytes = kEpiWarps * 32 * 32 * 4;        // epilogue transpose patches (4 KB per warp)

// TMEM accumulator tile -> +bias -> +residual -> global stores.
// tcgen05.ld hands every thread one accumulator ROW; storing rows directly makes each warp-wide store touch 32 different
// lines with 16 B each.  So each warp transposes 32 x 32 slabs through a private shared-memory patch: rows go in, and
// come back out as (4 rows x 8 lanes x float4) so that every warp-wide load of the residual and store of the result
// covers four full 128-byte lines.  All eight residual loads of a slab are issued before the accumulator is read, to
// overlap their latency.  Strided-channel / ragged-N outputs (final_conv -> NCHW) keep the simple per-row path.
constexpr int kEpiLd = 32;   // floats per staged row; 16-byte chunks are XOR-swizzled with the row index (conflict-free)

template <int BLOCK_N>
__device__ __forceinline__ void epilogue_tile(const ConvTcArgs& args, uint32_t taddr, int n0, long long pix, bool valid,
                                              float* stage /* this warp's [32][kEpiLd] patch */, int c_begin, int c_end,
                                              int b_img) {
    const int lane = threadIdx.x & 31;
    if (args.out_sc != 1 || n0 + BLOCK_N > args.n_valid || (BLOCK_N % 32) != 0) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
            uint32_t v[16];
            ptx::tmem_ld_x16(taddr + c, v);
            ptx::tmem_ld_wait();
            if (valid) {
                const int n = n0 + c;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (n + i < args.n_valid) {
                        float f = __uint_as_float(v[i]) + (args.bias ? __ldg(args.bias + n + i) : 0.f);
                        const long long o = pix + (long long)(n + i) * args.out_sc;
                        if (args.residual && args.out_sc == 1) f += args.residual[o];
                        if (args.out_f32) args.out_f32[o] = f;
                        if (args.out_f16) args.out_f16[o] = sat_half(f);
                    }
                }
            }
        }
        return;
    }
    const int sub = lane >> 3;            // which of the 4 rows a quarter-warp handles per step
    const int cv = (lane & 7) * 4;        // its 4 columns inside the 32-column slab
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const int pix_lo = (int)(pix & 0xffffffffLL), pix_hi = (int)(pix >> 32);
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
        const int n = n0 + c;
        float4 res[8];
        if (args.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + sub;
                const long long p = ((long long)__shfl_sync(0xffffffffu, pix_hi, r) << 32) |
                                    (unsigned)__shfl_sync(0xffffffffu, pix_lo, r);
                res[i] = ((vmask >> r) & 1u) ? *reinterpret_cast<const float4*>(args.residual + p + n + cv)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        uint32_t v0[16], v1[16];
        ptx::tmem_ld_x16(taddr + c, v0);
        ptx::tmem_ld_x16(taddr + c + 16, v1);
        ptx::tmem_ld_wait();
        float* srow = stage + lane * kEpiLd;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (args.bias) {
                b0 = __ldg(reinterpret_cast<const float4*>(args.bias + n + i));
                b1 = __ldg(reinterpret_cast<const float4*>(args.bias + n + 16 + i));
            }
            *reinterpret_cast<float4*>(srow + ((((i >> 2)) ^ (lane & 7)) << 2)) =
                make_float4(__uint_as_float(v0[i]) + b0.x, __uint_as_float(v0[i + 1]) + b0.y,
                            __uint_as_float(v0[i + 2]) + b0.z, __uint_as_float(v0[i + 3]) + b0.w);
            *reinterpret_cast<float4*>(srow + ((((i >> 2) + 4) ^ (lane & 7)) << 2)) =
                make_float4(__uint_as_float(v1[i]) + b1.x, __uint_as_float(v1[i + 1]) + b1.y,
                            __uint_as_float(v1[i + 2]) + b1.z, __uint_as_float(v1[i + 3]) + b1.w);
        }
        __syncwarp();
        float st_s = 0.f, st_q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + sub;
            float4 f = *reinterpret_cast<const float4*>(stage + r * kEpiLd + (((lane & 7) ^ (r & 7)) << 2));
            const long long p = ((long long)__shfl_sync(0xffffffffu, pix_hi, r) << 32) |
                                (unsigned)__shfl_sync(0xffffffffu, pix_lo, r);
            if ((vmask >> r) & 1u) {
                if (args.residual) { f.x += res[i].x; f.y += res[i].y; f.z += res[i].z; f.w += res[i].w; }
                st_s += (f.x + f.y) + (f.z + f.w);
                st_q += (f.x * f.x + f.y * f.y) + (f.z * f.z + f.w * f.w);
                if (args.out_f32) *reinterpret_cast<float4*>(args.out_f32 + p + n + cv) = f;
                if (args.out_f16) {
                    __half2 lo = sat_half2(f.x, f.y), hi = sat_half2(f.z, f.w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(args.out_f16 + p + n + cv) = pk;
                }
            }
        }
        if (args.stats) {
            // GroupNorm statistics of the tensor being written, per (image, 16-channel block): this warp's 32 rows x 32
            // columns reduce to two (sum, sum of squares) pairs -- lanes with bit 2 clear hold block 0, set: block 1
#pragma unroll
            for (int o = 1; o <= 16; o <<= 1) {
                if (o == 4) continue;
                st_s += __shfl_xor_sync(0xffffffffu, st_s, o);
                st_q += __shfl_xor_sync(0xffffffffu, st_q, o);
            }
            const int b0 = __shfl_sync(0xffffffffu, b_img, vmask ? (__ffs(vmask) - 1) : 0);
            if ((lane == 0 || lane == 4) && vmask != 0u) {
                double* dst = args.stats + ((long long)b0 * args.stats_blocks + (n >> 4) + (lane >> 2)) * 2;
                atomicAdd(dst, (double)st_s);
                atomicAdd(dst + 1, (double)st_q);
            }
        }
        __syncwarp();
    }
}

}  // namespace
}  // namespace mi
