// Fused Block.forward (minimagen/layers.py:131-145):  GroupNorm -> (scale + 1, shift) -> SiLU -> Conv2d 3x3 in ONE kernel;
// the normalised / activated tensor never exists in HBM.
//
// This is conv3x3_halo_t_kernel<G32x8> (conv_tc.cu: swapped operands, D^T[128 channels][256 pixels] = W_tile x window^T, the
// nine taps as descriptor windows into one (32+2) x (8+2)-pixel halo tile per 64-channel chunk) with the TMA load of the
// activation halo replaced by a PROLOGUE executed by twelve dedicated warps:
//
//   raw input    the fp32 NHWC residual-stream tensor(s) (optionally the virtual concat cat(x, skip * s), Unet.py:445), read
//                straight from global memory / L2 into registers with coalesced 128-bit loads (one pixel's 64 channels =
//                256 contiguous bytes = eight lanes) -- no fp32 staging buffer in shared memory, so the shared-memory
//                traffic of the kernel is the same as the un-fused conv's (the MMA operand reads already use most of the
//                128 B/clk port);
//   transform    y = SiLU(x * A[b,c] + Bc[b,c]); the per-(image, channel) coefficients fold the GroupNorm mean / rstd (from
//                the producers' epilogue block statistics), its affine, the FiLM (scale + 1, shift) and the skip-connection
//                scale; halo pixels outside the image are written as 0 (the conv zero-pads the ACTIVATED tensor);
//   operand      written as fp16 directly in the 128-byte-swizzled K-major layout tcgen05.mma consumes (16-byte chunk q of
//                halo pixel p lives at p*128 + ((q ^ (p & 7)) << 4)), fence.proxy.async + mbarrier hand-off to the issuer;
//   main loop / epilogue   as conv3x3_halo_t_kernel: weights [C_out][9*C_in] by TMA, accumulators double-buffered in TMEM,
//                lane = channel epilogue with bias / fp32 residual / fp32 + fp16 stores / GroupNorm block statistics.
//
// Warp roles (768 threads): 0 TMA (weights), 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue, 12-23 transform.
#include "conv_tc.cuh"

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "sat_half.cuh"

namespace mi {

namespace {

constexpr int kGnThreads = 768;
constexpr int kXformWarp0 = 12, kXformThreads = 384, kGnEpiWarps = 8;
constexpr int kXformRows = kXformThreads / 8;                            // halo pixels covered per iteration (48: a multiple of 8)
constexpr int kTH = 32, kTW = 8, kBoxW = kTW + 2, kBoxH = kTH + 2, kHaloPix = kBoxH * kBoxW;   // 340 halo pixels
constexpr uint32_t kHaloBytes = kHaloPix * 128;                          // fp16 operand tile: 43520 B
constexpr uint32_t kHaloStride = (kHaloBytes + 1023) & ~1023u;           // 44032
constexpr uint32_t kWBytes = 128 * kConvBlockK * 2;                      // one (tap, chunk) weight tile: 16 KiB
constexpr int kHStages = 2, kWStages = 6;
constexpr int kPix = 256;                                                // UMMA N
constexpr uint32_t kTmemCols = 2 * kPix;
constexpr uint32_t kAuxBytes = 512;                                      // barriers + group mean / rstd
constexpr uint32_t kSmemBase = kHStages * kHaloStride + kWStages * kWBytes + 1024 + kAuxBytes;   // + 8 * C_in (coefficient table)
constexpr uint32_t kSmemMax = 227 * 1024;
constexpr int kItems = kHaloPix * 8;                                     // 16-byte operand chunks per halo tile
constexpr int kIters = (kItems + kXformThreads - 1) / kXformThreads;     // 8
constexpr int kBatch = 4;                                                // loads in flight per thread: kBatch x 32 B

__device__ __forceinline__ uint64_t make_win_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;        // stride between 8-pixel row segments
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ void xform_bar_sync() {           // the twelve transform warps only
    asm volatile("bar.sync 1, 384;" ::: "memory");
}

__global__ void __launch_bounds__(kGnThreads, 1)
conv3x3_gn_t_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvTcArgs args,
                    const __grid_constant__ GnPrologueArgs gn) {
    constexpr int NH = kHStages, NW = kWStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_w = smem + NH * kHaloStride;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + NW * kWBytes);
    uint64_t* fullH = bars;
    uint64_t* emptyH = bars + NH;
    uint64_t* fullW = bars + 2 * NH;
    uint64_t* emptyW = bars + 2 * NH + NW;
    uint64_t* tfull_bar = bars + 2 * NH + 2 * NW;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    static_assert((2 * NH + 2 * NW + 4) * 8 + 8 <= 256, "barrier block too large");
    float* s_mean = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);     // [32]
    float* s_rstd = s_mean + 32;                                                          // [32]
    float* s_coef = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);     // A[C_in] then Bc[C_in] of the current image

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int* err = args.err_flag;

    if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tmB);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NH; ++i) { ptx::mbar_init(&fullH[i], kXformThreads / 32); ptx::mbar_init(&emptyH[i], 1); }
        for (int i = 0; i < NW; ++i) { ptx::mbar_init(&fullW[i], 1); ptx::mbar_init(&emptyW[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], 32 * kGnEpiWarps); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();      // everything above is independent of the previous kernel's output

    const int chunks = args.chunks_per_tap;
    const int tiles_m = args.tiles_w * args.tiles_h * args.tiles_b;
    const int total_tiles = tiles_m * args.tiles_n;
    const int Cin = chunks * kConvBlockK;

    if (warp == 0) {
        // ===================== TMA producer: weight tiles only =====================
        int sw = 0;
        uint32_t pw = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n0 = (tile % args.tiles_n) * 128;
            for (int j = 0; j < chunks; ++j) {
                for (int t = 0; t < 9; ++t) {
                    ptx::mbar_wait(&emptyW[sw], pw ^ 1, err, 5200 + sw);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&fullW[sw], kWBytes);
                        ptx::tma_load_2d(&tmB, &fullW[sw], smem_w + sw * kWBytes, t * Cin + j * kConvBlockK, n0);
                    }
                    if (++sw == NW) { sw = 0; pw ^= 1; }
                }
            }
        }
        pdl_trigger();      // last weight loads issued: the next kernel may be scheduled behind this one's final tile(s)
    } else if (warp == 1) {
        // ===================== MMA issuer: D^T[128 ch][256 px] += W_tile[128][64] x window^T =====================
        constexpr uint32_t idesc = ptx::make_idesc_f16(128, kPix, 0);
        int sh = 0, sw = 0;
        uint32_t ph = 0, pw = 0;
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, err, 5300 + as);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * kPix;
            for (int j = 0; j < chunks; ++j) {
                ptx::mbar_wait(&fullH[sh], ph, err, 5400 + sh);          // the transformed halo tile is in shared memory
                const uint32_t h_base = ptx::smem_u32(smem + sh * kHaloStride);
                for (int t = 0; t < 9; ++t) {
                    ptx::mbar_wait(&fullW[sw], pw, err, 5500 + sw);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_w + sw * kWBytes));
                        // tap t = dh*3 + dw starts (dh*10 + dw) pixels into the halo tile; 32 row segments one halo row apart
                        const uint64_t db = make_win_desc(h_base + ((t / 3) * kBoxW + (t % 3)) * 128, kBoxW * 128);
#pragma unroll
                        for (int k = 0; k < kConvBlockK / 16; ++k)
                            ptx::umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (j | t | k) != 0);
                        ptx::umma_commit(&emptyW[sw]);
                    }
                    if (++sw == NW) { sw = 0; pw ^= 1; }
                }
                if (ptx::elect_one()) ptx::umma_commit(&emptyH[sh]);
                if (++sh == NH) { sh = 0; ph ^= 1; }
            }
            if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[as]);
        }
    } else if (warp >= kXformWarp0) {
        // ===================== transform: fp32 global -> GroupNorm/FiLM/SiLU -> fp16 swizzled halo operand =====================
        const int tt = threadIdx.x - kXformWarp0 * 32;      // 0..383
        const int lq = tt & 7;                              // logical 16-byte chunk: channels [8*lq, 8*lq + 8) of the k-chunk
        const int prow = tt >> 3;                           // halo pixel of iteration it: p = it*48 + prow; p & 7 == prow & 7
        const int qo = lq ^ (prow & 7);                     // physical (swizzled) chunk position inside the 128-byte row
        const int C0 = gn.C0, C1 = gn.C1, Ctot = C0 + C1, Cg = Ctot / gn.groups;
        const int H = args.H, W = args.W;
        float* sA = s_coef;
        float* sB = s_coef + Ctot;
        int sh = 0;
        uint32_t ph = 0;
        int cur_b = -1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / args.tiles_n;
            const int w0 = (mt % args.tiles_w) * kTW;
            const int h0 = ((mt / args.tiles_w) % args.tiles_h) * kTH;
            const int b = mt / (args.tiles_w * args.tiles_h);
            if (b != cur_b) {
                // Coefficient table of image b for ALL input channels: y = SiLU(x * A[c] + Bc[c]) (same arithmetic as
                // gn_apply_silu_kernel).  Rebuilt only when the image changes; no per-chunk barrier or global load afterwards.
                xform_bar_sync();                            // every thread has finished reading the previous table
                if (tt < gn.groups) {
                    const int g = tt;
                    double su = 0.0, sq = 0.0;
                    const int lo = g * Cg, hi = lo + Cg;
                    const int lo0 = min(lo, C0), hi0 = min(hi, C0);
                    for (int e = lo0 / 16; e < hi0 / 16; ++e) {
                        su += gn.stats0[((long long)b * (C0 / 16) + e) * 2];
                        sq += gn.stats0[((long long)b * (C0 / 16) + e) * 2 + 1];
                    }
                    const int lo1 = max(lo, C0) - C0, hi1 = max(hi, C0) - C0;
                    for (int e = lo1 / 16; e < hi1 / 16; ++e) {
                        su += (double)gn.scale1 * gn.stats1[((long long)b * (C1 / 16) + e) * 2];
                        sq += (double)gn.scale1 * (double)gn.scale1 * gn.stats1[((long long)b * (C1 / 16) + e) * 2 + 1];
                    }
                    const double n = (double)Cg * H * W;
                    const double mean = su / n;
                    double var = sq / n - mean * mean;
                    if (var < 0) var = 0;
                    s_mean[g] = (float)mean;
                    s_rstd[g] = (float)(1.0 / sqrt(var + (double)gn.eps));
                }
                xform_bar_sync();
                for (int cc = tt; cc < Ctot; cc += kXformThreads) {
                    const int g = cc / Cg;
                    float a = s_rstd[g] * gn.gamma[cc];
                    float bb = gn.beta[cc] - s_mean[g] * a;
                    if (gn.scale_shift) {
                        const float sc = gn.scale_shift[(long long)b * gn.ss_ld + cc] + 1.0f;
                        const float shv = gn.scale_shift[(long long)b * gn.ss_ld + Ctot + cc];
                        a *= sc;
                        bb = bb * sc + shv;
                    }
                    if (cc >= C0) a *= gn.scale1;                        // skip * 2^-1/2 folded into the multiplier
                    sA[cc] = a;
                    sB[cc] = bb;
                }
                cur_b = b;
                xform_bar_sync();
            }
            const long long img = (long long)b * H * W;
            for (int j = 0; j < chunks; ++j) {
                const float* cA = sA + j * kConvBlockK + lq * 8;
                const float* cB = sB + j * kConvBlockK + lq * 8;
                const float4 a0 = *reinterpret_cast<const float4*>(cA), a1 = *reinterpret_cast<const float4*>(cA + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(cB), b1 = *reinterpret_cast<const float4*>(cB + 4);

                const bool first = j < args.a_split;
                const int Cs = first ? C0 : C1;
                const float* src = (first ? gn.src0 + (long long)j * kConvBlockK
                                          : gn.src1 + (long long)(j - args.a_split) * kConvBlockK) + img * Cs + lq * 8;

                ptx::mbar_wait(&emptyH[sh], ph ^ 1, err, 5600 + sh);     // operand slot free (its MMAs retired)
                uint8_t* op = smem + sh * kHaloStride + qo * 16;
#pragma unroll 1
                for (int it0 = 0; it0 < kIters; it0 += kBatch) {
                    float4 x0[kBatch], x1[kBatch];
                    bool ok[kBatch];
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int p = (it0 + u) * kXformRows + prow;
                        const int hr = p / kBoxW, hc = p - hr * kBoxW;
                        const int gh = h0 - 1 + hr, gw = w0 - 1 + hc;
                        ok[u] = (p < kHaloPix) && gh >= 0 && gh < H && gw >= 0 && gw < W;
                        if (ok[u]) {
                            const float4* s4 = reinterpret_cast<const float4*>(src + ((long long)gh * W + gw) * Cs);
                            x0[u] = __ldg(s4);
                            x1[u] = __ldg(s4 + 1);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int p = (it0 + u) * kXformRows + prow;
                        if (p >= kHaloPix) continue;
                        uint4 o = make_uint4(0u, 0u, 0u, 0u);
                        if (ok[u]) {
                            float v[8] = {fmaf(x0[u].x, a0.x, b0.x), fmaf(x0[u].y, a0.y, b0.y), fmaf(x0[u].z, a0.z, b0.z),
                                          fmaf(x0[u].w, a0.w, b0.w), fmaf(x1[u].x, a1.x, b1.x), fmaf(x1[u].y, a1.y, b1.y),
                                          fmaf(x1[u].z, a1.z, b1.z), fmaf(x1[u].w, a1.w, b1.w)};
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = __fdividef(v[e], 1.0f + __expf(-v[e]));
                            const __half2 h0_ = sat_half2(v[0], v[1]), h1_ = sat_half2(v[2], v[3]);
                            const __half2 h2_ = sat_half2(v[4], v[5]), h3_ = sat_half2(v[6], v[7]);
                            o.x = *reinterpret_cast<const uint32_t*>(&h0_); o.y = *reinterpret_cast<const uint32_t*>(&h1_);
                            o.z = *reinterpret_cast<const uint32_t*>(&h2_); o.w = *reinterpret_cast<const uint32_t*>(&h3_);
                        }
                        *reinterpret_cast<uint4*>(op + p * 128) = o;
                    }
                }
                ptx::fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();                         // one arrival per warp: 384 serialised arrivals per stage cost more than the math
                if (lane == 0) ptx::mbar_arrive(&fullH[sh]);
                if (++sh == NH) { sh = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: lane = channel, columns = pixels (16 pixels per step) =====================
        const int q = warp & 3;                       // TMEM lane quarter -> channels [32q, 32q + 32) of the tile
        const int half = warp >= 8 ? 1 : 0;           // pixel columns [128*half, 128*half + 128) = tile rows [16*half, +16)
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int nt = tile % args.tiles_n;
            const int mt = tile / args.tiles_n;
            const int w0 = (mt % args.tiles_w) * kTW;
            const int h0 = ((mt / args.tiles_w) % args.tiles_h) * kTH;
            const int b = mt / (args.tiles_w * args.tiles_h);
            const int n = nt * 128 + q * 32 + lane;   // this thread's output channel
            const float bias_v = args.bias ? __ldg(args.bias + n) : 0.f;
            const long long base = (long long)b * args.out_sb + (long long)(h0 + half * (kTH / 2)) * args.out_sh +
                                   (long long)w0 * args.out_sw + n;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            ptx::mbar_wait(&tfull_bar[as], aphase, err, 5700 + as);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kPix + half * 128;
            float st_s = 0.f, st_q = 0.f;
#pragma unroll 1
            for (int c = 0; c < 128; c += 16) {       // 16 pixels = two tile rows of 8 per step
                uint32_t v[16];
                ptx::tmem_ld_x16(taddr + c, v);
                const long long rowb = base + (long long)(c >> 3) * args.out_sh;
                float r[16];
                if (args.residual) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        r[i] = args.residual[rowb + (long long)(i >> 3) * args.out_sh + (long long)(i & 7) * args.out_sw];
                }
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float f = __uint_as_float(v[i]) + bias_v;
                    if (args.residual) f += r[i];
                    st_s += f;
                    st_q += f * f;
                    const long long o = rowb + (long long)(i >> 3) * args.out_sh + (long long)(i & 7) * args.out_sw;
                    if (args.out_f32) args.out_f32[o] = f;
                    if (args.out_f16) args.out_f16[o] = sat_half(f);
                }
            }
            if (args.stats) {
                // 16-channel blocks = half warps: lanes 0-15 and 16-31
#pragma unroll
                for (int o = 1; o <= 8; o <<= 1) {
                    st_s += __shfl_xor_sync(0xffffffffu, st_s, o);
                    st_q += __shfl_xor_sync(0xffffffffu, st_q, o);
                }
                if ((lane & 15) == 0) {
                    double* dst = args.stats + ((long long)b * args.stats_blocks + (n >> 4)) * 2;
                    atomicAdd(dst, (double)st_s);
                    atomicAdd(dst + 1, (double)st_q);
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty_bar[as]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace

bool conv_gn_supported(int H, int W, int C0, int C1, int Cout, int groups) {
    const int C = C0 + C1;
    if (H <= 0 || W <= 0 || H % kTH || W % kTW || C0 <= 0 || C0 % 64 || C1 < 0 || C1 % 64 || Cout <= 0 || Cout % 128) return false;
    if (groups < 1 || groups > 32 || C % groups) return false;
    if (kSmemBase + 8u * (uint32_t)C > kSmemMax) return false;          // coefficient table A[C], Bc[C] in shared memory
    return (C / groups) % 16 == 0;
}

int conv_gn_launch(const ConvGnProblem& p, cudaStream_t stream) {
    if (!conv_gn_supported(p.H, p.W, p.C0, p.C1, p.Cout, p.groups)) return -3;
    if (p.C1 && (!p.src1 || !p.stats1)) return -8;
    if (!p.src0 || !p.stats0 || !p.gamma || !p.beta) return -8;
    if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.src1) & 15) ||
        (reinterpret_cast<uintptr_t>(p.wpacked) & 15))
        return -8;
    PFN_tmaEncodeTiled enc = get_tma_encode();
    if (!enc) return -5;
    const int C = p.C0 + p.C1;

    ConvTcArgs a{};
    a.num_taps = 9;
    a.chunks_per_tap = C / kConvBlockK;
    a.a_split = p.C0 / kConvBlockK;
    a.tiles_w = p.W / kTW; a.tiles_h = p.H / kTH; a.tiles_b = p.B; a.tiles_n = p.Cout / 128;
    a.B = p.B; a.H = p.H; a.W = p.W;
    a.out_sb = (long long)p.H * p.W * p.Cout; a.out_sh = (long long)p.W * p.Cout; a.out_sw = p.Cout; a.out_sc = 1;
    a.n_valid = p.Cout;
    a.out_f32 = p.out_f32; a.out_f16 = p.out_f16; a.bias = p.bias; a.residual = p.residual; a.err_flag = p.err_flag;
    a.stats = p.out_stats; a.stats_blocks = p.Cout / 16;

    GnPrologueArgs g{};
    g.src0 = p.src0; g.src1 = p.src1;
    g.C0 = p.C0; g.C1 = p.C1; g.groups = p.groups; g.scale1 = p.scale1; g.eps = p.eps;
    g.stats0 = p.stats0; g.stats1 = p.stats1; g.gamma = p.gamma; g.beta = p.beta;
    g.scale_shift = p.scale_shift; g.ss_ld = p.ss_ld;

    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);

    CUtensorMap tmB;
    const cuuint64_t K = (cuuint64_t)9 * C;
    cuuint64_t wdim[2] = {K, (cuuint64_t)p.Cout};
    cuuint64_t wstr[1] = {K * 2};
    cuuint32_t wbox[2] = {kConvBlockK, 128};
    cuuint32_t westr[2] = {1, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(p.wpacked), wdim, wstr, wbox, westr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -7;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv3x3_gn_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax) != cudaSuccess)
            return -10;
        attr_set = true;
    }
    const int total = a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n;
    const int grid = total < num_sms ? total : num_sms;
    launch_k(conv3x3_gn_t_kernel, grid, kGnThreads, kSmemBase + 8u * (uint32_t)C, stream, tmB, a, g);
    return cudaGetLastError() == cudaSuccess ? 0 : -11;
}

}  // namespace mi
