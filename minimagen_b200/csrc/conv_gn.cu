// Fused Block.forward (minimagen/layers.py:131-145):  GroupNorm -> (scale + 1, shift) -> SiLU -> Conv2d 3x3,
// one kernel, the normalised / activated tensor never touches HBM.
//
//   raw input    fp32 NHWC residual-stream tensor(s) (optionally the virtual concat cat(x, skip * s), Unet.py:445),
//                fetched as (16+2) x (8+2) pixel HALO tiles of 64 channels by TMA (two 32-channel boxes, zero-filled
//                outside the image) into a staging buffer;
//   prologue     4 "transform" warps turn the staged fp32 halo tile into the fp16, 128B-swizzled K-major MMA operand:
//                y = SiLU(x * A[b,c] + Bc[b,c]) with the per-(image, channel) coefficients folding GroupNorm (mean / rstd
//                from the producers' epilogue block statistics), its affine, and the FiLM (scale + 1, shift); halo
//                positions outside the image are forced to 0 (the conv pads the ACTIVATED tensor with zeros);
//   main loop    nine taps = nine tcgen05.mma descriptor windows into the transformed halo tile (see conv_tc.cu,
//                conv3x3_halo_kernel), weights [C_out][9*C_in] streamed per tap by TMA;
//   epilogue     shared with conv_tc.cu: +bias, +fp32 residual, fp32 and/or fp16 stores, GroupNorm block statistics of
//                the output for the next Block.
//
// Warp roles (512 threads): 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue, 12-15 transform.
#include "conv_tc.cuh"

#include <cuda_runtime.h>
#include <mutex>

#include "conv_epilogue.cuh"
#include "ptx.cuh"
#include "launch.cuh"

namespace mi {

namespace {

constexpr int kGnThreads = 512;
constexpr int kTH = 16, kTW = 8, kHW = kTW + 2, kHH = kTH + 2, kHaloPix = kHH * kHW;   // 180 halo pixels
constexpr uint32_t kOpBytes = kHaloPix * 128;                          // fp16 operand tile: 23040 B
constexpr uint32_t kOpStride = (kOpBytes + 1023) & ~1023u;             // 23552
constexpr uint32_t kRawHalf = kHaloPix * 128;                          // fp32, 32 channels: 23040 B
constexpr uint32_t kRawBytes = 2 * kRawHalf;                           // 64 channels: 46080 B
constexpr uint32_t kRawStride = (kRawBytes + 1023) & ~1023u;           // 46080 (already a multiple of 1024)
constexpr int kOpStages = 2;

template <int BLOCK_N>
struct CfgG {
    static constexpr uint32_t kBBytes = BLOCK_N * kConvBlockK * 2;
    static constexpr int kBStages = BLOCK_N == 256 ? 3 : 3;
    static constexpr int kRawStages = BLOCK_N == 256 ? 1 : 2;       // the raw fp32 staging ring
    static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    static constexpr uint32_t kAux = 2048;   // barriers, tmem ptr, group stats, coefficients
    static constexpr uint32_t kSmemBytes = kRawStages * kRawStride + kOpStages * kOpStride + kBStages * kBBytes + kEpiBytes + kAux + 1024;
};

__device__ __forceinline__ uint64_t make_halo_desc_g(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((kHW * 128u) >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kGnThreads, 1)
conv3x3_gn_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvTcArgs args,
                  const __grid_constant__ GnPrologueArgs gn) {
    pdl_trigger();
    using C = CfgG<BLOCK_N>;
    constexpr int NB = C::kBStages;
    constexpr int NR = C::kRawStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_rawbuf = smem;
    uint8_t* s_op = smem + NR * kRawStride;
    uint8_t* s_b = s_op + kOpStages * kOpStride;
    float* epi_base = reinterpret_cast<float*>(s_b + NB * C::kBBytes);
    uint8_t* aux = reinterpret_cast<uint8_t*>(epi_base) + kEpiBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
    uint64_t* fullRaw = bars;                      // [NR] TMA -> transform
    uint64_t* emptyRaw = bars + 2;                 // [NR] transform -> TMA
    uint64_t* xformed = bars + 4;                  // [kOpStages] transform -> MMA
    uint64_t* emptyOp = bars + 4 + kOpStages;      // [kOpStages] MMA -> transform
    uint64_t* fullB = bars + 4 + 2 * kOpStages;    // [NB]
    uint64_t* emptyB = fullB + NB;                 // [NB]
    uint64_t* tfull_bar = emptyB + NB;             // [2]
    uint64_t* tempty_bar = tfull_bar + 2;          // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_mean = reinterpret_cast<float*>(aux + 512);     // [32]
    float* s_rstd = s_mean + 32;                             // [32]
    float* s_coef = s_rstd + 32;                             // [2][64] A, then [2][64] B  (double-buffered by chunk parity)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    int* err = args.err_flag;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmR0);
        ptx::prefetch_tensormap(&tmR1);
        ptx::prefetch_tensormap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NR; ++i) { ptx::mbar_init(&fullRaw[i], 1); ptx::mbar_init(&emptyRaw[i], 128); }
        for (int i = 0; i < kOpStages; ++i) { ptx::mbar_init(&xformed[i], 128); ptx::mbar_init(&emptyOp[i], 1); }
        for (int i = 0; i < NB; ++i) { ptx::mbar_init(&fullB[i], 1); ptx::mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], 32 * kEpiWarps); }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr_smem, C::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();      // everything above is independent of the previous kernel's output

    const int chunks = args.chunks_per_tap;             // 64-channel chunks of the (concatenated) input
    const int tiles_m = args.tiles_w * args.tiles_h * args.tiles_b;
    const int total_tiles = tiles_m * args.tiles_n;
    const int Cin = chunks * kConvBlockK;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int sb = 0, sr = 0;
            uint32_t praw = 0, pb = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % args.tiles_n;
                const int mt = tile / args.tiles_n;
                const int w0 = (mt % args.tiles_w) * kTW;
                const int h0 = ((mt / args.tiles_w) % args.tiles_h) * kTH;
                const int b0 = mt / (args.tiles_w * args.tiles_h);
                const int n0 = nt * BLOCK_N;
                for (int j = 0; j < chunks; ++j) {
                    ptx::mbar_wait(&emptyRaw[sr], praw ^ 1, err, 3100 + sr);
                    ptx::mbar_arrive_expect_tx(&fullRaw[sr], kRawBytes);
                    const bool first = j < args.a_split;
                    const CUtensorMap* tm = first ? &tmR0 : &tmR1;
                    const int c0 = (first ? j : j - args.a_split) * kConvBlockK;
                    uint8_t* rb = s_rawbuf + sr * kRawStride;
                    ptx::tma_load_5d(tm, &fullRaw[sr], rb, c0, w0 - 1, h0 - 1, 0, b0);
                    ptx::tma_load_5d(tm, &fullRaw[sr], rb + kRawHalf, c0 + 32, w0 - 1, h0 - 1, 0, b0);
                    if (++sr == NR) { sr = 0; praw ^= 1; }
                    for (int t = 0; t < 9; ++t) {
                        ptx::mbar_wait(&emptyB[sb], pb ^ 1, err, 3200 + sb);
                        ptx::mbar_arrive_expect_tx(&fullB[sb], C::kBBytes);
                        ptx::tma_load_2d(&tmB, &fullB[sb], s_b + sb * C::kBBytes, t * Cin + j * kConvBlockK, n0);
                        if (++sb == NB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = ptx::make_idesc_f16(kConvBlockM, BLOCK_N, 0);
            int so = 0, sb = 0;
            uint32_t po = 0, pb = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, err, 3300 + as);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int j = 0; j < chunks; ++j) {
                    ptx::mbar_wait(&xformed[so], po, err, 3400 + so);
                    const uint32_t a_base = ptx::smem_u32(s_op + so * kOpStride);
                    for (int t = 0; t < 9; ++t) {
                        ptx::mbar_wait(&fullB[sb], pb, err, 3500 + sb);
                        ptx::tc_fence_after();
                        const uint64_t da = make_halo_desc_g(a_base + ((t / 3) * kHW + (t % 3)) * 128);
                        const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_b + sb * C::kBBytes));
#pragma unroll
                        for (int k = 0; k < kConvBlockK / 16; ++k)
                            ptx::umma_f16(tmem_d, da + 2 * k, db + 2 * k, idesc, (j | t | k) != 0);
                        ptx::umma_commit(&emptyB[sb]);
                        if (++sb == NB) { sb = 0; pb ^= 1; }
                    }
                    ptx::umma_commit(&emptyOp[so]);
                    if (++so == kOpStages) { so = 0; po ^= 1; }
                }
                ptx::umma_commit(&tfull_bar[as]);
            }
        }
    } else if (warp >= 12) {
        // ===================== transform: fp32 halo tile -> GroupNorm/FiLM/SiLU -> fp16 swizzled operand ==========
        const int tt = threadIdx.x - 12 * 32;          // 0..127
        const int qo = tt & 7;                         // physical 16-byte chunk of the operand row this thread writes
        const int p0 = (tt >> 3) & 7;                  // rows p == p0 (mod 8)  ->  the swizzle phase is fixed per thread
        const int lq = qo ^ p0;                        // logical chunk: channels [8*lq, 8*lq + 8) of the 64-channel chunk
        const int half = tt >> 6;                      // two thread halves interleave the rows
        int so = 0, sr = 0;
        uint32_t po = 0, praw = 0;
        const int C0 = gn.C0, Ctot = gn.C0 + gn.C1, Cg = Ctot / gn.groups;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / args.tiles_n;
            const int w0 = (mt % args.tiles_w) * kTW;
            const int h0 = ((mt / args.tiles_w) % args.tiles_h) * kTH;
            const int b = mt / (args.tiles_w * args.tiles_h);
            // --- per-image group statistics from the producers' block statistics
            named_bar_sync(1, 128);
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
                    su += (double)gn.scale1 * gn.stats1[((long long)b * (gn.C1 / 16) + e) * 2];
                    sq += (double)gn.scale1 * (double)gn.scale1 * gn.stats1[((long long)b * (gn.C1 / 16) + e) * 2 + 1];
                }
                const double n = (double)Cg * args.H * args.W;
                const double mean = su / n;
                double var = sq / n - mean * mean;
                if (var < 0) var = 0;
                s_mean[g] = (float)mean;
                s_rstd[g] = rsqrtf((float)var + gn.eps);
            }
            named_bar_sync(1, 128);
            for (int j = 0; j < chunks; ++j) {
                // --- coefficients of this 64-channel chunk (threads 0..63), double-buffered by chunk parity
                float* cA = s_coef + (j & 1) * 64;
                float* cB = s_coef + 128 + (j & 1) * 64;
                if (tt < 64) {
                    const int cc = j * kConvBlockK + tt;                  // channel of the virtual concat
                    const int g = cc / Cg;
                    float a = s_rstd[g] * gn.gamma[cc];
                    float bb = gn.beta[cc] - s_mean[g] * a;
                    if (gn.scale_shift) {
                        const float sc = gn.scale_shift[(long long)b * gn.ss_ld + cc] + 1.0f;
                        const float sh = gn.scale_shift[(long long)b * gn.ss_ld + Ctot + cc];
                        a *= sc;
                        bb = bb * sc + sh;
                    }
                    if (cc >= C0) a *= gn.scale1;                        // skip * 2^-1/2 folded into the multiplier
                    cA[tt] = a;
                    cB[tt] = bb;
                }
                named_bar_sync(1, 128);
                float ca[8], cb[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { ca[e] = cA[lq * 8 + e]; cb[e] = cB[lq * 8 + e]; }

                ptx::mbar_wait(&fullRaw[sr], praw, err, 3600 + sr);     // raw tile landed
                ptx::mbar_wait(&emptyOp[so], po ^ 1, err, 3700 + so);   // operand slot free (its MMAs retired)
                uint8_t* op = s_op + so * kOpStride;
                const uint8_t* rawp = s_rawbuf + sr * kRawStride + (lq >> 2) * kRawHalf + (lq & 3) * 32;   // 8 fp32 = 32 B
                for (int p = p0 + 8 * half; p < kHaloPix; p += 16) {
                    const int ph = p / kHW, pw = p - ph * kHW;
                    const int gh = h0 - 1 + ph, gw = w0 - 1 + pw;
                    uint4 o = make_uint4(0u, 0u, 0u, 0u);
                    if (gh >= 0 && gh < args.H && gw >= 0 && gw < args.W) {
                        const float4 x0 = *reinterpret_cast<const float4*>(rawp + p * 128);
                        const float4 x1 = *reinterpret_cast<const float4*>(rawp + p * 128 + 16);
                        float v[8] = {fmaf(x0.x, ca[0], cb[0]), fmaf(x0.y, ca[1], cb[1]), fmaf(x0.z, ca[2], cb[2]),
                                      fmaf(x0.w, ca[3], cb[3]), fmaf(x1.x, ca[4], cb[4]), fmaf(x1.y, ca[5], cb[5]),
                                      fmaf(x1.z, ca[6], cb[6]), fmaf(x1.w, ca[7], cb[7])};
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __fdividef(v[e], 1.0f + __expf(-v[e]));
                        __half2 h0_ = __floats2half2_rn(v[0], v[1]), h1_ = __floats2half2_rn(v[2], v[3]);
                        __half2 h2_ = __floats2half2_rn(v[4], v[5]), h3_ = __floats2half2_rn(v[6], v[7]);
                        o.x = *reinterpret_cast<uint32_t*>(&h0_); o.y = *reinterpret_cast<uint32_t*>(&h1_);
                        o.z = *reinterpret_cast<uint32_t*>(&h2_); o.w = *reinterpret_cast<uint32_t*>(&h3_);
                    }
                    *reinterpret_cast<uint4*>(op + p * 128 + qo * 16) = o;
                }
                ptx::fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core (async proxy)
                ptx::mbar_arrive(&xformed[so]);
                ptx::mbar_arrive(&emptyRaw[sr]);
                if (++sr == NR) { sr = 0; praw ^= 1; }
                if (++so == kOpStages) { so = 0; po ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp & 3;
        const int c_half = BLOCK_N / 2;
        const int c_begin = warp >= 8 ? c_half : 0;
        const int c_end = c_begin + c_half;
        float* epi_stage = epi_base + (warp - 4) * 32 * kEpiLd;
        const int m = ew * 32 + lane;
        const int bw = m & (kTW - 1);
        const int bh = m >> 3;
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int nt = tile % args.tiles_n;
            const int mt = tile / args.tiles_n;
            const int w = (mt % args.tiles_w) * kTW + bw;
            const int h = ((mt / args.tiles_w) % args.tiles_h) * kTH + bh;
            const int b = mt / (args.tiles_w * args.tiles_h);
            const int n0 = nt * BLOCK_N;
            const bool valid = (b < args.B) && (h < args.H) && (w < args.W);
            const long long pix = (long long)b * args.out_sb + (long long)h * args.out_sh + (long long)w * args.out_sw;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            ptx::mbar_wait(&tfull_bar[as], aphase, err, 3800 + as);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BLOCK_N;
            epilogue_tile<BLOCK_N>(args, taddr, n0, pix, valid, epi_stage, c_begin, c_end, b);
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty_bar[as]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, C::kTmemCols);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_g() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

template <int BLOCK_N>
int launch_gn(const CUtensorMap& r0, const CUtensorMap& r1, const CUtensorMap& tmB, const ConvTcArgs& a,
              const GnPrologueArgs& g, int total, int num_sms, cudaStream_t st) {
    using C = CfgG<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv3x3_gn_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes) !=
            cudaSuccess)
            return -10;
        attr_set = true;
    }
    const int grid = total < num_sms ? total : num_sms;
    launch_k(conv3x3_gn_kernel<BLOCK_N>, grid, kGnThreads, C::kSmemBytes, st, r0, r1, tmB, a, g);
    return cudaGetLastError() == cudaSuccess ? 0 : -11;
}

int encode_raw(PFN_encodeTiled enc, CUtensorMap* tm, const float* src, int C, int B, int H, int W) {
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1, (cuuint64_t)B};
    cuuint64_t gstr[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[5] = {32, kHW, kHH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(src), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0
               : -6;
}

}  // namespace

bool conv_gn_supported(int H, int W, int C0, int C1, int Cout, int groups) {
    const int C = C0 + C1;
    if (H % kTH || W % kTW || C0 % 64 || C1 % 64 || C <= 0 || Cout % 128) return false;
    if (groups < 1 || groups > 32 || C % groups) return false;
    const int Cg = C / groups;
    return Cg % 16 == 0;
}

int conv_gn_launch(const ConvGnProblem& p, cudaStream_t stream) {
    if (!conv_gn_supported(p.H, p.W, p.C0, p.C1, p.Cout, p.groups)) return -3;
    if (p.C1 && (!p.src1 || !p.stats1)) return -8;
    if ((reinterpret_cast<uintptr_t>(p.src0) & 15) || (reinterpret_cast<uintptr_t>(p.wpacked) & 15)) return -8;
    PFN_encodeTiled enc = get_encode_g();
    if (!enc) return -5;
    const int C = p.C0 + p.C1;

    ConvTcArgs a{};
    a.num_taps = 9;
    a.chunks_per_tap = C / kConvBlockK;
    a.a_split = p.C0 / kConvBlockK;
    a.tiles_w = p.W / kTW; a.tiles_h = p.H / kTH; a.tiles_b = p.B;
    a.B = p.B; a.H = p.H; a.W = p.W;
    a.out_sb = (long long)p.H * p.W * p.Cout; a.out_sh = (long long)p.W * p.Cout; a.out_sw = p.Cout; a.out_sc = 1;
    a.n_valid = p.Cout;
    a.out_f32 = p.out_f32; a.out_f16 = p.out_f16; a.bias = p.bias; a.residual = p.residual; a.err_flag = p.err_flag;
    a.stats = p.out_stats; a.stats_blocks = p.Cout / 16;

    GnPrologueArgs g{};
    g.C0 = p.C0; g.C1 = p.C1; g.groups = p.groups; g.scale1 = p.scale1; g.eps = p.eps;
    g.stats0 = p.stats0; g.stats1 = p.stats1; g.gamma = p.gamma; g.beta = p.beta;
    g.scale_shift = p.scale_shift; g.ss_ld = p.ss_ld;

    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    const int bn = (p.Cout % 256 == 0) ? 256 : 128;
    a.tiles_n = p.Cout / bn;

    CUtensorMap r0, r1, tmB;
    if (encode_raw(enc, &r0, p.src0, p.C0, p.B, p.H, p.W)) return -6;
    if (p.C1) { if (encode_raw(enc, &r1, p.src1, p.C1, p.B, p.H, p.W)) return -6; }
    else r1 = r0;
    const cuuint64_t K = (cuuint64_t)9 * C;
    cuuint64_t wdim[2] = {K, (cuuint64_t)p.Cout};
    cuuint64_t wstr[1] = {K * 2};
    cuuint32_t wbox[2] = {kConvBlockK, (cuuint32_t)bn};
    cuuint32_t westr[2] = {1, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(p.wpacked), wdim, wstr, wbox, westr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -7;
    const int total = a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n;
    return bn == 256 ? launch_gn<256>(r0, r1, tmB, a, g, total, num_sms, stream)
                     : launch_gn<128>(r0, r1, tmB, a, g, total, num_sms, stream);
}

}  // namespace mi
