// Micro-benchmark: what does one tcgen05.mma (M = 128, K = 16, fp16 -> fp32) cost on B200, by N, by where A lives (shared memory
// "SS" / tensor memory "TS"), and by how many accumulators the stream alternates between?  One CTA per SM, one issuing thread,
// operands resident (no TMA, no epilogue); clocks per instruction = (clock64 around issue + final commit wait) / #instructions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I minimagen_b200/csrc tools/mma_cost.cu -o tools/mma_cost
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace mi;

struct Cfg { int N, ts, accs, iters, kchunk; };

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(Cfg c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 128 rows x 64 K (16 KB), zeros
    uint8_t* sB = smem + 16384;         // 256 rows x 64 K (32 KB)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) { ptx::tmem_alloc(&tmem_ptr, 512); ptx::tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::make_idesc_f16(128, c.N, 0);
        const uint64_t da = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sA));
        const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB));
        const uint32_t ta = tm + 448;                        // A operand region in TMEM (garbage values are fine)
        long long t0 = 0, t1 = 0;
        if (ptx::elect_one()) {
            t0 = clock64();
            for (int it = 0; it < c.iters; ++it) {
                const uint32_t d = tm + (it % c.accs) * c.N;
#pragma unroll 4
                for (int k = 0; k < c.kchunk; ++k) {
                    if (c.ts) ptx::umma_f16_ts(d, ta + 8 * (k & 3), db + 2 * (k & 3), idesc, 1);
                    else ptx::umma_f16(d, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
                }
            }
            ptx::umma_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0, nullptr, 0);
        if (ptx::elect_one()) {
            t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

int main() {
    long long* out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("# M=128 K=16 fp16 tcgen05.mma, clocks per instruction (SM clock attr %d kHz); ideal = N/2\n", clk_khz);
    for (int grid : {1, 148})
        for (int ts = 0; ts < 2; ++ts)
            for (int N : {64, 128, 256})
                for (int accs : {1, 2}) {
                    if (accs * N > 448) continue;
                    Cfg c{N, ts, accs, 2000, 4};
                    long long h = 0;
                    for (int rep = 0; rep < 2; ++rep) {
                        mma_cost_kernel<<<grid, 128, 64 * 1024>>>(c, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
                    printf("grid %3d  %s  N=%3d  accumulators=%d : %7.1f clk / MMA  (ideal %d)\n", grid, ts ? "TS" : "SS", N, accs,
                           (double)h / (c.iters * c.kchunk), N / 2);
                }
    return 0;
}
