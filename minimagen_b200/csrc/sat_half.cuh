// Saturating fp32 -> fp16 conversion of ACTIVATIONS (tensor-core operands): values beyond +-65504 become the largest finite
// half instead of +-inf, so that one out-of-range activation of a real checkpoint degrades gracefully (bounded error) rather
// than poisoning every later layer with inf/NaN.  One F2FP.SATFINITE instruction, same cost as the plain conversion.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace mi {

__device__ __forceinline__ __half sat_half(float x) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return __ushort_as_half(r);
}
__device__ __forceinline__ __half2 sat_half2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return *reinterpret_cast<__half2*>(&r);
}

}  // namespace mi
