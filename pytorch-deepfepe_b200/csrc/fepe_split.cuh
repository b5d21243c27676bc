// Split-fp16 helpers shared by the fp32-parity tensor-core kernels (fepe_mlp32.cu forward, fepe_mlp32_bwd.cu backward).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fepe {
namespace m32 {

// the power of two that scales a tensor whose largest magnitude has the bit pattern `bits` into [2^13, 2^14)
__device__ __forceinline__ float pow2_scale(unsigned bits) {
    const float amax = __uint_as_float(bits);
    if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.f;
    int ex;
    frexpf(amax, &ex);                                     // amax in [2^(ex-1), 2^ex)
    int e = 14 - ex;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    return ldexpf(1.f, e);
}

// x = hi + lo in fp16 (round to nearest, saturating): the low half of each result holds the first element
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - f.y), "f"(x0 - f.x));
}


}  // namespace m32
}  // namespace fepe
