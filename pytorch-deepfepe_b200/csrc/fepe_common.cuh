// Shared device helpers: mbarrier / bulk-copy (TMA, 1-D) PTX wrappers, warp reductions, and the
// persistent "pair ring" every streaming kernel of this library uses.
//
// Pair ring.  One CTA per SM, resident for the whole launch.  Warp `kConsumers` is the producer:
// lane 0 issues, for each image pair assigned to the CTA, two cp.async.bulk copies (the pair's
// [N,4] coordinates and [N] weights, 20 B per correspondence) into one of S shared-memory stages
// and arms that stage's `full` mbarrier with the byte count.  Consumer warp w owns pairs
// w, w+C, w+2C, ... of the CTA: it waits on `full`, makes all of its passes over the pair out of
// shared memory (HBM is read exactly once), writes its outputs and releases the stage through the
// `empty` mbarrier.  With S > C the producer is always S-C pairs ahead, so HBM latency is hidden
// without needing occupancy.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fepe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// Programmatic dependent launch: let a dependent grid of the stream (one launched with the programmatic-serialization
// attribute, e.g. fepe_pose_fwd) be scheduled early; it still waits for this grid's completion before it reads our
// outputs (griddepcontrol.wait).  No effect on ordinary successors.
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion counted on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce-scatter of LEN per-lane doubles across the warp: after the call lane l holds in v[0..cnt)
// the warp-wide sums of elements base .. base+cnt of the original vector (cnt <= 2 for LEN = 36).
// Each butterfly level halves the live vector instead of exchanging all of it: 37 exchanges instead
// of 180 for LEN = 36.
template <int LEN, int XOR>
struct ReduceScatter {
    template <int CAP>
    static __device__ __forceinline__ void run(double (&v)[CAP], int lane, int& base, int& cnt) {
        constexpr int H = (LEN + 1) / 2;
        const bool up = (lane & XOR) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const double lower = v[i];
            const double upper = (i + H < LEN) ? v[i + H] : 0.0;
            const double send = up ? lower : upper;
            const double keep = up ? upper : lower;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, XOR);
        }
        if (up) { base += H; cnt = max(cnt - H, 0); } else { cnt = min(cnt, H); }
        if constexpr (XOR > 1) ReduceScatter<H, XOR / 2>::run(v, lane, base, cnt);
    }
};

struct RingLayout {
    int stages;        // S
    int consumers;     // C
    int stage_bytes;   // multiple of 128
    int bar_off;       // byte offset of full[S], empty[S]
    int scratch_off;   // byte offset of per-consumer scratch
    int total_bytes;
};

}  // namespace fepe
