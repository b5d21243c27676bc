// Backward pass of the fp32-parity tensor-core MLP (forward: fepe_mlp32.cu).  What autograd does for the reference's
// ErrorEstimator (Conv1d k=1 / InstanceNorm1d(affine) / LeakyReLU, deepFEPE/models/ErrorEstimators.py:46-64;
// loss.backward() at Train_model_pipeline.py:595), with fp32 tensors in memory and the two GEMM-shaped gradients on
// tcgen05 with split-fp16 operands (three MMAs per product, fp32 accumulation):
//
//   fepe_mlp32_last_bwd     final Conv1d: dX' = dlogits W, dW, db (CUDA cores; Co = 1 or 4)
//   fepe_mlp32_normbwd      InstanceNorm(affine) + LeakyReLU adjoint; the block output x' = LeakyReLU(a y + d) is
//                           RECOMPUTED from the saved pre-norm y (the forward never stored it):
//                               dZ = dX' (t > 0 ? 1 : slope),  t = a y + d
//                               A1 = sum_n dZ,  A2 = sum_n dZ yhat,  yhat = (y - mean) rstd          (fp64 sums)
//                               dY = rstd gamma (dZ - A1/N - yhat A2/N),  padded rows zero
//                           and max |dY| (the power-of-two scale of the tensor-core consumers)
//   fepe_mlp32_wgrad        dW[Co,Ci] += dY[M,Co]^T X'[M,Ci], X' recomputed from the previous block's y on the operand
//                           path.  Both operands are MN-major for the tensor core (the GEMM-K index is the row m), read
//                           in place from the row-major fp32 activations: a TMA box is 64 rows x 32 channels (128 B),
//                           two boxes = one 64-channel group, which the transform threads (one per row) turn IN PLACE
//                           into the fp16 hi tile (first box) and lo tile (second box) of the same swizzled geometry.
//   data gradient           dX' = dY W is fepe_mlp32_gemm with ss = NULL, (W^T)_hi/lo and a_amax = max |dY|
//   fepe_mlp32_first_bwd    layer 1 (Ci <= 16): dX0 = dY W, dW += dY^T X0 on CUDA cores
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_common.cuh"
#include "fepe_dispatch.cuh"
#include "fepe_split.cuh"
#include "fepe_umma.cuh"

namespace fepe {
namespace m32 {

// ------------------------------------------------------------------------------------------------
// last layer: logits[b,o,n] = x'[m,:] . W[o,:] + bias[o], x' = LeakyReLU(a y + d).  One CTA per (`rows`-row slab, pair),
// thread = channel (Ci = 256); rows = 128, or less for batches that would not fill the machine with 128-row slabs.
template <int CO>
__global__ void __launch_bounds__(256) last_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ Y,
                                                       const float2* __restrict__ ss, float slope,
                                                       const float* __restrict__ W, float* __restrict__ dX,
                                                       float* __restrict__ dW, float* __restrict__ db, int N, int Npad,
                                                       int Ci, int rows) {
    __shared__ float dl[CO][128];
    const int b = blockIdx.y, r0 = blockIdx.x * rows;
    for (int i = threadIdx.x; i < CO * 128; i += 256) {
        const int o = i >> 7, r = i & 127;
        dl[o][r] = (r < rows && r0 + r < N) ? dlogits[(static_cast<size_t>(b) * CO + o) * N + r0 + r] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Ci; k += 256) {
        const float2 ad = __ldg(ss + static_cast<size_t>(b) * Ci + k);
        float wk[CO], acc[CO];
#pragma unroll
        for (int o = 0; o < CO; ++o) { wk[o] = __ldg(W + o * Ci + k); acc[o] = 0.f; }
        const size_t base = (static_cast<size_t>(b) * Npad + r0) * Ci + k;
#pragma unroll 8
        for (int r = 0; r < rows; ++r) {
            const float t = fmaf(Y[base + static_cast<size_t>(r) * Ci], ad.x, ad.y);
            const float x = fmaxf(t, slope * t);
            float g = 0.f;
#pragma unroll
            for (int o = 0; o < CO; ++o) {
                g = fmaf(dl[o][r], wk[o], g);
                acc[o] = fmaf(dl[o][r], x, acc[o]);
            }
            dX[base + static_cast<size_t>(r) * Ci] = g;
        }
#pragma unroll
        for (int o = 0; o < CO; ++o) atomicAdd(dW + o * Ci + k, acc[o]);
    }
    if (threadIdx.x < CO) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += dl[threadIdx.x][r];
        atomicAdd(db + threadIdx.x, s);
    }
}

// dgamma[c] += sum_b A[b,c,1], dbeta[c] += sum_b A[b,c,0] for up to kAffineMaxSeg blocks whose A arrays lie back to back
// ([B,C_0,2], [B,C_1,2], ...): ONE launch for the affine gradients of a whole estimator, accumulating straight into the
// caller's gradient buffers (a thread per channel; launches on one stream are serial, so a plain read-modify-write).
constexpr int kAffineMaxSeg = 8;
struct AffineGradParams {
    const double* A;
    int B, nseg;
    int C[kAffineMaxSeg];
    float* dgamma[kAffineMaxSeg];
    float* dbeta[kAffineMaxSeg];
};
__global__ void __launch_bounds__(128) affine_grads_kernel(const AffineGradParams p) {
    int c = blockIdx.x * 128 + threadIdx.x;
    const double* A = p.A;
    for (int s = 0; s < p.nseg; ++s) {
        const int C = p.C[s];
        if (c < C) {
            double s1 = 0.0, s2 = 0.0;
            for (int b = 0; b < p.B; ++b) {
                const double2 v = *reinterpret_cast<const double2*>(A + (static_cast<size_t>(b) * C + c) * 2);
                s1 += v.x; s2 += v.y;
            }
            p.dbeta[s][c] += static_cast<float>(s1);
            p.dgamma[s][c] += static_cast<float>(s2);
            return;
        }
        c -= C;
        A += static_cast<size_t>(p.B) * C * 2;
    }
}

// ------------------------------------------------------------------------------------------------
// InstanceNorm + LeakyReLU adjoint.  One CTA per (128-row slab, pair); a thread owns 4 fixed channels (their a, d, mean,
// rstd live in registers) and walks rows.  C / 4 must be a power of two <= 256 or a multiple of 256.
struct NormBwdParams {
    const float* dX;        // [M,C] gradient of the block output
    const float* Y;         // [M,C] saved pre-norm output
    const float2* ss;       // [B,C] (a, d)
    const float2* mr;       // [B,C] (mean, rstd)
    const float* gamma;     // [C]
    double* A;              // [B,C,2] (A1, A2), zeroed by the caller
    float* dY;              // [M,C]
    unsigned* amax;         // bits of max |dY|, zeroed by the caller
    int C, Npad, Nvalid;
    float slope;
    int rows;               // rows of a pair per CTA: 128, or less when B * Npad / 128 CTAs would not fill the machine
};

__global__ void __launch_bounds__(256) normbwd_reduce_kernel(const NormBwdParams p) {
    extern __shared__ float acc[];           // [2*C]
    const int b = blockIdx.y, r0 = blockIdx.x * p.rows;
    const int C = p.C, vpr = C / 4;
    for (int i = threadIdx.x; i < 2 * C; i += 256) acc[i] = 0.f;
    __syncthreads();
    const int cols = vpr < 256 ? vpr : 256;              // channel vectors handled side by side
    const int nrg = 256 / cols, rg = threadIdx.x / cols;
    int rows = p.Nvalid - r0;
    rows = rows > p.rows ? p.rows : rows;
    if (rows <= 0) return;                               // a tile of padding rows only: nothing to add
    for (int v = threadIdx.x % cols; v < vpr; v += cols) {
        const int c0 = v * 4;
        float a[4], d[4], mu[4], rs[4], a1[4], a2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 s2 = __ldg(p.ss + static_cast<size_t>(b) * C + c0 + k);
            const float2 m2 = __ldg(p.mr + static_cast<size_t>(b) * C + c0 + k);
            a[k] = s2.x; d[k] = s2.y; mu[k] = m2.x; rs[k] = m2.y; a1[k] = 0.f; a2[k] = 0.f;
        }
        const size_t base = (static_cast<size_t>(b) * p.Npad + r0) * C + c0;
#pragma unroll 4
        for (int r = rg; r < rows; r += nrg) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.dX + base + static_cast<size_t>(r) * C));
            const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.Y + base + static_cast<size_t>(r) * C));
            const float g[4] = {g4.x, g4.y, g4.z, g4.w}, y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float t = fmaf(y[k], a[k], d[k]);
                const float dz = t > 0.f ? g[k] : p.slope * g[k];
                a1[k] += dz;
                a2[k] = fmaf(dz, (y[k] - mu[k]) * rs[k], a2[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { atomicAdd(&acc[2 * (c0 + k)], a1[k]); atomicAdd(&acc[2 * (c0 + k) + 1], a2[k]); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += 256) atomicAdd(p.A + static_cast<size_t>(b) * C * 2 + i, static_cast<double>(acc[i]));
}

__global__ void __launch_bounds__(256) normbwd_apply_kernel(const NormBwdParams p) {
    const int b = blockIdx.y, r0 = blockIdx.x * p.rows;
    const int C = p.C, vpr = C / 4;
    const int cols = vpr < 256 ? vpr : 256;
    const int nrg = 256 / cols, rg = threadIdx.x / cols;
    int rows = p.Nvalid - r0;
    rows = rows > p.rows ? p.rows : (rows < 0 ? 0 : rows);
    const double invN = 1.0 / static_cast<double>(p.Nvalid);
    float amax = 0.f;
    for (int v = threadIdx.x % cols; v < vpr; v += cols) {
        const int c0 = v * 4;
        float a[4], d[4], mu[4], rs[4], k0[4], k1[4], k2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t ch = static_cast<size_t>(b) * C + c0 + k;
            const float2 s2 = __ldg(p.ss + ch);
            const float2 m2 = __ldg(p.mr + ch);
            a[k] = s2.x; d[k] = s2.y; mu[k] = m2.x; rs[k] = m2.y;
            k0[k] = m2.y * __ldg(p.gamma + c0 + k);
            k1[k] = static_cast<float>(p.A[ch * 2] * invN);
            k2[k] = static_cast<float>(p.A[ch * 2 + 1] * invN);
        }
        const size_t base = (static_cast<size_t>(b) * p.Npad + r0) * C + c0;
#pragma unroll 4
        for (int r = rg; r < p.rows; r += nrg) {
            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);                    // padded rows stay zero
            if (r < rows) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.dX + base + static_cast<size_t>(r) * C));
                const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.Y + base + static_cast<size_t>(r) * C));
                const float g[4] = {g4.x, g4.y, g4.z, g4.w}, y[4] = {y4.x, y4.y, y4.z, y4.w};
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float t = fmaf(y[k], a[k], d[k]);
                    const float dz = t > 0.f ? g[k] : p.slope * g[k];
                    o[k] = k0[k] * (dz - k1[k] - (y[k] - mu[k]) * rs[k] * k2[k]);
                    amax = fmaxf(amax, fabsf(o[k]));
                }
                out = make_float4(o[0], o[1], o[2], o[3]);
            }
            *reinterpret_cast<float4*>(p.dY + base + static_cast<size_t>(r) * C) = out;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0 && amax > 0.f && amax < 3.0e38f) atomicMax(p.amax, __float_as_uint(amax));
}

// ------------------------------------------------------------------------------------------------
// layer 1 backward (Ci <= 16, Co = 64): dX0[b,n,k] = sum_c dY[m,c] W[c,k], dW[c,k] += sum_m dY[m,c] X0[m,k].
constexpr int kFirstMaxCi = 16;
__global__ void __launch_bounds__(128) first_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X0,
                                                        const float* __restrict__ W, float* __restrict__ dX0,
                                                        float* __restrict__ dW, int N, int Npad, int Ci) {
    __shared__ float w_s[64 * kFirstMaxCi];
    __shared__ float dy_s[128][64 + 1];
    __shared__ float x_s[128][kFirstMaxCi];
    const int b = blockIdx.y, r0 = blockIdx.x * 128, r = r0 + threadIdx.x;
    for (int i = threadIdx.x; i < 64 * Ci; i += 128) w_s[i] = W[i];
    const bool valid = r < N;
#pragma unroll
    for (int k = 0; k < kFirstMaxCi; ++k) x_s[threadIdx.x][k] = (valid && k < Ci) ? X0[(static_cast<size_t>(b) * N + r) * Ci + k] : 0.f;
    // coalesced load of the 128 x 64 tile of dY (rows >= N of the pair are zero in dY already)
    for (int i = threadIdx.x; i < 128 * 16; i += 128) {
        const int rr = i >> 4, c4 = (i & 15) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(dY + (static_cast<size_t>(b) * Npad + r0 + rr) * 64 + c4));
        dy_s[rr][c4] = v.x; dy_s[rr][c4 + 1] = v.y; dy_s[rr][c4 + 2] = v.z; dy_s[rr][c4 + 3] = v.w;
    }
    __syncthreads();
    if (valid && dX0 != nullptr) {
        float g[kFirstMaxCi];
#pragma unroll
        for (int k = 0; k < kFirstMaxCi; ++k) g[k] = 0.f;
        for (int c = 0; c < 64; ++c) {
            const float dd = dy_s[threadIdx.x][c];
#pragma unroll
            for (int k = 0; k < kFirstMaxCi; ++k)
                if (k < Ci) g[k] = fmaf(dd, w_s[c * Ci + k], g[k]);
        }
#pragma unroll
        for (int k = 0; k < kFirstMaxCi; ++k)
            if (k < Ci) dX0[(static_cast<size_t>(b) * N + r) * Ci + k] = g[k];
    }
    for (int o = threadIdx.x; o < 64 * Ci; o += 128) {
        const int c = o / Ci, k = o % Ci;
        float acc = 0.f;
        for (int rr = 0; rr < 128; ++rr) acc = fmaf(dy_s[rr][c], x_s[rr][k], acc);
        atomicAdd(dW + o, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on tensor cores.  Output tile 128 (co) x BN (ci) per CTA and row slab; k-block = 64 rows.
// A stage holds G = 2 + BN/64 channel GROUPS of 16 KB: [box 0: 64 rows x channels 0..31 fp32 | box 1: channels 32..63]
// -> after the transform [hi tile: 64 rows x 64 fp16 | lo tile].  Operand A (dY, 128 channels) = groups 0, 1; operand B
// (X', BN channels) = groups 2...  MN-major descriptors: 64-channel groups are 16 KB apart (leading-byte-offset), 8-row
// groups 1024 B (stride-byte-offset), one MMA consumes 16 rows = 2048 B.
// 16 warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = epilogue, 8.. = transform (one thread per
// (group, row)).
constexpr int kWgThreads = 512;
// pipeline depth by tile width: a stage is (2 + BN / 64) groups of 16 KB
template <int BN> struct WgStages { static constexpr int value = (BN == 256) ? 2 : 3; };
constexpr int kGroupBytes = 64 * 64 * 4;      // 16 KB

__device__ __forceinline__ uint64_t umma_desc_mn_sw128_g(const void* smem_ptr) {
    const uint64_t addr = static_cast<uint64_t>(smem_u32(smem_ptr));
    return ((addr >> 4) & 0x3FFFull) | (static_cast<uint64_t>(kGroupBytes >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}

struct WgradParams {
    int M, Co, Ci, Npad;
    int rows_per_slab;           // multiple of 64
    const float* ss;             // [B, Ci, 2] (a, d) of the block that produced X (x' = LeakyReLU(a y + d))
    float slope;
    const unsigned* dy_amax;     // bits of max |dY|
    float* dW;                   // [Co, Ci] fp32, accumulated with atomics (zeroed by the caller)
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
fepe_mlp32_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                        const WgradParams p) {
    constexpr int G = 2 + BN / 64;
    constexpr int kWgStages = WgStages<BN>::value;
    constexpr int kStageBytes = G * kGroupBytes;
    constexpr int kSsBytes = BN * 8;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ss_ring = smem + kWgStages * kStageBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(ss_ring + kWgStages * kSsBytes);
    uint64_t* empty = full + kWgStages;
    uint64_t* ready = empty + kWgStages;
    uint64_t* tmem_full = ready + kWgStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ci0 = blockIdx.x * BN;
    const int co0 = blockIdx.y * 128;
    const int row0 = blockIdx.z * p.rows_per_slab;
    int rows = p.M - row0;
    if (rows > p.rows_per_slab) rows = p.rows_per_slab;
    const int num_kb = rows / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 2 * G); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(static_cast<uint32_t>(BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---------------- transform: one thread per (group, row) of a stage ----------------
    // 2 G warps: warps 8..15 take groups 0..3; at BN = 256 the four epilogue warps (idle until the last k-block) take
    // groups 4 and 5 before they turn to the accumulator.
    auto transform = [&](const int tt) {
        const int g = tt >> 6, row = tt & 63;
        const int sw = row & 7;
        const bool is_dy = g < 2;
        const float dy_scale = pow2_scale(__ldg(p.dy_amax));
        const float slope = p.slope;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kWgStages;
            const uint32_t ph = static_cast<uint32_t>(kb / kWgStages) & 1u;
            mbar_wait(&full[s], ph);
            unsigned char* a0 = smem + s * kStageBytes + g * kGroupBytes + row * 128;
            unsigned char* a1 = a0 + kGroupBytes / 2;
            float v[64];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 t0 = *reinterpret_cast<const float4*>(a0 + ((c ^ sw) << 4));
                const float4 t1 = *reinterpret_cast<const float4*>(a1 + ((c ^ sw) << 4));
                v[4 * c + 0] = t0.x; v[4 * c + 1] = t0.y; v[4 * c + 2] = t0.z; v[4 * c + 3] = t0.w;
                v[32 + 4 * c + 0] = t1.x; v[32 + 4 * c + 1] = t1.y; v[32 + 4 * c + 2] = t1.z; v[32 + 4 * c + 3] = t1.w;
            }
            if (is_dy) {
#pragma unroll
                for (int q = 0; q < 64; ++q) v[q] *= dy_scale;
            } else {
                const float4* cf = reinterpret_cast<const float4*>(ss_ring + s * kSsBytes) + (g - 2) * 32;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float4 c4 = cf[q];
                    const float t0 = fmaf(v[2 * q], c4.x, c4.y), t1 = fmaf(v[2 * q + 1], c4.z, c4.w);
                    v[2 * q] = fmaxf(t0, slope * t0);
                    v[2 * q + 1] = fmaxf(t1, slope * t1);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) split2(v[8 * c + 2 * q], v[8 * c + 2 * q + 1], h[q], l[q]);
                *reinterpret_cast<uint4*>(a0 + ((c ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(a1 + ((c ^ sw) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
        }
    };

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kWgStages;
                const uint32_t ph = static_cast<uint32_t>(kb / kWgStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* sa = smem + s * kStageBytes;
                mbar_arrive_expect_tx(&full[s], kStageBytes + kSsBytes);
                const int r = row0 + kb * 64;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    tma_load_2d(sa + g * kGroupBytes, &map_dy, co0 + g * 64, r, &full[s]);
                    tma_load_2d(sa + g * kGroupBytes + kGroupBytes / 2, &map_dy, co0 + g * 64 + 32, r, &full[s]);
                }
#pragma unroll
                for (int g = 0; g < BN / 64; ++g) {
                    tma_load_2d(sa + (2 + g) * kGroupBytes, &map_x, ci0 + g * 64, r, &full[s]);
                    tma_load_2d(sa + (2 + g) * kGroupBytes + kGroupBytes / 2, &map_x, ci0 + g * 64 + 32, r, &full[s]);
                }
                const int pair = r / p.Npad;
                bulk_g2s(ss_ring + s * kSsBytes, p.ss + (static_cast<size_t>(pair) * p.Ci + ci0) * 2, kSsBytes, &full[s]);
            }
        }
    } else if (warp == 1) {
        // D = f32, A = B = f16 (format 0), BOTH MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                   (static_cast<uint32_t>(128 >> 4) << 24);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kWgStages;
            const uint32_t ph = static_cast<uint32_t>(kb / kWgStages) & 1u;
            mbar_wait(&ready[s], ph);
            tcgen05_fence_after();
            if (lane == 0) {
                const unsigned char* sa = smem + s * kStageBytes;
                const uint64_t ah = umma_desc_mn_sw128_g(sa), al = umma_desc_mn_sw128_g(sa + kGroupBytes / 2);
                const uint64_t bh = umma_desc_mn_sw128_g(sa + 2 * kGroupBytes), bl = umma_desc_mn_sw128_g(sa + 2 * kGroupBytes + kGroupBytes / 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {      // 16 rows = 2048 B per MMA
                    const uint64_t o = static_cast<uint64_t>(k * 128);
                    umma_bf16(tmem_base, ah + o, bh + o, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_bf16(tmem_base, al + o, bh + o, idesc, 1u);
                    umma_bf16(tmem_base, ah + o, bl + o, idesc, 1u);
                }
                tcgen05_commit(&empty[s]);
                if (kb == num_kb - 1) tcgen05_commit(tmem_full);
            }
            __syncwarp();
        }
    } else if (warp >= 8 && warp < 8 + (2 * G < 8 ? 2 * G : 8)) {
        transform(static_cast<int>(threadIdx.x) - 256);
    } else if (warp >= 4 && warp < 8 && num_kb > 0) {
        if constexpr (2 * G > 8) transform(256 + static_cast<int>(threadIdx.x) - 128);
        const int q = warp & 3;
        const int row = q * 32 + lane;                         // output row = channel co0 + row
        const float inv = 1.f / pow2_scale(__ldg(p.dy_amax));
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        float* out = p.dW + static_cast<size_t>(co0 + row) * p.Ci + ci0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
            // vector reductions (REDG.ADD.F32x4): a quarter of the L2 transactions of scalar atomics -- with one row per
            // lane every lane touches its own cache line, so the epilogue is bound by the number of requests
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c + j),
                             "f"(__uint_as_float(v[j]) * inv), "f"(__uint_as_float(v[j + 1]) * inv),
                             "f"(__uint_as_float(v[j + 2]) * inv), "f"(__uint_as_float(v[j + 3]) * inv)
                             : "memory");
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(static_cast<uint32_t>(BN)));
    }
}

template <int BN>
static int launch_wgrad(const float* dY, const float* X, const WgradParams& p, int slabs, cudaStream_t stream) {
    CUtensorMap my, mx;
    if (!make_map_2d(&my, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dY, p.M, p.Co, 32, 64) ||
        !make_map_2d(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, p.M, p.Ci, 32, 64))
        return FEPE_E_NODEVICE;
    constexpr int smem = WgStages<BN>::value * ((2 + BN / 64) * kGroupBytes + BN * 8) + 256 + 1024;
    static_assert(smem <= 232448, "stage ring does not fit");
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fepe_mlp32_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured[dev & 63] = true;
    }
    dim3 grid(p.Ci / BN, p.Co / 128, slabs);
    fepe_mlp32_wgrad_kernel<BN><<<grid, kWgThreads, smem, stream>>>(my, mx, p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace m32
}  // namespace fepe

extern "C" {

int fepe_mlp32_last_bwd(const float* dlogits, const float* Y, const float* ss, float slope, const float* W, float* dX,
                        float* dW, float* db, int B, int N, int Npad, int Ci, int Co, void* stream) {
    if (!dlogits || !Y || !ss || !W || !dX || !dW || !db || B <= 0 || N <= 0 || Npad < N || (Npad % 128) != 0 || Ci <= 0 ||
        (Co != 1 && Co != 4))
        return FEPE_E_BADARG;
    int rows = 128;
    while (rows > 16 && (Npad / rows) * B < 4 * 148) rows >>= 1;
    dim3 grid(Npad / rows, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* s2 = reinterpret_cast<const float2*>(ss);
    if (Co == 1) fepe::m32::last_bwd_kernel<1><<<grid, 256, 0, st>>>(dlogits, Y, s2, slope, W, dX, dW, db, N, Npad, Ci, rows);
    else fepe::m32::last_bwd_kernel<4><<<grid, 256, 0, st>>>(dlogits, Y, s2, slope, W, dX, dW, db, N, Npad, Ci, rows);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_normbwd(const float* dX, const float* Y, const float* ss, const float* mean_rstd, const float* gamma,
                       float slope, double* A, float* dY, unsigned* dy_amax, int B, int Npad, int Nvalid, int C,
                       void* stream) {
    const int vpr = C / 4;
    if (!dX || !Y || !ss || !mean_rstd || !gamma || !A || !dY || !dy_amax || B <= 0 || C <= 0 || (C % 4) != 0 ||
        (Npad % 128) != 0 || Nvalid <= 0 || Nvalid > Npad || !((vpr <= 256 && (vpr & (vpr - 1)) == 0) || (vpr % 256) == 0) ||
        2 * C * sizeof(float) > 48 * 1024 || (reinterpret_cast<uintptr_t>(dX) & 15u) || (reinterpret_cast<uintptr_t>(Y) & 15u) ||
        (reinterpret_cast<uintptr_t>(dY) & 15u))
        return FEPE_E_BADARG;
    fepe::m32::NormBwdParams p{dX, Y, reinterpret_cast<const float2*>(ss), reinterpret_cast<const float2*>(mean_rstd), gamma,
                               A, dY, dy_amax, C, Npad, Nvalid, slope, 128};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Small batches (a training step has 16 pairs per GPU): 128-row CTAs would leave most SMs without work and every
    // thread with one pair of loads in flight; shrink the row tile until ~4 CTAs per SM exist (>= 32 rows: every CTA
    // also pays 2 C shared + 2 C global atomics for the statistics).
    while (p.rows > 32 && (Npad / p.rows) * B < 4 * 148) p.rows >>= 1;
    dim3 grid(Npad / p.rows, B);
    fepe::m32::normbwd_reduce_kernel<<<grid, 256, 2 * C * sizeof(float), st>>>(p);
    fepe::m32::normbwd_apply_kernel<<<grid, 256, 0, st>>>(p);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_affine_grads(const double* A, int B, int nseg, const int* C, float* const* dgamma, float* const* dbeta,
                            void* stream) {
    if (!A || !C || !dgamma || !dbeta || B <= 0 || nseg <= 0 || nseg > fepe::m32::kAffineMaxSeg) return FEPE_E_BADARG;
    fepe::m32::AffineGradParams p{};
    p.A = A; p.B = B; p.nseg = nseg;
    int total = 0;
    for (int i = 0; i < nseg; ++i) {
        if (C[i] <= 0 || !dgamma[i] || !dbeta[i]) return FEPE_E_BADARG;
        p.C[i] = C[i]; p.dgamma[i] = dgamma[i]; p.dbeta[i] = dbeta[i];
        total += C[i];
    }
    fepe::m32::affine_grads_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_first_bwd(const float* dY, const float* X0, const float* W, float* dX0, float* dW, int B, int N, int Npad,
                         int Ci, int Co, void* stream) {
    if (!dY || !X0 || !W || !dW || B <= 0 || N <= 0 || Ci <= 0 || Ci > fepe::m32::kFirstMaxCi || Co != 64 || (Npad % 128) != 0 ||
        Npad < N || (reinterpret_cast<uintptr_t>(dY) & 15u))
        return FEPE_E_BADARG;
    dim3 grid(Npad / 128, B);
    fepe::m32::first_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(dY, X0, W, dX0, dW, N, Npad, Ci);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_wgrad(const float* dY, const unsigned* dy_amax, const float* Yprev, const float* ss_prev, float slope,
                     float* dW, int M, int Npad, int Co, int Ci, void* stream) {
    if (!dY || !dy_amax || !Yprev || !ss_prev || !dW || M <= 0 || Npad <= 0 || (Npad % 128) != 0 || (M % Npad) != 0 ||
        (Co % 128) != 0 || (Ci % 64) != 0 || Co <= 0 || Ci <= 0 || (reinterpret_cast<uintptr_t>(ss_prev) & 15u) ||
        (reinterpret_cast<uintptr_t>(dY) & 15u) || (reinterpret_cast<uintptr_t>(Yprev) & 15u) ||
        (reinterpret_cast<uintptr_t>(dW) & 15u) || !(slope >= 0.f && slope <= 1.f))
        return FEPE_E_BADARG;
    // 128 x 256 tiles where the layer allows it: the dY tile is split once per 256 instead of 128 input channels, which
    // takes the operand transform (both operands are fp32 in memory) off the critical path of the tensor pipe
    const int wide = fepe::dispatch_get(FEPE_DISPATCH_WGRAD);                  // 0 automatic | 1 never | 2 whenever possible
    const int bn = (Ci % 256 == 0 && wide != 1) ? 256 : (Ci % 128 == 0) ? 128 : 64;
    const int tiles = (Co / 128) * (Ci / bn);
    // 2 waves of 148 SMs, never a CTA more (a third, nearly empty wave cost 1.4x on the 1024 -> 512 layer: 320 CTAs);
    // every slab pays a prologue and 128 x BN reductions
    int slabs = 296 / tiles;
    const int max_slabs = M / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    int rows_per_slab = ((M + slabs - 1) / slabs + 63) / 64 * 64;
    slabs = (M + rows_per_slab - 1) / rows_per_slab;
    fepe::m32::WgradParams p{M, Co, Ci, Npad, rows_per_slab, ss_prev, slope, dy_amax, dW};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return bn == 256   ? fepe::m32::launch_wgrad<256>(dY, Yprev, p, slabs, st)
           : bn == 128 ? fepe::m32::launch_wgrad<128>(dY, Yprev, p, slabs, st)
                       : fepe::m32::launch_wgrad<64>(dY, Yprev, p, slabs, st);
}

}  // extern "C"
