// Tensor-core path of the per-correspondence weight MLP (ErrorEstimator,
// deepFEPE/models/ErrorEstimators.py:46-64): Conv1d(k=1) == a GEMM over the channel dimension,
// followed by InstanceNorm1d(affine) (statistics over the N correspondences of ONE pair) and
// LeakyReLU(0.01).
//
//   fepe_mlp_gemm_kernel   Y[M,Co] = X[M,Ci] . W[Co,Ci]^T + b   bf16 in, fp32 accumulate in TMEM
//                          (tcgen05.mma, operands staged by TMA with 128-byte swizzle), bf16 out,
//                          plus per-(pair, channel) sum / sum of squares for the InstanceNorm that follows.
//   fepe_mlp_norm_kernel   X'[M,Co] = LeakyReLU(gamma (Y - mean) rstd + beta), bf16 (memory bound).
//   fepe_mlp_first_kernel  layer 1 (Ci = 4..8: too thin for a GEMM tile) on CUDA cores.
//   fepe_mlp_last_kernel   layer 6 (Co = 1) + softmax over the N correspondences of a pair.
//
// Rows: M = B * Npad, Npad = N rounded up to 128 so that a 128-row tile never straddles two pairs;
// padded rows are excluded from the statistics and written as zeros.
//
// Warp roles in the GEMM CTA (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/fepe_b200.h"
#include "fepe_common.cuh"

namespace fepe {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;          // 64 bf16 = 128 B = one swizzle atom
constexpr int kGemmThreads = 256;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// K-major operand tile in shared memory, 128-byte swizzle (what a TMA box of 64 bf16 x rows produces):
// start address >> 4, stride-byte-offset = 8 rows * 128 B, descriptor version 1, layout SWIZZLE_128B
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* smem_ptr) {
    const uint64_t addr = static_cast<uint64_t>(smem_u32(smem_ptr));
    // leading-byte-offset field = 1 (unused for swizzled K-major, CUTLASS sets the canonical value 1)
    return ((addr >> 4) & 0x3FFFull) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct GemmParams {
    int M, K, Co;          // M = B * Npad
    int Npad, Nvalid;      // rows per pair (padded / real)
    const float* bias;     // [Co]
    __nv_bfloat16* Y;      // [M, Co]
    float* stats;          // [B, Co, 2] (sum, sum of squares), zeroed by the caller
};

// STAGES = 4 with one CTA per SM, or STAGES = 2 with two CTAs per SM: in the second configuration the
// epilogue of one CTA (TMEM -> bf16 tile -> statistics -> store) overlaps the main loop of the other, which is
// what the thin layers (K = 64 / 128: two k-blocks of MMA, then a long epilogue) need.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, (STAGES <= 2) ? 2 : 1)
fepe_mlp_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 128-byte-swizzled operand tiles must start on a 1024-byte boundary
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kABytes = kGemmBM * kGemmBK * 2;       // 16 KB
    constexpr int kBBytes = BN * kGemmBK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    unsigned char* tile_y = smem + STAGES * kStageBytes;                     // [128][BN] bf16
    uint64_t* full = reinterpret_cast<uint64_t*>(tile_y + kGemmBM * BN * 2);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kGemmBM;     // n-tiles vary fastest: the CTAs sharing an A tile are co-scheduled
    const int n0 = blockIdx.x * BN;
    const int num_kb = p.K / kGemmBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(static_cast<uint32_t>(BN < 32 ? 32 : BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = static_cast<uint32_t>(kb / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* sa = smem + s * kStageBytes;
                mbar_arrive_expect_tx(&full[s], kStageBytes);
                tma_load_2d(sa, &map_a, kb * kGemmBK, m0, &full[s]);
                tma_load_2d(sa + kABytes, &map_w, kb * kGemmBK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D=f32, A=B=bf16, K-major both,
        // N>>3 at bit 17, M>>4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                   (static_cast<uint32_t>(kGemmBM >> 4) << 24);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = static_cast<uint32_t>(kb / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            tcgen05_fence_after();
            if (lane == 0) {                                 // one fixed thread issues every MMA and commit
                const unsigned char* sa = smem + s * kStageBytes;
                const uint64_t da = umma_desc_k_sw128(sa);
                const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
                for (int k = 0; k < kGemmBK / 16; ++k) {     // UMMA_K = 16 bf16 = 32 B: advance the start address
                    umma_bf16(tmem_base, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                              (kb | k) != 0 ? 1u : 0u);
                }
                tcgen05_commit(&empty[s]);                  // frees the stage when these MMAs have read it
                if (kb == num_kb - 1) tcgen05_commit(tmem_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---------------- epilogue: TMEM -> registers -> bf16 tile in smem ----------------
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                        // row inside the tile
        const int pair_row = (m0 % p.Npad) + row;             // row inside its pair
        const bool valid = pair_row < p.Nvalid;
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        __nv_bfloat16* ty = reinterpret_cast<__nv_bfloat16*>(tile_y);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float y0 = valid ? __uint_as_float(v[j]) + __ldg(p.bias + n0 + c + j) : 0.f;
                const float y1 = valid ? __uint_as_float(v[j + 1]) + __ldg(p.bias + n0 + c + j + 1) : 0.f;
                // column-rotated placement keeps the 32 rows of a warp off the same bank
                *reinterpret_cast<__nv_bfloat162*>(ty + row * BN + ((c + j + 2 * row) % BN)) =
                    __floats2bfloat162_rn(y0, y1);
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();

    // ---- all 256 threads: column statistics from the bf16 tile, then coalesced store ----
    {
        const __nv_bfloat16* ty = reinterpret_cast<const __nv_bfloat16*>(tile_y);
        const int pair = m0 / p.Npad;
        // every thread sums half a column (BN <= 128 columns x 2 halves = 256 threads)
        for (int item = threadIdx.x; item < 2 * BN; item += kGemmThreads) {
            const int col = item % BN, half = item / BN;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
            for (int r = half * (kGemmBM / 2); r < (half + 1) * (kGemmBM / 2); ++r) {
                const float y = __bfloat162float(ty[r * BN + ((col + 2 * r) % BN)]);
                s1 += y;
                s2 = fmaf(y, y, s2);
            }
            atomicAdd(p.stats + (static_cast<size_t>(pair) * p.Co + n0 + col) * 2, s1);
            atomicAdd(p.stats + (static_cast<size_t>(pair) * p.Co + n0 + col) * 2 + 1, s2);
        }
        // store: each row of the tile is BN*2 bytes contiguous in Y
        for (int idx = threadIdx.x; idx < kGemmBM * (BN / 2); idx += kGemmThreads) {
            const int r = idx / (BN / 2), c2 = (idx % (BN / 2)) * 2;
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(ty + r * BN + ((c2 + 2 * r) % BN));
            *reinterpret_cast<__nv_bfloat162*>(p.Y + static_cast<size_t>(m0 + r) * p.Co + n0 + c2) = v;
        }
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(static_cast<uint32_t>(BN < 32 ? 32 : BN)));
    }
}

// X'[m][c] = LeakyReLU(gamma[c] (Y[m][c] - mean[b][c]) rstd[b][c] + beta[c]); padded rows -> 0.
// One CTA per (pair, 128-row slab): the per-channel scale / shift are computed once into shared memory,
// then every thread streams 16-byte vectors (8 bf16).
__global__ void __launch_bounds__(256) fepe_mlp_norm_kernel(const __nv_bfloat16* __restrict__ Y,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ X, int Co, int Npad, int Nvalid,
                                                            float eps, float slope) {
    extern __shared__ float sc[];            // [Co] scale, [Co] shift
    float* sh = sc + Co;
    const int b = blockIdx.y;
    const int r0 = blockIdx.x * 128;
    const float invN = 1.0f / static_cast<float>(Nvalid);
    for (int c = threadIdx.x; c < Co; c += blockDim.x) {
        const float s1 = stats[(static_cast<size_t>(b) * Co + c) * 2];
        const float s2 = stats[(static_cast<size_t>(b) * Co + c) * 2 + 1];
        const float mean = s1 * invN;
        const float var = fmaxf(s2 * invN - mean * mean, 0.f);           // biased variance, like InstanceNorm1d
        const float a = rsqrtf(var + eps) * gamma[c];
        sc[c] = a;
        sh[c] = beta[c] - mean * a;
    }
    __syncthreads();
    const int vec_per_row = Co / 8;
    const size_t base = (static_cast<size_t>(b) * Npad + r0) * Co;
    for (int idx = threadIdx.x; idx < 128 * vec_per_row; idx += blockDim.x) {
        const int r = idx / vec_per_row, c = (idx % vec_per_row) * 8;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < Nvalid) {
            const uint4 in = *reinterpret_cast<const uint4*>(Y + base + static_cast<size_t>(r) * Co + c);
            const uint32_t w[4] = {in.x, in.y, in.z, in.w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
                float t0 = fmaf(__low2float(y), sc[c + 2 * k], sh[c + 2 * k]);
                float t1 = fmaf(__high2float(y), sc[c + 2 * k + 1], sh[c + 2 * k + 1]);
                t0 = t0 > 0.f ? t0 : slope * t0;
                t1 = t1 > 0.f ? t1 : slope * t1;
                const __nv_bfloat162 v = __floats2bfloat162_rn(t0, t1);
                o[k] = *reinterpret_cast<const uint32_t*>(&v);
            }
            out = make_uint4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<uint4*>(X + base + static_cast<size_t>(r) * Co + c) = out;
    }
}

// layer 1: X0 [B,N,Ci] fp32 (Ci <= 8) -> Y [B*Npad, Co=64] bf16 + stats.  One CTA per (pair, 128-row slab):
// thread = row computes its 64 outputs into a shared tile, then thread = channel sums the slab's column.
__global__ void __launch_bounds__(128) fepe_mlp_first_kernel(const float* __restrict__ X0, const float* __restrict__ W,
                                                             const float* __restrict__ bias,
                                                             __nv_bfloat16* __restrict__ Y, float* __restrict__ stats,
                                                             int B, int N, int Npad, int Ci, int Co) {
    __shared__ __nv_bfloat16 tile[128][64 + 2];
    __shared__ float w_s[64 * 8], b_s[64];
    const int b = blockIdx.y;
    const int r = blockIdx.x * 128 + threadIdx.x;
    for (int i = threadIdx.x; i < Co * Ci; i += 128) w_s[i] = W[i];
    for (int i = threadIdx.x; i < Co; i += 128) b_s[i] = bias[i];
    __syncthreads();
    float x[8];
    const bool valid = r < N;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (valid && k < Ci) ? X0[(static_cast<size_t>(b) * N + r) * Ci + k] : 0.f;
    for (int c = 0; c < Co; ++c) {
        float y = 0.f;
        if (valid) {
            y = b_s[c];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < Ci) y = fmaf(w_s[c * Ci + k], x[k], y);
        }
        tile[threadIdx.x][c] = __float2bfloat16(y);
    }
    __syncthreads();
    if (threadIdx.x < Co) {
        float s1 = 0.f, s2 = 0.f;
        for (int rr = 0; rr < 128; ++rr) {
            const float y = __bfloat162float(tile[rr][threadIdx.x]);
            s1 += y;
            s2 = fmaf(y, y, s2);
        }
        atomicAdd(stats + (static_cast<size_t>(b) * Co + threadIdx.x) * 2, s1);
        atomicAdd(stats + (static_cast<size_t>(b) * Co + threadIdx.x) * 2 + 1, s2);
    }
    for (int idx = threadIdx.x; idx < 128 * (Co / 2); idx += 128) {
        const int rr = idx / (Co / 2), c2 = (idx % (Co / 2)) * 2;
        *reinterpret_cast<__nv_bfloat162*>(Y + (static_cast<size_t>(b) * Npad + blockIdx.x * 128 + rr) * Co + c2) =
            *reinterpret_cast<const __nv_bfloat162*>(&tile[rr][c2]);
    }
}

// last layer (Ci -> 1) + softmax over the N rows of a pair.  One CTA per pair, 256 threads.
__global__ void fepe_mlp_last_kernel(const __nv_bfloat16* __restrict__ X, const float* __restrict__ W, float bias,
                                     float* __restrict__ logits, float* __restrict__ weights, int N, int Npad, int Ci) {
    extern __shared__ float sh[];            // [Npad] logits
    __shared__ float red[8];
    const int b = blockIdx.x;
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const __nv_bfloat16* x = X + (static_cast<size_t>(b) * Npad + r) * Ci;
        float acc = bias;
        for (int k = 0; k < Ci; k += 2) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + k);
            acc = fmaf(__low2float(v), W[k], acc);
            acc = fmaf(__high2float(v), W[k + 1], acc);
        }
        sh[r] = acc;
        logits[static_cast<size_t>(b) * N + r] = acc;
    }
    __syncthreads();
    float mx = -3.4e38f;
    for (int r = threadIdx.x; r < N; r += blockDim.x) mx = fmaxf(mx, sh[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
        const float e = __expf(sh[r] - mx);
        sh[r] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) sum += red[w];
    const float inv = 1.0f / sum;
    for (int r = threadIdx.x; r < N; r += blockDim.x) weights[static_cast<size_t>(b) * N + r] = sh[r] * inv;
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D bf16 row-major [rows, K] tensor, box = 64 (K) x box_rows, 128-byte swizzle
static bool make_map(CUtensorMap* map, const void* ptr, int rows, int K, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
static int launch_gemm(const void* X, const void* W, const GemmParams& p, cudaStream_t stream) {
    CUtensorMap ma, mw;
    if (!make_map(&ma, X, p.M, p.K, kGemmBM) || !make_map(&mw, W, p.Co, p.K, BN)) return FEPE_E_NODEVICE;
    constexpr int smem = STAGES * (kGemmBM * kGemmBK * 2 + BN * kGemmBK * 2) + kGemmBM * BN * 2 + 256 + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(fepe_mlp_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured = true;
    }
    dim3 grid(p.Co / BN, p.M / kGemmBM);
    fepe_mlp_gemm_kernel<BN, STAGES><<<grid, kGemmThreads, smem, stream>>>(ma, mw, p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace fepe

extern "C" {

// Y = X W^T + b (bf16 in / out, fp32 accumulate on tcgen05) with per-(pair, channel) statistics.
int fepe_mlp_gemm(const void* X, const void* W, const float* bias, void* Y, float* stats, int B, int Npad,
                  int Nvalid, int K, int Co, void* stream) {
    if (!X || !W || !bias || !Y || !stats || B <= 0 || Npad <= 0 || (Npad % fepe::kGemmBM) != 0 || Nvalid > Npad ||
        (K % fepe::kGemmBK) != 0 || (Co % 64) != 0)
        return FEPE_E_BADARG;
    fepe::GemmParams p{B * Npad, K, Co, Npad, Nvalid, bias, static_cast<__nv_bfloat16*>(Y), stats};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // 2 stages / 2 CTAs per SM: the epilogue of one CTA overlaps the main loop of the other.  Measured faster
    // than 4 stages / 1 CTA per SM on every layer (profiles/r1_mlp_timing.txt).  FEPE_MLP_STAGES=2|4 overrides.
    int stages = 2;
    if (const char* ev = getenv("FEPE_MLP_STAGES")) stages = (ev[0] == '2') ? 2 : 4;
    if (Co % 128 == 0)
        return stages == 2 ? fepe::launch_gemm<128, 2>(X, W, p, st) : fepe::launch_gemm<128, 4>(X, W, p, st);
    return stages == 2 ? fepe::launch_gemm<64, 2>(X, W, p, st) : fepe::launch_gemm<64, 4>(X, W, p, st);
}

int fepe_mlp_norm(const void* Y, const float* stats, const float* gamma, const float* beta, void* X, int B, int Npad,
                  int Nvalid, int Co, float eps, float slope, void* stream) {
    if (!Y || !stats || !gamma || !beta || !X || B <= 0 || (Co & 7) || (Npad % 128) != 0) return FEPE_E_BADARG;
    dim3 grid(Npad / 128, B);
    fepe::fepe_mlp_norm_kernel<<<grid, 256, 2 * Co * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(Y), stats, gamma, beta, static_cast<__nv_bfloat16*>(X), Co, Npad, Nvalid, eps,
        slope);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp_first(const float* X0, const float* W, const float* bias, void* Y, float* stats, int B, int N, int Npad,
                   int Ci, int Co, void* stream) {
    if (!X0 || !W || !bias || !Y || !stats || B <= 0 || Ci <= 0 || Ci > 8 || Co != 64 || (Npad % 128) != 0) return FEPE_E_BADARG;
    dim3 grid(Npad / 128, B);
    fepe::fepe_mlp_first_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        X0, W, bias, static_cast<__nv_bfloat16*>(Y), stats, B, N, Npad, Ci, Co);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp_last(const void* X, const float* W, float bias, float* logits, float* weights, int B, int N, int Npad,
                  int Ci, void* stream) {
    if (!X || !W || !logits || !weights || B <= 0 || (Ci & 1)) return FEPE_E_BADARG;
    fepe::fepe_mlp_last_kernel<<<B, 256, Npad * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(X), W, bias, logits, weights, N, Npad, Ci);
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
