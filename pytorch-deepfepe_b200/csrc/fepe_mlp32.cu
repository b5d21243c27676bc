// fp32-parity tensor-core path of the per-correspondence weight MLP (ErrorEstimator,
// deepFEPE/models/ErrorEstimators.py:46-64; the reference computes it in fp32).
//
// The 1x1 convolutions are GEMMs over the channel dimension.  tcgen05 has no fp32 operand type, so every operand is
// SPLIT into two fp16 numbers, x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (11 + 11 significant bits), and a
// product is formed from three tensor-core MMAs accumulated in fp32 in tensor memory:
//        x w  ~=  hi_x hi_w + lo_x hi_w + hi_x lo_w          (dropped: lo_x lo_w <= 2^-22 |x w|)
// i.e. each product carries a relative error of ~2^-21, the accuracy class of an fp32 FMA chain (2^-24 per step over a
// K-long sum), not of bf16 (2^-9).  fp16 (not tf32) because kind::f16 runs at twice the tf32 rate with half the
// operand bytes, and its 5-bit exponent is enough here: the A operand is the post-InstanceNorm activation (|x'| <=
// |gamma| sqrt(N) + |beta|), the weights are pre-multiplied by a power of two chosen from max |W| (undone exactly in
// the epilogue), and the conversions saturate instead of overflowing.
//
//   fepe_mlp32_prepare_weights   W fp32 [Co,K] -> (W_hi, W_lo) fp16 + the power-of-two scale
//   fepe_mlp32_first             layer 1 (Ci <= 16: too thin for a tile) on CUDA cores, reading the model's inputs in
//                                place: the pixel matches (affine + (x+1)/2 applied on load, DeepFNet.get_input
//                                :377-389) and up to four extra channel groups (quality; weights, epi_res, residual
//                                of the previous iteration, DeepFNet.py:487) -- no torch.cat / permute in front of it
//   fepe_mlp32_gemm              Y[M,Co] = act(Yprev)[M,K] . W[Co,K]^T (+ b), fp32 in HBM on both sides.  act = the
//                                PREVIOUS block's InstanceNorm + LeakyReLU, applied to the operand tile in shared memory
//                                by transform warps that also produce the hi / lo fp16 tiles IN PLACE of the fp32 tile
//                                the TMA delivered; persistent CTAs, accumulator double-buffered in TMEM; the epilogue
//                                emits per-(pair, channel) sum / sum of squares in fp64 (pivoted, so the variance does
//                                not cancel) for the InstanceNorm that follows.  With ss = NULL the operand is used as
//                                is: the data-gradient GEMM of the backward pass.
//   fepe_mlp32_scale_shift       statistics -> per-(pair, channel) (a, d), x' = LeakyReLU(a y + d), in fp64
//   fepe_mlp32_last              last block's norm + final Conv1d (Co = 1: + softmax over N, DeepFNet.py:443,512;
//                                Co = 4: the offsets network, DeepFNet.py:341-342) on CUDA cores
//
// Rows: M = B * Npad, Npad = N rounded up to 128 so that a 128-row tile never straddles two pairs; padded rows are
// written as zeros and excluded from the statistics.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_common.cuh"
#include "fepe_split.cuh"
#include "fepe_umma.cuh"

namespace fepe {
namespace m32 {

constexpr int kBM = 128;
constexpr int kBK = 64;                       // channels per k-block: 64 fp16 = 128 B = one swizzle atom = two fp32 boxes of 32
constexpr int kThreads = 512;
constexpr int kABytes = kBM * kBK * 4;        // 32 KB: the fp32 tile as delivered == hi tile (16 KB) + lo tile (16 KB) after the split
constexpr int kHalfA = kABytes / 2;
constexpr int kPassCols = 32;                 // accumulator columns per epilogue pass
constexpr int kTileBytes = kBM * kPassCols * 4;      // 16 KB
constexpr int kSsBytes = kBK * 8;             // (a, d) fp32 of the 64 channels of a k-block

struct GemmParams {
    int M, K, Co;            // M = B * Npad
    int Npad, Nvalid;        // rows per pair (padded / real)
    const float* ss;         // [B, K, 2] = (a_c, d_c), x' = LeakyReLU(a y + d); null: operand used as is
    float slope;
    const float* wscale;     // [>= 2]: (s, 1/s), the power of two the weights were multiplied by
    const float* bias;       // [Co] or null (a bias in front of an InstanceNorm cancels in the normalisation)
    float* Y;                // [M, Co]
    double* stats;           // [B, Co, 2] (sum, sum of squares), zeroed by the caller; or null
    const unsigned* a_amax;  // ss == null only: bits of max |A| (device), the operand is pre-multiplied by the power of
                             // two that puts it at 2^13..2^14 (gradients can be far below fp16's range); or null
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(128) : "memory"); }

// A 128 x 32 fp32 tile in shared memory as 16-byte chunks of 4 columns, chunk k of row r at slot (k + r) mod 8: a quarter
// warp (8 consecutive rows of one chunk, or the 8 chunks of one row) always touches 8 different slots = all 32 banks once.
__device__ __forceinline__ unsigned char* tile_chunk(unsigned char* tile, int r, int k) {
    return tile + (r * 8 + ((k + r) & 7)) * 16;
}

// Column statistics of such a tile, called by the four warps (wq = 0..3) of a 128-thread group.  Warp wq owns chunks
// 2 wq and 2 wq + 1; lane = (chunk half, row residue rg = lane % 16) reads rows rg, rg + 16, ... of its chunk.  The sums
// are PIVOTED on the tile's first row (always a real row): T1 = sum (y - c), T2 = sum (y - c)^2 have no cancellation in
// fp32, and sum y = T1 + n c, sum y^2 = T2 + 2 c T1 + n c^2 are formed in fp64 -- so mean^2 can be subtracted from
// E[y^2] downstream (fepe_mlp32_scale_shift) without losing the variance when |mean| >> std.  A reduce-scatter over the
// row residues leaves (T1, T2) of the chunk's four columns in lanes rg = 0..7; those lanes issue one fp64 atomic each.
__device__ __forceinline__ void tile_stats(unsigned char* tile, int wq, int lane, int rows_valid, double* stats32) {
    const int kc = 2 * wq + (lane >> 4);
    const int rg = lane & 15;
    const float4 piv = *reinterpret_cast<const float4*>(tile_chunk(tile, 0, kc));
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.f;
#pragma unroll
    for (int i = 0; i < kBM / 16; ++i) {
        const int r = rg + 16 * i;
        if (r < rows_valid) {
            const float4 u = *reinterpret_cast<const float4*>(tile_chunk(tile, r, kc));
            const float d0 = u.x - piv.x, d1 = u.y - piv.y, d2 = u.z - piv.z, d3 = u.w - piv.w;
            a[0] += d0; a[1] = fmaf(d0, d0, a[1]);
            a[2] += d1; a[3] = fmaf(d1, d1, a[3]);
            a[4] += d2; a[5] = fmaf(d2, d2, a[5]);
            a[6] += d3; a[7] = fmaf(d3, d3, a[7]);
        }
    }
#pragma unroll
    for (int h = 4; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? a[i] : a[i + h];
            const float keep = up ? a[i + h] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    const float t = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 8);     // value index rg & 7 = 2 * column + (0: T1, 1: T2)
    const float other = __shfl_xor_sync(0xffffffffu, t, 1);
    const int v = rg & 7, j = v >> 1;
    const float c = (j == 0) ? piv.x : (j == 1) ? piv.y : (j == 2) ? piv.z : piv.w;
    const double n = static_cast<double>(rows_valid), cd = static_cast<double>(c);
    const double out = (v & 1) ? static_cast<double>(t) + 2.0 * cd * static_cast<double>(other) + n * cd * cd
                               : static_cast<double>(t) + n * cd;
    if (rg < 8) atomicAdd(stats32 + (kc * 4 + j) * 2 + (v & 1), out);
}

// Coalesced copy of the tile to Y (row pitch ld floats): thread et of the group moves 8 chunks, 8 consecutive threads
// one 128-byte row segment.
__device__ __forceinline__ void tile_store(unsigned char* tile, int et, float* y, int ld) {
#pragma unroll
    for (int idx = et; idx < kBM * 8; idx += 128) {
        const int r = idx >> 3, k = idx & 7;
        *reinterpret_cast<float4*>(y + static_cast<size_t>(r) * ld + k * 4) =
            *reinterpret_cast<const float4*>(tile_chunk(tile, r, k));
    }
}

// ------------------------------------------------------------------------------------------------
// 16 warps: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator, warps 4-11 = two
// epilogue groups of four warps (TMEM lane quarter = warp % 4; group g drains the columns [g BN/2, (g+1) BN/2) of the
// accumulator in passes of 32 through its own 16 KB tile), warps 12-15 = operand transform, one THREAD PER ROW of the
// 128 x 64 A tile.
//
// A stage holds [A box 0 (channels 0..31) | A box 1 (32..63)] as fp32 rows of 128 bytes (two TMA boxes, 128-byte swizzle)
// and [W_hi | W_lo] (BN rows of 64 fp16 = 128 bytes).  The fp16 hi tile of the transformed operand has the same
// geometry as A box 0 (128 rows of 128 swizzled bytes) and the lo tile that of A box 1, and the swizzle only depends on
// (row & 7): the transform thread of row r reads its 2 x 128 bytes, normalises, splits and writes the hi / lo rows
// back over the very bytes it read -- in place, no second buffer, no cross-thread hazard.
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
fepe_mlp32_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_wh,
                       const __grid_constant__ CUtensorMap map_wl, const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kWBytes = BN * kBK * 2;                // one of W_hi / W_lo
    constexpr int kStageBytes = kABytes + 2 * kWBytes;
    constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;     // 128, 256 or 512: a power of two
    constexpr int kPasses = BN / 2 / kPassCols;          // epilogue passes per group and tile
    unsigned char* epi_tiles = smem + STAGES * kStageBytes;          // [2 groups][128][32] fp32
    unsigned char* ss_ring = epi_tiles + 2 * kTileBytes;             // [STAGES][512 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(ss_ring + STAGES * kSsBytes);
    uint64_t* empty = full + STAGES;
    uint64_t* ready = empty + STAGES;                    // the A tile of the stage has been transformed
    uint64_t* tmem_full = ready + STAGES;                // [2]
    uint64_t* tmem_empty = tmem_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = p.K / kBK;
    const int n_tiles = p.Co / BN;
    const int tiles = (p.M / kBM) * n_tiles;
    const int n_local = (tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);
    const bool has_ss = p.ss != nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer: the ring runs across tile boundaries ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int j = 0; j < n_local; ++j) {
                const int t = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
                const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * BN;
                const int pair = m0 / p.Npad;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* sa = smem + s * kStageBytes;
                    mbar_arrive_expect_tx(&full[s], kStageBytes + (has_ss ? kSsBytes : 0));
                    tma_load_2d(sa, &map_a, kb * kBK, m0, &full[s]);
                    tma_load_2d(sa + kHalfA, &map_a, kb * kBK + 32, m0, &full[s]);
                    tma_load_2d(sa + kABytes, &map_wh, kb * kBK, n0, &full[s]);
                    tma_load_2d(sa + kABytes + kWBytes, &map_wl, kb * kBK, n0, &full[s]);
                    if (has_ss)
                        bulk_g2s(ss_ring + s * kSsBytes, p.ss + (static_cast<size_t>(pair) * p.K + kb * kBK) * 2, kSsBytes,
                                 &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: three products per 16-channel step ----------------
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp: InstrDescriptor): D = f32 (bit 4), A = B = f16 (format 0),
        // both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(kBM >> 4) << 24);
        uint32_t it = 0;
        for (int j = 0; j < n_local; ++j) {
            const uint32_t b = static_cast<uint32_t>(j) & 1u;
            const uint32_t use = static_cast<uint32_t>(j) >> 1;                // how often buffer b was used before
            mbar_wait(&tmem_empty[b], (use & 1u) ^ 1u);                        // drained by the epilogue (free at first use)
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + b * static_cast<uint32_t>(BN);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(&ready[s], ph);
                tcgen05_fence_after();
                if (lane == 0) {
                    const unsigned char* sa = smem + s * kStageBytes;
                    const uint64_t ah = umma_desc_k_sw128(sa), al = umma_desc_k_sw128(sa + kHalfA);
                    const uint64_t wh = umma_desc_k_sw128(sa + kABytes), wl = umma_desc_k_sw128(sa + kABytes + kWBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {               // UMMA_K = 16 fp16 = 32 B: advance the start address
                        const uint64_t o = static_cast<uint64_t>(k * 2);
                        umma_bf16(tmem_d, ah + o, wh + o, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_bf16(tmem_d, al + o, wh + o, idesc, 1u);
                        umma_bf16(tmem_d, ah + o, wl + o, idesc, 1u);
                    }
                    tcgen05_commit(&empty[s]);
                    if (kb == num_kb - 1) tcgen05_commit(&tmem_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 12) {
        // ---------------- operand transform: norm + LeakyReLU of the previous block, hi / lo split ----------------
        const int row = static_cast<int>(threadIdx.x) - 384;
        const int sw = row & 7;
        const float slope = p.slope;
        const float a_scale = (!has_ss && p.a_amax != nullptr) ? pow2_scale(__ldg(p.a_amax)) : 1.f;
        uint32_t it = 0;
        for (int j = 0; j < n_local; ++j) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                unsigned char* a0 = smem + s * kStageBytes + row * 128;       // row of box 0 == row of the hi tile
                unsigned char* a1 = a0 + kHalfA;                               // row of box 1 == row of the lo tile
                float v[kBK];
#pragma unroll
                for (int c = 0; c < 8; ++c) {                                 // logical chunk c = channels 4c .. 4c+3 of the box
                    const float4 t0 = *reinterpret_cast<const float4*>(a0 + ((c ^ sw) << 4));
                    const float4 t1 = *reinterpret_cast<const float4*>(a1 + ((c ^ sw) << 4));
                    v[4 * c + 0] = t0.x; v[4 * c + 1] = t0.y; v[4 * c + 2] = t0.z; v[4 * c + 3] = t0.w;
                    v[32 + 4 * c + 0] = t1.x; v[32 + 4 * c + 1] = t1.y; v[32 + 4 * c + 2] = t1.z; v[32 + 4 * c + 3] = t1.w;
                }
                if (has_ss) {
                    const float4* cf = reinterpret_cast<const float4*>(ss_ring + s * kSsBytes);   // (a_c, d_c, a_c+1, d_c+1)
#pragma unroll
                    for (int q = 0; q < kBK / 2; ++q) {
                        const float4 c4 = cf[q];                              // same address in every lane: broadcast
                        const float t0 = fmaf(v[2 * q], c4.x, c4.y), t1 = fmaf(v[2 * q + 1], c4.z, c4.w);
                        v[2 * q] = fmaxf(t0, slope * t0);
                        v[2 * q + 1] = fmaxf(t1, slope * t1);
                    }
                } else if (a_scale != 1.f) {
#pragma unroll
                    for (int q = 0; q < kBK; ++q) v[q] *= a_scale;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {                                 // fp16 chunk c = channels 8c .. 8c+7
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) split2(v[8 * c + 2 * q], v[8 * c + 2 * q + 1], h[q], l[q]);
                    *reinterpret_cast<uint4*>(a0 + ((c ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(a1 + ((c ^ sw) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
                }
                fence_proxy_async();                            // generic-proxy writes -> visible to the MMA's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[s]);
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue groups ----------------
        const int ew = warp - 4;
        const int g = ew >> 2;                                // column half of the accumulator
        const int wq = ew & 3;                                // warp inside the group == TMEM lane quarter (warp % 4)
        const int row = wq * 32 + lane;
        const int et = wq * 32 + lane;                        // thread inside the group
        unsigned char* tile = epi_tiles + g * kTileBytes;
        const bool has_bias = p.bias != nullptr;
        const float inv_scale = __ldg(p.wscale + 1) / ((!has_ss && p.a_amax != nullptr) ? pow2_scale(__ldg(p.a_amax)) : 1.f);
        for (int j = 0; j < n_local; ++j) {
            const int t = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
            const int m0 = (t / n_tiles) * kBM, n0 = (t % n_tiles) * BN;
            const uint32_t b = static_cast<uint32_t>(j) & 1u;
            const uint32_t use = static_cast<uint32_t>(j) >> 1;
            const int pair = m0 / p.Npad;
            int rows_valid = p.Nvalid - (m0 - pair * p.Npad);
            rows_valid = rows_valid > kBM ? kBM : (rows_valid < 0 ? 0 : rows_valid);
            const bool valid = row < rows_valid;
            mbar_wait(&tmem_full[b], use & 1u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int pass = 0; pass < kPasses; ++pass) {
                const int col0 = g * (BN / 2) + pass * kPassCols;          // first accumulator column of this pass
                uint32_t v[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + b * static_cast<uint32_t>(BN) +
                              static_cast<uint32_t>(col0), v);
                if (pass == kPasses - 1) {                    // last read of this accumulator buffer: hand it back
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[b]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 y = make_float4(__uint_as_float(v[4 * k]) * inv_scale, __uint_as_float(v[4 * k + 1]) * inv_scale,
                                           __uint_as_float(v[4 * k + 2]) * inv_scale, __uint_as_float(v[4 * k + 3]) * inv_scale);
                    if (has_bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col0) + k);   // 16-byte aligned
                        y.x += bb.x; y.y += bb.y; y.z += bb.z; y.w += bb.w;
                    }
                    if (!valid) y = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(tile_chunk(tile, row, k)) = y;
                }
                group_sync(g);
                if (p.stats != nullptr)
                    tile_stats(tile, wq, lane, rows_valid, p.stats + (static_cast<size_t>(pair) * p.Co + n0 + col0) * 2);
                tile_store(tile, et, p.Y + static_cast<size_t>(m0) * p.Co + n0 + col0, p.Co);
                group_sync(g);                                // the tile is rewritten by the next pass
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ------------------------------------------------------------------------------------------------
// W [Co, K] fp32 -> W_hi, W_lo fp16 of W * s, s = the power of two that puts max |W| into [2^13, 2^14): the lo parts
// are then normal fp16 numbers (not subnormals) for every weight down to 2^-17 of the largest, and nothing overflows.
// wsc[0] = s, wsc[1] = 1/s, wsc[2] = bits of max |W| (zeroed by the launcher).
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ W, size_t n, float* __restrict__ wsc) {
    float m = 0.f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float a = fabsf(W[i]);
        if (a < 3.0e38f) m = fmaxf(m, a);                 // ignore inf / nan: they saturate later
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(wsc) + 2, __float_as_uint(m));
}

__global__ void __launch_bounds__(256) split_weights_kernel(const float* __restrict__ W, size_t n, __half* __restrict__ Whi,
                                                            __half* __restrict__ Wlo, float* __restrict__ wsc) {
    const float s = pow2_scale(reinterpret_cast<const unsigned int*>(wsc)[2]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { wsc[0] = s; wsc[1] = 1.f / s; }
    for (size_t i = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) * 2; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x * 2) {
        const float x0 = W[i] * s, x1 = (i + 1 < n) ? W[i + 1] * s : 0.f;
        uint32_t h, l;
        split2(x0, x1, h, l);
        if (i + 1 < n) {
            *reinterpret_cast<uint32_t*>(Whi + i) = h;
            *reinterpret_cast<uint32_t*>(Wlo + i) = l;
        } else {
            Whi[i] = *reinterpret_cast<const __half*>(&h);
            Wlo[i] = *reinterpret_cast<const __half*>(&l);
        }
    }
}

// (a, d) of the fused norm: x' = LeakyReLU(a y + d), a = gamma rstd, d = beta - mean a, from the fp64 statistics
// (biased variance like InstanceNorm1d).  One thread per (pair, channel); ss[b][c] = (a_c, d_c).  With `clear` the
// statistics are zeroed for the next accumulation.
__global__ void __launch_bounds__(256) scale_shift_kernel(double* __restrict__ stats, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float2* __restrict__ ss,
                                                          float2* __restrict__ mr, int total, int Co, int Nvalid, float eps,
                                                          int clear) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = i % Co;
    double2* sp = reinterpret_cast<double2*>(stats) + i;
    const double2 v = *sp;
    const double invN = 1.0 / static_cast<double>(Nvalid);
    const double mean = v.x * invN;
    double var = v.y * invN - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double a = static_cast<double>(gamma[c]) * rstd;
    ss[i] = make_float2(static_cast<float>(a), static_cast<float>(static_cast<double>(beta[c]) - mean * a));
    if (mr != nullptr) mr[i] = make_float2(static_cast<float>(mean), static_cast<float>(rstd));   // for the backward pass
    if (clear) *sp = make_double2(0.0, 0.0);
}

// ------------------------------------------------------------------------------------------------
// Layer 1: the model's inputs, read in place, -> Y [B*Npad, 64] fp32 + statistics.  One CTA of 128 threads per (pair,
// 128-row slab): thread = row.  Channels, in the order of the reference's torch.cat (DeepFNet.py:387-389, :487):
//   matches != null: ((ax x1 + bx) + 1) / 2, ((ay y1 + by) + 1) / 2, the same for (x2, y2)      [4]
//   then extra[0..3], each [B, N, ec[i]] (quality; weights; epi_res; residual -- or any fp32 features)
// Weights transposed in shared memory and read as broadcast vectors; the tile's column statistics and coalesced stores
// are the GEMM epilogue's own (two passes of 32 columns).
constexpr int kFirstMaxCi = 16;
struct FirstParams {
    const float* matches;      // [B,N,4] or null
    float ax, bx, ay, by;
    const float* extra[4];
    int ec[4];
    const float* W;            // [64, Ci]
    const float* bias;         // [64] or null
    float* Y;
    double* stats;
    float* X0_out;             // [B,N,Ci] the assembled input features (kept for the backward pass) or null
    int B, N, Npad, Ci;
};

__global__ void __launch_bounds__(128) fepe_mlp32_first_kernel(const FirstParams p) {
    __shared__ __align__(16) unsigned char tile[kTileBytes];
    __shared__ __align__(16) float w_t[kFirstMaxCi * 64];       // [k][c]: W^T
    __shared__ __align__(16) float b_s[64];
    const int b = blockIdx.y;
    const int row = threadIdx.x;
    const int r = blockIdx.x * 128 + row;
    const int Ci = p.Ci;
    for (int i = threadIdx.x; i < kFirstMaxCi * 64; i += 128) {
        const int k = i >> 6, c = i & 63;
        w_t[i] = (k < Ci) ? p.W[c * Ci + k] : 0.f;
    }
    if (threadIdx.x < 64) b_s[threadIdx.x] = p.bias != nullptr ? p.bias[threadIdx.x] : 0.f;
    float x[kFirstMaxCi];
#pragma unroll
    for (int k = 0; k < kFirstMaxCi; ++k) x[k] = 0.f;
    const bool valid = r < p.N;
    if (valid) {
        const size_t pt = static_cast<size_t>(b) * p.N + r;
        int c = 0;
        if (p.matches != nullptr) {
            const float4 m = *reinterpret_cast<const float4*>(p.matches + pt * 4);
            x[0] = (fmaf(p.ax, m.x, p.bx) + 1.f) * 0.5f;
            x[1] = (fmaf(p.ay, m.y, p.by) + 1.f) * 0.5f;
            x[2] = (fmaf(p.ax, m.z, p.bx) + 1.f) * 0.5f;
            x[3] = (fmaf(p.ay, m.w, p.by) + 1.f) * 0.5f;
            c = 4;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (p.extra[e] == nullptr) continue;
            for (int q = 0; q < p.ec[e]; ++q) {
                const float val = p.extra[e][pt * p.ec[e] + q];
#pragma unroll
                for (int k = 0; k < kFirstMaxCi; ++k)
                    if (k == c) x[k] = val;              // static register indices
                ++c;
            }
        }
    }
    if (valid && p.X0_out != nullptr) {
        float* xo = p.X0_out + (static_cast<size_t>(b) * p.N + r) * Ci;
#pragma unroll
        for (int k = 0; k < kFirstMaxCi; ++k)
            if (k < Ci) xo[k] = x[k];
    }
    __syncthreads();
    int rows_valid = p.N - blockIdx.x * 128;
    rows_valid = rows_valid > 128 ? 128 : rows_valid;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {                         // 4 channels = one 16-byte chunk of the tile
            const int c = pass * 32 + k4 * 4;
            float4 y = *reinterpret_cast<const float4*>(b_s + c);
#pragma unroll
            for (int k = 0; k < kFirstMaxCi; ++k) {              // zero weights beyond Ci
                const float4 w4 = *reinterpret_cast<const float4*>(w_t + k * 64 + c);
                y.x = fmaf(w4.x, x[k], y.x); y.y = fmaf(w4.y, x[k], y.y);
                y.z = fmaf(w4.z, x[k], y.z); y.w = fmaf(w4.w, x[k], y.w);
            }
            if (!valid) y = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(tile_chunk(tile, row, k4)) = y;
        }
        __syncthreads();
        tile_stats(tile, static_cast<int>(threadIdx.x) >> 5, static_cast<int>(threadIdx.x) & 31, rows_valid,
                   p.stats + (static_cast<size_t>(b) * 64 + pass * 32) * 2);
        tile_store(tile, static_cast<int>(threadIdx.x), p.Y + (static_cast<size_t>(b) * p.Npad + blockIdx.x * 128) * 64 + pass * 32, 64);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Last block: x' = LeakyReLU(a y + d) of the PRE-norm output Y [B*Npad, Ci] (fp32), logits[b, o, n] = x' . W[o, :] + bias[o]
// for CO outputs, and for CO == 1 the softmax over the N rows of the pair (DeepFNet.py:443,512).  A pair is spread over the
// S = gridDim.x CTAs of one thread-block cluster (S = 1 for batches that fill the machine on their own, up to 8 for the 16
// pairs of a training step): every CTA owns a contiguous range of rows -- a warp per row, lane = 4 consecutive channels
// of every 128 (coalesced 512-byte reads), (a, d) and W of the lane's channels in registers -- keeps its logits in shared
// memory, and the softmax's max and sum are exchanged through distributed shared memory (two cluster barriers).
template <int CO, int CI>
__global__ void __launch_bounds__(256) fepe_mlp32_last_kernel(const float* __restrict__ Y, const float2* __restrict__ ss,
                                                              float slope, const float* __restrict__ W,
                                                              const float* __restrict__ bias, float* __restrict__ logits,
                                                              float* __restrict__ weights, int N, int Npad) {
    namespace cg = cooperative_groups;
    extern __shared__ float sh[];            // [rows of this CTA] logits (CO == 1)
    __shared__ float red[8];
    __shared__ float part[2];                // this CTA's max / sum of exponentials, read by the whole cluster
    constexpr int KPL = CI / 128;            // channel groups of 4 per lane
    const int b = blockIdx.y;
    const int S = gridDim.x;
    const int chunk = (((N + S - 1) / S) + 1) & ~1;
    const int n0 = blockIdx.x * chunk;
    const int n1 = (n0 + chunk < N) ? n0 + chunk : N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float a[KPL][4], d[KPL][4], w[CO][KPL][4];
#pragma unroll
    for (int g = 0; g < KPL; ++g) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = g * 128 + lane * 4 + q;
            const float2 s2 = __ldg(ss + static_cast<size_t>(b) * CI + c);
            a[g][q] = s2.x; d[g][q] = s2.y;
#pragma unroll
            for (int o = 0; o < CO; ++o) w[o][g][q] = __ldg(W + o * CI + c);
        }
    }
    float bs[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) bs[o] = bias != nullptr ? __ldg(bias + o) : 0.f;
    const float* yb = Y + static_cast<size_t>(b) * Npad * CI;
    for (int r0 = n0 + warp * 2; r0 < n1; r0 += nwarp * 2) {
        float4 in[2][KPL];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int g = 0; g < KPL; ++g) {
                in[u][g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r0 + u < n1) in[u][g] = __ldcs(reinterpret_cast<const float4*>(yb + static_cast<size_t>(r0 + u) * CI + g * 128 + lane * 4));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float acc[CO];
#pragma unroll
            for (int o = 0; o < CO; ++o) acc[o] = 0.f;
#pragma unroll
            for (int g = 0; g < KPL; ++g) {
                const float yv[4] = {in[u][g].x, in[u][g].y, in[u][g].z, in[u][g].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float t = fmaf(yv[q], a[g][q], d[g][q]);
                    const float xq = fmaxf(t, slope * t);
#pragma unroll
                    for (int o = 0; o < CO; ++o) acc[o] = fmaf(xq, w[o][g][q], acc[o]);
                }
            }
#pragma unroll
            for (int o = 0; o < CO; ++o) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
            }
            if (lane == 0 && r0 + u < n1) {
#pragma unroll
                for (int o = 0; o < CO; ++o) {
                    const float val = acc[o] + bs[o];
                    logits[(static_cast<size_t>(b) * CO + o) * N + r0 + u] = val;
                    if (CO == 1) sh[r0 + u - n0] = val;
                }
            }
        }
    }
    if (CO != 1 || weights == nullptr) return;          // uniform over the grid: no CTA waits for one that left
    cg::cluster_group cluster = cg::this_cluster();
    const int rows = n1 > n0 ? n1 - n0 : 0;
    __syncthreads();
    float mx = -3.4e38f;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) mx = fmaxf(mx, sh[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        mx = red[0];
        for (int i = 1; i < nwarp; ++i) mx = fmaxf(mx, red[i]);
        part[0] = mx;
    }
    cluster.sync();
    mx = -3.4e38f;
    for (int r = 0; r < S; ++r) mx = fmaxf(mx, *cluster.map_shared_rank(&part[0], r));
    float sum = 0.f;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const float e = expf(sh[r] - mx);
        sh[r] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        sum = 0.f;
        for (int i = 0; i < nwarp; ++i) sum += red[i];
        part[1] = sum;
    }
    cluster.sync();
    sum = 0.f;
    for (int r = 0; r < S; ++r) sum += *cluster.map_shared_rank(&part[1], r);
    const float inv = 1.0f / sum;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) weights[static_cast<size_t>(b) * N + n0 + r] = sh[r] * inv;
    cluster.sync();                                      // nobody's shared memory goes away while a peer still reads it
}

template <int BN, int STAGES>
static int launch_gemm(const void* X, const void* Whi, const void* Wlo, const GemmParams& p, cudaStream_t stream) {
    CUtensorMap ma, mh, ml;
    if (!make_map_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, X, p.M, p.K, 32, kBM) ||
        !make_map_2d(&mh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Whi, p.Co, p.K, kBK, BN) ||
        !make_map_2d(&ml, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, Wlo, p.Co, p.K, kBK, BN))
        return FEPE_E_NODEVICE;
    constexpr int smem = STAGES * (kABytes + 2 * BN * kBK * 2) + 2 * kTileBytes + STAGES * kSsBytes + 256 + 1024;
    static_assert(smem <= 232448, "stage ring does not fit the 227 KB of shared memory a CTA may use");
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sms[dev & 63] == 0) {
        cudaError_t e = cudaFuncSetAttribute(fepe_mlp32_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || n <= 0) return FEPE_E_NODEVICE;
        sms[dev & 63] = n;
    }
    const int tiles = (p.M / kBM) * (p.Co / BN);
    const int grid = tiles < sms[dev & 63] ? tiles : sms[dev & 63];
    fepe_mlp32_gemm_kernel<BN, STAGES><<<grid, kThreads, smem, stream>>>(ma, mh, ml, p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace m32
}  // namespace fepe

extern "C" {

int fepe_mlp32_prepare_weights(const float* W, void* Whi, void* Wlo, float* wscale, int Co, int K, void* stream) {
    if (!W || !Whi || !Wlo || !wscale || Co <= 0 || K <= 0 || (reinterpret_cast<uintptr_t>(Whi) & 3u) ||
        (reinterpret_cast<uintptr_t>(Wlo) & 3u))
        return FEPE_E_BADARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n = static_cast<size_t>(Co) * K;
    cudaError_t e = cudaMemsetAsync(wscale, 0, 4 * sizeof(float), st);
    if (e != cudaSuccess) return static_cast<int>(e);
    int blocks = static_cast<int>((n + 255) / 256);
    blocks = blocks > 592 ? 592 : blocks;
    fepe::m32::absmax_kernel<<<blocks, 256, 0, st>>>(W, n, wscale);
    fepe::m32::split_weights_kernel<<<blocks, 256, 0, st>>>(W, n, static_cast<__half*>(Whi), static_cast<__half*>(Wlo), wscale);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_first(const float* matches, float ax, float bx, float ay, float by, const float* extra0, int c0,
                     const float* extra1, int c1, const float* extra2, int c2, const float* extra3, int c3,
                     const float* W, const float* bias, float* Y, double* stats, float* X0_out, int B, int N, int Npad,
                     int Co, void* stream) {
    const int Ci = (matches ? 4 : 0) + (extra0 ? c0 : 0) + (extra1 ? c1 : 0) + (extra2 ? c2 : 0) + (extra3 ? c3 : 0);
    if (!W || !Y || !stats || B <= 0 || N <= 0 || Ci <= 0 || Ci > fepe::m32::kFirstMaxCi || Co != 64 || (Npad % 128) != 0 ||
        Npad < N || c0 < 0 || c1 < 0 || c2 < 0 || c3 < 0 || (reinterpret_cast<uintptr_t>(matches) & 15u))
        return FEPE_E_BADARG;
    fepe::m32::FirstParams p{matches, ax, bx, ay, by, {extra0, extra1, extra2, extra3}, {c0, c1, c2, c3}, W, bias, Y, stats,
                             X0_out, B, N, Npad, Ci};
    dim3 grid(Npad / 128, B);
    fepe::m32::fepe_mlp32_first_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_scale_shift(double* stats, const float* gamma, const float* beta, float* ss, float* mean_rstd, int B, int Co,
                           int Nvalid, float eps, int clear_stats, void* stream) {
    if (!stats || !gamma || !beta || !ss || B <= 0 || Co <= 0 || Nvalid <= 0 || (reinterpret_cast<uintptr_t>(stats) & 15u) ||
        (reinterpret_cast<uintptr_t>(ss) & 15u))
        return FEPE_E_BADARG;
    const int total = B * Co;
    fepe::m32::scale_shift_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, gamma, beta, reinterpret_cast<float2*>(ss), reinterpret_cast<float2*>(mean_rstd), total, Co, Nvalid, eps,
        clear_stats);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp32_gemm(const float* Yprev, const float* ss, float slope, const unsigned* a_amax, const void* Whi,
                    const void* Wlo, const float* wscale, const float* bias, float* Y, double* stats, int B, int Npad,
                    int Nvalid, int K, int Co, void* stream) {
    using namespace fepe::m32;
    if (!Yprev || !Whi || !Wlo || !wscale || !Y || B <= 0 || Npad <= 0 || (Npad % kBM) != 0 || Nvalid > Npad || Nvalid <= 0 ||
        K <= 0 || (K % kBK) != 0 || Co <= 0 || (Co % 64) != 0 || (reinterpret_cast<uintptr_t>(ss) & 15u) ||
        (reinterpret_cast<uintptr_t>(bias) & 15u) || (reinterpret_cast<uintptr_t>(Y) & 15u) ||
        (reinterpret_cast<uintptr_t>(Yprev) & 15u) || (ss != nullptr && !(slope >= 0.f && slope <= 1.f)))
        return FEPE_E_BADARG;
    GemmParams p{B * Npad, K, Co, Npad, Nvalid, ss, slope, wscale, bias, Y, stats, ss == nullptr ? a_amax : nullptr};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (Co % 256 == 0) return launch_gemm<256, 2>(Yprev, Whi, Wlo, p, st);
    if (Co % 128 == 0) return launch_gemm<128, 3>(Yprev, Whi, Wlo, p, st);
    return launch_gemm<64, 3>(Yprev, Whi, Wlo, p, st);
}

int fepe_mlp32_last(const float* Y, const float* ss, float slope, const float* W, const float* bias, float* logits,
                    float* weights, int B, int N, int Npad, int Ci, int Co, void* stream) {
    if (!Y || !ss || !W || !logits || B <= 0 || N <= 0 || Npad < N || Ci != 256 || (Co != 1 && Co != 4) ||
        (Co != 1 && weights != nullptr) || (reinterpret_cast<uintptr_t>(Y) & 15u) || !(slope >= 0.f && slope <= 1.f))
        return FEPE_E_BADARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* s2 = reinterpret_cast<const float2*>(ss);
    int S = 1;                                           // CTAs (= cluster size) per pair: ~2 CTAs per SM, at most 8
    while (S < 8 && B * S < 2 * 148) S *= 2;
    const int chunk = (((N + S - 1) / S) + 1) & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S, B);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = (Co == 1) ? chunk * sizeof(float) : 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    float* no_weights = nullptr;
    cudaError_t e;
    if (Co == 1) {
        if (cfg.dynamicSmemBytes > 48 * 1024) {
            e = cudaFuncSetAttribute(fepe::m32::fepe_mlp32_last_kernel<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(cfg.dynamicSmemBytes));
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        e = cudaLaunchKernelEx(&cfg, fepe::m32::fepe_mlp32_last_kernel<1, 256>, Y, s2, slope, W, bias, logits, weights, N, Npad);
    } else {
        e = cudaLaunchKernelEx(&cfg, fepe::m32::fepe_mlp32_last_kernel<4, 256>, Y, s2, slope, W, bias, logits, no_weights, N, Npad);
    }
    return static_cast<int>(e);
}

}  // extern "C"
