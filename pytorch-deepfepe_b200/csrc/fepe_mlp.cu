// Tensor-core path of the per-correspondence weight MLP (ErrorEstimator,
// deepFEPE/models/ErrorEstimators.py:46-64): Conv1d(k=1) == a GEMM over the channel dimension,
// followed by InstanceNorm1d(affine) (statistics over the N correspondences of ONE pair) and
// LeakyReLU(0.01).
//
//   fepe_mlp_gemm_persist_kernel   Y[M,Co] = X[M,Ci] . W[Co,Ci]^T + b   bf16 in, fp32 accumulate in TMEM
//                          (tcgen05.mma, operands staged by TMA with 128-byte swizzle), bf16 out, plus per-(pair,
//                          channel) sum / sum of squares for the InstanceNorm that follows.  Persistent CTAs, accumulator
//                          double-buffered in TMEM, 128 x 256 tiles; optionally with the PREVIOUS block's InstanceNorm +
//                          LeakyReLU applied to the operand tiles in shared memory (fepe_mlp_gemm_norm).
//   fepe_mlp_gemm_kernel   the same GEMM, one tile per CTA (Co % 128 != 0, and the A/B reference of the persistent one).
//   fepe_mlp_norm_kernel   X'[M,Co] = LeakyReLU(gamma (Y - mean) rstd + beta), bf16 (memory bound).
//   fepe_mlp_first_kernel  layer 1 (Ci = 4..8: too thin for a GEMM tile) on CUDA cores.
//   fepe_mlp_last_kernel   layer 6 (Co = 1) + softmax over the N correspondences of a pair.
//
// Rows: M = B * Npad, Npad = N rounded up to 128 so that a 128-row tile never straddles two pairs;
// padded rows are excluded from the statistics and written as zeros.
//
// Warp roles in the one-tile GEMM CTA (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM lane quarter = warp % 4); the persistent kernel's roles
// are described in front of it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/fepe_b200.h"
#include "fepe_common.cuh"
#include "fepe_dispatch.cuh"
#include "fepe_umma.cuh"

namespace fepe {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;          // 64 bf16 = 128 B = one swizzle atom
constexpr int kGemmThreads = 256;

struct GemmParams {
    int M, K, Co;          // M = B * Npad
    int Npad, Nvalid;      // rows per pair (padded / real)
    const float* bias;     // [Co], or null (a bias in front of an InstanceNorm cancels in the normalisation)
    __nv_bfloat16* Y;      // [M, Co]
    float* stats;          // [B, Co, 2] (sum, sum of squares), zeroed by the caller
    const float* ss;       // fused-norm variant only: [B, K/2, 4] = (a_2q, a_2q+1, d_2q, d_2q+1), x' = LeakyReLU(a y + d)
    float slope;
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
    constexpr int kChunks = BN / 8;          // 16-byte chunks per tile row
    __shared__ __align__(16) float sbias[BN];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kGemmBM;     // n-tiles vary fastest: the CTAs sharing an A tile are co-scheduled
    const int n0 = blockIdx.x * BN;
    const int num_kb = p.K / kGemmBK;
    const bool has_bias = p.bias != nullptr;
    if (has_bias && threadIdx.x >= 128 && threadIdx.x < 128 + BN) sbias[threadIdx.x - 128] = __ldg(p.bias + n0 + threadIdx.x - 128);

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
        // The tile is kept as 16-byte chunks of 8 columns, chunk k of row r at slot (k + r) mod kChunks: a quarter warp
        // (8 consecutive rows, or 8 consecutive chunks of one row) always touches 8 different slots = all 32 banks once,
        // for the 16-byte stores here, the 4-byte statistics loads and the 16-byte loads of the final copy alike.
        const int q = warp & 3;                               // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                        // row inside the tile
        const int pair_row = (m0 % p.Npad) + row;             // row inside its pair
        const bool valid = pair_row < p.Nvalid;
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float y0 = __uint_as_float(v[g * 8 + 2 * j]), y1 = __uint_as_float(v[g * 8 + 2 * j + 1]);
                    if (has_bias) {
                        const float2 bb = *reinterpret_cast<const float2*>(sbias + c + g * 8 + 2 * j);
                        y0 += bb.x; y1 += bb.y;
                    }
                    const __nv_bfloat162 h = __floats2bfloat162_rn(valid ? y0 : 0.f, valid ? y1 : 0.f);
                    pk[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                const int chunk = (c >> 3) + g;
                *reinterpret_cast<uint4*>(tile_y + (row * kChunks + ((chunk + row) & (kChunks - 1))) * 16) =
                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();

    // ---- all 256 threads: column statistics from the bf16 tile, then coalesced store ----
    {
        const int pair = m0 / p.Npad;
        if (p.stats != nullptr) {        // (the data-gradient GEMM of the backward pass needs no statistics)
            // a warp owns one chunk of 8 columns at a time: lane = (row residue rs = lane / 4, column pair cq = lane % 4)
            // sums rows rs, rs + 8, ... of its two columns; the 8 residues are folded by 3 shuffle levels and the 4
            // lanes with rs = 0 issue the 16 atomics of the chunk (BN x 2 atomics per tile)
            const int rs = lane >> 2, cq = lane & 3;
            for (int chunk = warp; chunk < kChunks; chunk += kGemmThreads / 32) {
                float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll 4
                for (int r = rs; r < kGemmBM; r += 8) {
                    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(
                        tile_y + (r * kChunks + ((chunk + r) & (kChunks - 1))) * 16 + cq * 4);
                    const float2 y = __bfloat1622float2(h);
                    s1a += y.x; s2a = fmaf(y.x, y.x, s2a);
                    s1b += y.y; s2b = fmaf(y.y, y.y, s2b);
                }
#pragma unroll
                for (int off = 4; off < 32; off <<= 1) {
                    s1a += __shfl_xor_sync(0xffffffffu, s1a, off);
                    s2a += __shfl_xor_sync(0xffffffffu, s2a, off);
                    s1b += __shfl_xor_sync(0xffffffffu, s1b, off);
                    s2b += __shfl_xor_sync(0xffffffffu, s2b, off);
                }
                if (rs == 0) {
                    float* st = p.stats + (static_cast<size_t>(pair) * p.Co + n0 + chunk * 8 + cq * 2) * 2;
                    atomicAdd(st, s1a); atomicAdd(st + 1, s2a);
                    atomicAdd(st + 2, s1b); atomicAdd(st + 3, s2b);
                }
            }
        }
        // store: each row of the tile is BN*2 bytes contiguous in Y, moved as 16-byte vectors
        for (int idx = threadIdx.x; idx < kGemmBM * kChunks; idx += kGemmThreads) {
            const int r = idx / kChunks, k = idx % kChunks;
            const uint4 v = *reinterpret_cast<const uint4*>(tile_y + (r * kChunks + ((k + r) & (kChunks - 1))) * 16);
            *reinterpret_cast<uint4*>(p.Y + static_cast<size_t>(m0 + r) * p.Co + n0 + k * 8) = v;
        }
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(static_cast<uint32_t>(BN < 32 ? 32 : BN)));
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant of the same GEMM: one CTA per SM walks over the output tiles (n fastest), the
// accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns) so that the epilogue of tile j overlaps the
// main loop of tile j+1, and the operand ring keeps running across tile boundaries.  Why: the
// one-tile-per-CTA kernel above pays TMEM allocation, barrier set-up and a serial load -> MMA ->
// epilogue chain per tile, which is most of the time of the thin layers (K = 128: two k-blocks), and
// with 128 x 128 tiles the fat layer (K = 1024) saturates the L2 (64 flop per operand byte against the
// ~12 TB/s the L2 slices deliver); BN = 256 raises that to 85 flop/B.
//
// 12 warps: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator,
// warps 4-11 = two epilogue groups of four warps (TMEM lane quarter = warp % 4); group g drains the
// columns [g BN/2, (g+1) BN/2) of the accumulator in chunks of 64 columns through its own 16 KB tile
// (same swizzled 16-byte-chunk layout, statistics and 16-byte stores as above).
// ------------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 384;
constexpr int kPersistFuseThreads = 512;             // + warps 12-15: operand transform of the fused-norm variant
constexpr int kEpiCols = 64;                         // columns per epilogue pass of a group
constexpr int kEpiTileBytes = kGemmBM * kEpiCols * 2;

__device__ __forceinline__ void epi_group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(128) : "memory");
}

// A 128 x 64 bf16 tile in shared memory as 16-byte chunks of 8 columns, chunk k of row r at slot (k + r) mod 8: a quarter
// warp (8 consecutive rows of one chunk, or the 8 chunks of one row) always touches 8 different slots = all 32 banks once.
__device__ __forceinline__ unsigned char* tile64_chunk(unsigned char* tile, int r, int k) {
    return tile + (r * (kEpiCols / 8) + ((k + r) & (kEpiCols / 8 - 1))) * 16;
}
// Column statistics of such a tile, called by the four warps (wq = 0..3) of a 128-thread group.  Warp wq owns the chunks
// 2 wq and 2 wq + 1; lane = (chunk half, row residue rg = lane % 16) reads rows rg, rg + 16, ... of its chunk as 16-byte
// vectors into 16 independent accumulators (packed FADD2 / FFMA2: a[4k] / a[4k+2] = sums, a[4k+1] / a[4k+3] = sums of
// squares of columns 2k / 2k+1).  A reduce-scatter over the 16 row residues (15 exchanges) leaves entry rg in lane rg,
// i.e. the chunk's 16 consecutive floats of `stats` ([column][sum, sum of squares]): ONE fully populated atomic
// instruction per warp.  `stats64` points at the statistics of the tile's first column.
__device__ __forceinline__ void tile64_stats(unsigned char* tile, int wq, int lane, float* stats64) {
    const int kc = 2 * wq + (lane >> 4);
    const int rg = lane & 15;
    float2 s1p[4], s2p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { s1p[k] = make_float2(0.f, 0.f); s2p[k] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int i = 0; i < kGemmBM / 16; ++i) {
        const uint4 u = *reinterpret_cast<const uint4*>(tile64_chunk(tile, rg + 16 * i, kc));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 y = make_float2(__uint_as_float(w[k] << 16), __uint_as_float(w[k] & 0xffff0000u));
            s1p[k] = __fadd2_rn(s1p[k], y);
            s2p[k] = __ffma2_rn(y, y, s2p[k]);
        }
    }
    float a[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[4 * k + 0] = s1p[k].x; a[4 * k + 1] = s2p[k].x; a[4 * k + 2] = s1p[k].y; a[4 * k + 3] = s2p[k].y;
    }
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? a[i] : a[i + h];
            const float keep = up ? a[i + h] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    atomicAdd(stats64 + kc * 16 + rg, a[0]);
}
// Coalesced copy of the tile to Y (row pitch ld elements): thread et of the group moves 8 chunks, 8 consecutive threads
// one 128-byte row segment.  `y64` points at the tile's first element.
__device__ __forceinline__ void tile64_store(unsigned char* tile, int et, __nv_bfloat16* y64, int ld) {
#pragma unroll
    for (int idx = et; idx < kGemmBM * (kEpiCols / 8); idx += 128) {
        const int r = idx / (kEpiCols / 8), k = idx % (kEpiCols / 8);
        *reinterpret_cast<uint4*>(y64 + static_cast<size_t>(r) * ld + k * 8) = *reinterpret_cast<const uint4*>(tile64_chunk(tile, r, k));
    }
}

// FUSE: 0 = plain GEMM (8 epilogue warps); 1 = fused norm with 8 epilogue + 4 transform warps, (a, d) read from global
// memory; 2 = fused norm with 4 epilogue warps (one group, four passes per 256-column tile: the fused layers are bound by
// the operand path, not by the epilogue) + 8 transform warps, (a, d) of the k-block delivered by the TMA producer into a
// 512-byte slot next to the stage (no global load, hence nothing for fence.proxy.async to wait for, in the transform).
template <int BN, int STAGES, int FUSE>
__global__ void __launch_bounds__(FUSE ? kPersistFuseThreads : kPersistThreads, 1)
fepe_mlp_gemm_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                             const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kABytes = kGemmBM * kGemmBK * 2;       // 16 KB
    constexpr int kBBytes = BN * kGemmBK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr uint32_t kTmemCols = 2 * BN;               // 256 or 512: a power of two
    constexpr int kNG = (FUSE == 2) ? 1 : 2;             // epilogue groups of four warps
    constexpr int kTW = (FUSE == 2) ? 8 : 4;             // transform warps (FUSE only): the last kTW warps of the CTA
    constexpr int kSsBytes = kGemmBK * 8;                // (a, d) fp32 of the 64 channels of a k-block
    unsigned char* epi_tiles = smem + STAGES * kStageBytes;                  // [2 groups][128][64] bf16
    unsigned char* ss_ring = epi_tiles + kEpiTileBytes;                      // FUSE == 2: [STAGES][512 B] in group 1's tile
    uint64_t* full = reinterpret_cast<uint64_t*>(epi_tiles + 2 * kEpiTileBytes);
    uint64_t* empty = full + STAGES;
    uint64_t* ready = empty + STAGES;                    // FUSE: the A tile of the stage has been transformed
    uint64_t* tmem_full = ready + STAGES;                // [2]
    uint64_t* tmem_empty = tmem_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = p.K / kGemmBK;
    const int n_tiles = p.Co / BN;
    const int tiles = (p.M / kGemmBM) * n_tiles;
    const int n_local = (tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], kTW); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4 * kNG); }
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
                const int m0 = (t / n_tiles) * kGemmBM, n0 = (t % n_tiles) * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const uint32_t s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* sa = smem + s * kStageBytes;
                    mbar_arrive_expect_tx(&full[s], kStageBytes + (FUSE == 2 ? kSsBytes : 0));
                    tma_load_2d(sa, &map_a, kb * kGemmBK, m0, &full[s]);
                    tma_load_2d(sa + kABytes, &map_w, kb * kGemmBK, n0, &full[s]);
                    if constexpr (FUSE == 2) {
                        const int pair = m0 / p.Npad;
                        bulk_g2s(ss_ring + s * kSsBytes, p.ss + (static_cast<size_t>(pair) * p.K + kb * kGemmBK) * 2, kSsBytes,
                                 &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                   (static_cast<uint32_t>(kGemmBM >> 4) << 24);
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
                mbar_wait(FUSE ? &ready[s] : &full[s], ph);
                tcgen05_fence_after();
                if (lane == 0) {
                    const unsigned char* sa = smem + s * kStageBytes;
                    const uint64_t da = umma_desc_k_sw128(sa);
                    const uint64_t db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 16; ++k) {
                        umma_bf16(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit(&empty[s]);
                    if (kb == num_kb - 1) tcgen05_commit(&tmem_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (FUSE && warp >= 16 - kTW) {
        // ---------------- operand transform (fused InstanceNorm + LeakyReLU of the previous layer) ----------------
        // The A tile holds the previous layer's PRE-norm output; x' = LeakyReLU(a y + d) with the per-(pair, channel)
        // (a, d) of fepe_mlp_scale_shift is applied in place before the MMA reads the stage.  Thread tt owns the
        // physical 16-byte chunk `pos` of rows r0, r0 + 16, ...; under the 128-byte swizzle that is the logical chunk
        // pos ^ (r0 & 7) for all of them (16 i leaves r & 7 unchanged), so its eight channels -- and their (a, d),
        // requested before the stage's data are waited for -- are the same for every row of a k-block.
        constexpr int kRowStep = kTW * 4;                       // rows between two chunks of a thread (16 or 32)
        const int tt = static_cast<int>(threadIdx.x) - (kPersistFuseThreads - kTW * 32);
        const int pos = tt & 7;
        const int r0 = tt >> 3;
        const int cl = pos ^ (r0 & 7);
        const float2 slope2 = make_float2(p.slope, p.slope);
        uint32_t it = 0;
        // (a, a, d, d) of channel pairs 0..3 of the chunk, requested one k-block ahead and before the wait for the
        // stage: this role is the slowest stage of the pipeline, so its data are usually waiting and a load issued next
        // to its use would be fully exposed (under the epilogue's store traffic an L1/L2 hit takes > 1000 cycles here).
        // Measured alternatives (profiles/r1_mlp_fused_norm.md): no prefetch 860 us, this 772 us, two k-blocks ahead
        // issued after the hand-over 950 us on the K = 1024 layer.
        auto ss_ptr = [&](int j, int kb) {
            const int t = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
            const int pair = ((t / n_tiles) * kGemmBM) / p.Npad;
            return reinterpret_cast<const float4*>(p.ss) + (static_cast<size_t>(pair) * p.K + kb * kGemmBK + cl * 8) / 2;
        };
        float4 cn[4];
        if (FUSE == 1 && n_local > 0) {
            const float4* sp0 = ss_ptr(0, 0);
#pragma unroll
            for (int q = 0; q < 4; ++q) cn[q] = __ldg(sp0 + q);
        }
        for (int j = 0; j < n_local; ++j) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1u;
                float4 c[4];
                if constexpr (FUSE == 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) c[q] = cn[q];
                    const bool last_kb = kb + 1 == num_kb;
                    if (!last_kb || j + 1 < n_local) {
                        const float4* spn = last_kb ? ss_ptr(j + 1, 0) : ss_ptr(j, kb + 1);
#pragma unroll
                        for (int q = 0; q < 4; ++q) cn[q] = __ldg(spn + q);
                    }
                }
                mbar_wait(&full[s], ph);
                if constexpr (FUSE == 2) {
                    const float4* sl = reinterpret_cast<const float4*>(ss_ring + s * kSsBytes) + cl * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) c[q] = sl[q];
                }
                unsigned char* sa = smem + s * kStageBytes + r0 * 128 + pos * 16;
#pragma unroll
                for (int i = 0; i < kGemmBM / kRowStep; ++i) {
                    uint4* ptr = reinterpret_cast<uint4*>(sa + i * (kRowStep * 128));
                    const uint4 u = *ptr;
                    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
                    uint32_t o[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 y = make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u));
                        const float2 tv = __ffma2_rn(y, make_float2(c[q].x, c[q].y), make_float2(c[q].z, c[q].w));
                        const float2 sv = __fmul2_rn(tv, slope2);
                        const __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(tv.x, sv.x), fmaxf(tv.y, sv.y));
                        o[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *ptr = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async();                            // generic-proxy writes -> visible to the MMA's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[s]);
            }
        }
    } else if (warp >= 4 && warp < 4 + 4 * kNG) {
        // ---------------- epilogue groups ----------------
        const int ew = warp - 4;
        const int g = ew >> 2;                                // column half of the accumulator
        const int wq = ew & 3;                                // warp inside the group
        const int q = warp & 3;                               // TMEM lane quarter this warp may read (== wq)
        const int row = q * 32 + lane;
        const int et = wq * 32 + lane;                        // thread inside the group
        unsigned char* tile_y = epi_tiles + g * kEpiTileBytes;
        const bool has_bias = p.bias != nullptr;
        for (int j = 0; j < n_local; ++j) {
            const int t = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
            const int m0 = (t / n_tiles) * kGemmBM, n0 = (t % n_tiles) * BN;
            const uint32_t b = static_cast<uint32_t>(j) & 1u;
            const uint32_t use = static_cast<uint32_t>(j) >> 1;
            const int pair = m0 / p.Npad;
            const bool valid = (m0 - pair * p.Npad) + row < p.Nvalid;
            mbar_wait(&tmem_full[b], use & 1u);
            tcgen05_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < BN / kNG; cc += kEpiCols) {
                const int col0 = g * (BN / kNG) + cc;         // first accumulator column of this pass
#pragma unroll
                for (int c = 0; c < kEpiCols; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * static_cast<uint32_t>(BN) +
                                  static_cast<uint32_t>(col0 + c), v);
                    if (has_bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + col0 + c);   // 16-byte aligned
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 bb = __ldg(bp + i);
                            v[4 * i + 0] = __float_as_uint(__uint_as_float(v[4 * i + 0]) + bb.x);
                            v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + bb.y);
                            v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + bb.z);
                            v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + bb.w);
                        }
                    }
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = 0u;
                    }
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg) {
                        uint32_t pk[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[gg * 8 + 2 * jj]),
                                                                           __uint_as_float(v[gg * 8 + 2 * jj + 1]));
                            pk[jj] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        const int chunk = (c >> 3) + gg;
                        *reinterpret_cast<uint4*>(tile64_chunk(tile_y, row, chunk)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if (cc + kEpiCols >= BN / kNG) {              // last read of this accumulator buffer: hand it back
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[b]);
                }
                epi_group_sync(g);
                if (p.stats != nullptr)
                    tile64_stats(tile_y, wq, lane, p.stats + (static_cast<size_t>(pair) * p.Co + n0 + col0) * 2);
                tile64_store(tile_y, et, p.Y + static_cast<size_t>(m0) * p.Co + n0 + col0, p.Co);
                epi_group_sync(g);                            // the tile is rewritten by the next pass
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

// (a, d) of the fused-norm GEMM: x' = LeakyReLU(a y + d), a = gamma rstd, d = beta - mean a (the same fp32 expressions
// as fepe_mlp_norm_kernel, so both paths produce the same bf16 activations).  One thread per channel pair; layout
// ss[b][c/2] = (a_c, a_c+1, d_c, d_c+1).  With `clear` the statistics are zeroed for the next layer's accumulation.
__global__ void __launch_bounds__(256) fepe_mlp_scale_shift_kernel(float* __restrict__ stats, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float4* __restrict__ ss,
                                                                   int total_pairs, int Co, int Nvalid, float eps, int clear) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // channel pair of (b, c)
    if (i >= total_pairs) return;
    const int c = (i % (Co / 2)) * 2;
    float4* sp = reinterpret_cast<float4*>(stats) + i;         // (s1_c, s2_c, s1_c+1, s2_c+1)
    const float4 v = *sp;
    const float invN = 1.0f / static_cast<float>(Nvalid);
    const float m0 = v.x * invN, m1 = v.z * invN;
    const float var0 = fmaxf(v.y * invN - m0 * m0, 0.f), var1 = fmaxf(v.w * invN - m1 * m1, 0.f);
    const float a0 = rsqrtf(var0 + eps) * gamma[c], a1 = rsqrtf(var1 + eps) * gamma[c + 1];
    ss[i] = make_float4(a0, a1, beta[c] - m0 * a0, beta[c + 1] - m1 * a1);
    if (clear) *sp = make_float4(0.f, 0.f, 0.f, 0.f);
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
    // A thread keeps ONE group of 8 channels (its scale / shift live in registers) and walks down the rows: per 16-byte
    // vector that is one load, 8 FMA, the LeakyReLU selects, 4 conversions and one store -- no shared-memory traffic in
    // the loop -- with four rows in flight.  A warp covers 512 contiguous bytes of a row (or whole rows when Co < 256).
    const int vec_per_row = Co / 8;
    const size_t base = (static_cast<size_t>(b) * Npad + r0) * Co;
    const int nthr = static_cast<int>(blockDim.x);
    const int rstep = (nthr >= vec_per_row) ? nthr / vec_per_row : 1;          // rows a pass of the CTA covers
    const int rfirst = static_cast<int>(threadIdx.x) / vec_per_row;
    const int rows_valid = (Nvalid - r0 < 128) ? ((Nvalid - r0 > 0) ? Nvalid - r0 : 0) : 128;
    for (int cg = static_cast<int>(threadIdx.x) % vec_per_row; cg < vec_per_row && rfirst < rstep; cg += nthr) {
        const int c = cg * 8;
        float a[8], d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] = sc[c + k]; d[k] = sh[c + k]; }
        for (int r = rfirst; r < 128; r += 4 * rstep) {
            uint4 in[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + u * rstep;
                in[u] = make_uint4(0u, 0u, 0u, 0u);
                if (rr < rows_valid) in[u] = __ldcs(reinterpret_cast<const uint4*>(Y + base + static_cast<size_t>(rr) * Co + c));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + u * rstep;
                if (rr >= 128) continue;
                uint4 out = make_uint4(0u, 0u, 0u, 0u);                       // padded rows stay zero
                if (rr < rows_valid) {
                    const uint32_t w[4] = {in[u].x, in[u].y, in[u].z, in[u].w};
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                        const float t0 = fmaf(y.x, a[2 * k], d[2 * k]), t1 = fmaf(y.y, a[2 * k + 1], d[2 * k + 1]);
                        const __nv_bfloat162 v = __floats2bfloat162_rn(t0 > 0.f ? t0 : slope * t0, t1 > 0.f ? t1 : slope * t1);
                        o[k] = *reinterpret_cast<const uint32_t*>(&v);
                    }
                    out = make_uint4(o[0], o[1], o[2], o[3]);
                }
                *reinterpret_cast<uint4*>(X + base + static_cast<size_t>(rr) * Co + c) = out;
            }
        }
    }
}

// layer 1: X0 [B,N,Ci] fp32 (Ci <= 8) -> Y [B*Npad, Co=64] bf16 + stats.  One CTA of 128 threads per (pair, 128-row
// slab): thread = row computes its 64 outputs four channels at a time (weights transposed in shared memory and read as
// broadcast 16-byte vectors, packed FFMA2), rounds them to bf16 into the same rotated-chunk tile as the GEMM epilogue, and
// the tile's column statistics and coalesced 16-byte stores are the GEMM epilogue's own (tile64_stats / tile64_store).
// CI > 0: compile-time channel count (4 and 7 are the reference's configurations); CI = 0: run-time Ci with guards.
template <int CI>
__global__ void __launch_bounds__(128) fepe_mlp_first_kernel(const float* __restrict__ X0, const float* __restrict__ W,
                                                             const float* __restrict__ bias,
                                                             __nv_bfloat16* __restrict__ Y, float* __restrict__ stats,
                                                             int B, int N, int Npad, int Ci_rt, int Co) {
    __shared__ __align__(16) unsigned char tile[kEpiTileBytes];   // [128][64] bf16, rotated 16-byte chunks
    __shared__ __align__(16) float w_t[8 * 64];                   // [k][c]: W^T, zero beyond Ci
    __shared__ __align__(16) float b_s[64];
    const int Ci = (CI > 0) ? CI : Ci_rt;
    const int b = blockIdx.y;
    const int row = threadIdx.x;
    const int r = blockIdx.x * 128 + row;
    for (int i = threadIdx.x; i < 8 * 64; i += 128) {
        const int k = i >> 6, c = i & 63;
        w_t[i] = (k < Ci) ? W[c * Ci + k] : 0.f;
    }
    if (threadIdx.x < 64) b_s[threadIdx.x] = bias[threadIdx.x];
    float x[8];
    const bool valid = r < N;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (valid && k < Ci) ? X0[(static_cast<size_t>(b) * N + r) * Ci + k] : 0.f;
    __syncthreads();
    constexpr int KMAX = (CI > 0) ? CI : 8;
#pragma unroll
    for (int chunk = 0; chunk < 8; ++chunk) {                     // 8 channels = one 16-byte chunk of the tile
        uint32_t pk[4];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c = chunk * 8 + half * 4;
            const float4 bb = *reinterpret_cast<const float4*>(b_s + c);
            float2 y01 = make_float2(bb.x, bb.y), y23 = make_float2(bb.z, bb.w);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {                      // same order of accumulation as the scalar loop: b, k = 0, 1, ...
                const float4 w4 = *reinterpret_cast<const float4*>(w_t + k * 64 + c);
                const float2 xx = make_float2(x[k], x[k]);
                y01 = __ffma2_rn(make_float2(w4.x, w4.y), xx, y01);
                y23 = __ffma2_rn(make_float2(w4.z, w4.w), xx, y23);
            }
            if (!valid) { y01 = make_float2(0.f, 0.f); y23 = make_float2(0.f, 0.f); }
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(y01.x, y01.y), h1 = __floats2bfloat162_rn(y23.x, y23.y);
            pk[half * 2 + 0] = *reinterpret_cast<const uint32_t*>(&h0);
            pk[half * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        }
        *reinterpret_cast<uint4*>(tile64_chunk(tile, row, chunk)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    __syncthreads();
    tile64_stats(tile, static_cast<int>(threadIdx.x) >> 5, static_cast<int>(threadIdx.x) & 31,
                 stats + static_cast<size_t>(b) * Co * 2);
    tile64_store(tile, static_cast<int>(threadIdx.x), Y + (static_cast<size_t>(b) * Npad + blockIdx.x * 128) * Co, Co);
}

// last layer (Ci -> 1) + softmax over the N rows of a pair.  One CTA per pair, 256 threads.
// FUSE: X is the PRE-norm output of layer 5 and x' = LeakyReLU(a y + d), rounded to bf16 like the stored activation of
// the unfused path, is formed on the fly from ss (fepe_mlp_scale_shift layout).
template <bool FUSE>
__global__ void fepe_mlp_last_kernel(const __nv_bfloat16* __restrict__ X, const float* __restrict__ W, float bias,
                                     float* __restrict__ logits, float* __restrict__ weights, int N, int Npad, int Ci,
                                     const float4* __restrict__ ss, float slope) {
    extern __shared__ float sh[];            // [Npad] logits
    __shared__ float red[8];
    const int b = blockIdx.x;
    {
        // a warp per row, 16-byte vectors (a 256-channel row is exactly one vector per lane), four rows in flight
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
        const __nv_bfloat16* xb = X + static_cast<size_t>(b) * Npad * Ci;
        for (int r = warp * 4; r < N; r += nwarp * 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = lane * 8; k < Ci; k += 256) {
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = __ldg(W + k + j);
                float4 cf[4];
                if constexpr (FUSE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) cf[j] = __ldg(ss + (static_cast<size_t>(b) * Ci + k) / 2 + j);
                }
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (r + u < N) v[u] = __ldcs(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r + u) * Ci + k));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t q[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q[j]));
                        if constexpr (FUSE) {
                            const float2 tv = __ffma2_rn(y, make_float2(cf[j].x, cf[j].y), make_float2(cf[j].z, cf[j].w));
                            const float2 sv = __fmul2_rn(tv, make_float2(slope, slope));
                            y = __bfloat1622float2(__floats2bfloat162_rn(fmaxf(tv.x, sv.x), fmaxf(tv.y, sv.y)));
                        }
                        acc[u] = fmaf(y.x, w[2 * j], acc[u]);
                        acc[u] = fmaf(y.y, w[2 * j + 1], acc[u]);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            }
            if (lane < 4 && r + lane < N) {
                const float v = ((lane == 0) ? acc[0] : (lane == 1) ? acc[1] : (lane == 2) ? acc[2] : acc[3]) + bias;
                sh[r + lane] = v;
                logits[static_cast<size_t>(b) * N + r + lane] = v;
            }
        }
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
// Backward of  X' = LeakyReLU(gamma * Yhat + beta),  Yhat = (Y - mean) rstd  (InstanceNorm over the N rows
// of a pair).  With dZ = dX' * (X' > 0 ? 1 : slope):
//   A1[b,c] = sum_n dZ,  A2[b,c] = sum_n dZ Yhat           (fepe_mlp_normbwd_reduce_kernel)
//   dY = rstd gamma (dZ - A1/N - Yhat A2/N)                 (fepe_mlp_normbwd_apply_kernel)
//   dgamma[c] = sum_b A2, dbeta[c] = sum_b A1 (host side, tiny);  the bias of the preceding conv gets an
//   exactly zero gradient (sum_n dY = 0): InstanceNorm cancels it.
// One CTA per (128-row slab, pair); a thread owns 8 fixed channels and walks rows, so sums stay in registers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4 v, float (&o)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        o[2 * k] = __low2float(t);
        o[2 * k + 1] = __high2float(t);
    }
}

__global__ void __launch_bounds__(256) fepe_mlp_normbwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ dX, const __nv_bfloat16* __restrict__ Xp, const __nv_bfloat16* __restrict__ Y,
    const float* __restrict__ stats, float* __restrict__ A, int Co, int Npad, int Nvalid, float eps, float slope) {
    extern __shared__ float sm[];            // [Co] mean, [Co] rstd, [2*Co] accumulators
    float* mean = sm;
    float* rstd = sm + Co;
    float* acc = sm + 2 * Co;
    const int b = blockIdx.y, r0 = blockIdx.x * 128;
    const float invN = 1.0f / static_cast<float>(Nvalid);
    for (int c = threadIdx.x; c < Co; c += 256) {
        const float s1 = stats[(static_cast<size_t>(b) * Co + c) * 2], s2 = stats[(static_cast<size_t>(b) * Co + c) * 2 + 1];
        const float mu = s1 * invN;
        mean[c] = mu;
        rstd[c] = rsqrtf(fmaxf(s2 * invN - mu * mu, 0.f) + eps);
        acc[2 * c] = 0.f;
        acc[2 * c + 1] = 0.f;
    }
    __syncthreads();
    const int vpr = Co / 8;                          // vectors per row (divides 256)
    const int v = threadIdx.x % vpr, rg = threadIdx.x / vpr, nrg = 256 / vpr;
    const int c0 = v * 8;
    float a1[8], a2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a1[k] = 0.f; a2[k] = 0.f; }
    const size_t base = (static_cast<size_t>(b) * Npad + r0) * Co + c0;
    for (int r = rg; r < 128 && r0 + r < Nvalid; r += nrg) {
        float g[8], xp[8], y[8];
        unpack8(*reinterpret_cast<const uint4*>(dX + base + static_cast<size_t>(r) * Co), g);
        unpack8(*reinterpret_cast<const uint4*>(Xp + base + static_cast<size_t>(r) * Co), xp);
        unpack8(*reinterpret_cast<const uint4*>(Y + base + static_cast<size_t>(r) * Co), y);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float dz = xp[k] > 0.f ? g[k] : slope * g[k];
            a1[k] += dz;
            a2[k] = fmaf(dz, (y[k] - mean[c0 + k]) * rstd[c0 + k], a2[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&acc[2 * (c0 + k)], a1[k]); atomicAdd(&acc[2 * (c0 + k) + 1], a2[k]); }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Co; i += 256) atomicAdd(A + static_cast<size_t>(b) * Co * 2 + i, acc[i]);
}

__global__ void __launch_bounds__(256) fepe_mlp_normbwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dX, const __nv_bfloat16* __restrict__ Xp, const __nv_bfloat16* __restrict__ Y,
    const float* __restrict__ stats, const float* __restrict__ A, const float* __restrict__ gamma,
    __nv_bfloat16* __restrict__ dY, int Co, int Npad, int Nvalid, float eps, float slope) {
    extern __shared__ float sm[];            // per channel: mean, rstd, k0 = rstd*gamma, k1 = A1/N, k2 = A2/N
    float* mean = sm; float* rstd = sm + Co; float* k0 = sm + 2 * Co; float* k1 = sm + 3 * Co; float* k2 = sm + 4 * Co;
    const int b = blockIdx.y, r0 = blockIdx.x * 128;
    const float invN = 1.0f / static_cast<float>(Nvalid);
    for (int c = threadIdx.x; c < Co; c += 256) {
        const float s1 = stats[(static_cast<size_t>(b) * Co + c) * 2], s2 = stats[(static_cast<size_t>(b) * Co + c) * 2 + 1];
        const float mu = s1 * invN;
        const float rs = rsqrtf(fmaxf(s2 * invN - mu * mu, 0.f) + eps);
        mean[c] = mu; rstd[c] = rs; k0[c] = rs * gamma[c];
        k1[c] = A[(static_cast<size_t>(b) * Co + c) * 2] * invN;
        k2[c] = A[(static_cast<size_t>(b) * Co + c) * 2 + 1] * invN;
    }
    __syncthreads();
    const int vpr = Co / 8;
    const size_t base = (static_cast<size_t>(b) * Npad + r0) * Co;
    for (int idx = threadIdx.x; idx < 128 * vpr; idx += 256) {
        const int r = idx / vpr, c0 = (idx % vpr) * 8;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < Nvalid) {
            float g[8], xp[8], y[8];
            const size_t off = base + static_cast<size_t>(r) * Co + c0;
            unpack8(*reinterpret_cast<const uint4*>(dX + off), g);
            unpack8(*reinterpret_cast<const uint4*>(Xp + off), xp);
            unpack8(*reinterpret_cast<const uint4*>(Y + off), y);
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float d[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = c0 + 2 * k + h;
                    const float dz = xp[2 * k + h] > 0.f ? g[2 * k + h] : slope * g[2 * k + h];
                    const float yh = (y[2 * k + h] - mean[c]) * rstd[c];
                    d[h] = k0[c] * (dz - k1[c] - yh * k2[c]);
                }
                const __nv_bfloat162 t = __floats2bfloat162_rn(d[0], d[1]);
                o[k] = *reinterpret_cast<const uint32_t*>(&t);
            }
            out = make_uint4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<uint4*>(dY + base + static_cast<size_t>(r) * Co + c0) = out;
    }
}

// Backward of the last layer  logit[m] = X[m,:] . w + b :  dX[m,k] = dl[m] w[k],  dw[k] += sum_m dl[m] X[m,k],
// db += sum_m dl[m].  One CTA per (128-row slab, pair), thread = channel (Ci = 256).
__global__ void __launch_bounds__(256) fepe_mlp_last_bwd_kernel(const float* __restrict__ dlogits,
                                                                const __nv_bfloat16* __restrict__ X,
                                                                const float* __restrict__ W,
                                                                __nv_bfloat16* __restrict__ dX, float* __restrict__ dW,
                                                                float* __restrict__ db, int N, int Npad, int Ci) {
    __shared__ float dl[128];
    const int b = blockIdx.y, r0 = blockIdx.x * 128;
    if (threadIdx.x < 128) dl[threadIdx.x] = (r0 + threadIdx.x < N) ? dlogits[static_cast<size_t>(b) * N + r0 + threadIdx.x] : 0.f;
    __syncthreads();
    for (int k = threadIdx.x; k < Ci; k += 256) {
        const float wk = W[k];
        float acc = 0.f;
        const size_t base = (static_cast<size_t>(b) * Npad + r0) * Ci + k;
        for (int r = 0; r < 128; ++r) {
            acc = fmaf(dl[r], __bfloat162float(X[base + static_cast<size_t>(r) * Ci]), acc);
            dX[base + static_cast<size_t>(r) * Ci] = __float2bfloat16(dl[r] * wk);
        }
        atomicAdd(dW + k, acc);
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int r = 0; r < 128; ++r) s += dl[r];
        atomicAdd(db, s);
    }
}

// Backward of the first layer (Ci <= 8, Co = 64): dX0[b,n,ci] = sum_c dY[m,c] W[c,ci] (fp32), dW[c,ci] += sum_m dY[m,c] X0[m,ci].
__global__ void __launch_bounds__(128) fepe_mlp_first_bwd_kernel(const __nv_bfloat16* __restrict__ dY,
                                                                 const float* __restrict__ X0,
                                                                 const float* __restrict__ W, float* __restrict__ dX0,
                                                                 float* __restrict__ dW, int N, int Npad, int Ci, int Co) {
    __shared__ float w_s[64 * 8];
    __shared__ float dy_s[128][64 + 1];
    __shared__ float x_s[128][8];
    const int b = blockIdx.y, r0 = blockIdx.x * 128, r = r0 + threadIdx.x;
    for (int i = threadIdx.x; i < Co * Ci; i += 128) w_s[i] = W[i];
    const bool valid = r < N;
#pragma unroll
    for (int k = 0; k < 8; ++k) x_s[threadIdx.x][k] = (valid && k < Ci) ? X0[(static_cast<size_t>(b) * N + r) * Ci + k] : 0.f;
    for (int c = 0; c < Co; ++c)
        dy_s[threadIdx.x][c] = valid ? __bfloat162float(dY[(static_cast<size_t>(b) * Npad + r) * Co + c]) : 0.f;
    __syncthreads();
    if (valid) {
        float g[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = 0.f;
        for (int c = 0; c < Co; ++c) {
            const float d = dy_s[threadIdx.x][c];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < Ci) g[k] = fmaf(d, w_s[c * Ci + k], g[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < Ci) dX0[(static_cast<size_t>(b) * N + r) * Ci + k] = g[k];
    }
    for (int o = threadIdx.x; o < Co * Ci; o += 128) {
        const int c = o / Ci, k = o % Ci;
        float acc = 0.f;
        for (int rr = 0; rr < 128; ++rr) acc = fmaf(dy_s[rr][c], x_s[rr][k], acc);
        atomicAdd(dW + o, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: dW[Co,Ci] += dY[M,Co]^T . X[M,Ci]   (split over K = M slabs, fp32 atomics).
// Both operands are "MN-major" for the tensor core: the GEMM-K index (the row m) is the strided one and
// the M / N index (channel) is contiguous -- exactly the row-major activation tensors, so no transposed
// copies exist.  A TMA box is 64 rows x 64 channels (128 B, SWIZZLE_128B); a 128-channel operand tile is
// two such boxes 8192 B apart (the descriptor's leading-byte-offset), 8-row groups are 1024 B apart
// (stride-byte-offset); one MMA consumes 16 rows = 2048 B.   (canonical layout: cute/atom/mma_traits_sm100.hpp,
// LayoutType::B128 MN-major  ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(const void* smem_ptr) {
    const uint64_t addr = static_cast<uint64_t>(smem_u32(smem_ptr));
    return ((addr >> 4) & 0x3FFFull) | (static_cast<uint64_t>(8192 >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}

struct WgradParams {
    int M, Co, Ci;
    int rows_per_slab;       // multiple of 64
    float* dW;               // [Co, Ci] fp32, accumulated with atomics (zeroed by the caller)
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 2)
fepe_mlp_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                      const WgradParams p) {
    constexpr int STAGES = 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kABytes = 64 * 128 * 2;                 // 64 rows x 128 channels (two 64x64 boxes)
    constexpr int kBBytes = 64 * BN * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
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
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = static_cast<uint32_t>(kb / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* sa = smem + s * kStageBytes;
                mbar_arrive_expect_tx(&full[s], kStageBytes);
                const int r = row0 + kb * 64;
                tma_load_2d(sa, &map_dy, co0, r, &full[s]);
                tma_load_2d(sa + 8192, &map_dy, co0 + 64, r, &full[s]);
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) tma_load_2d(sa + kABytes + j * 8192, &map_x, ci0 + j * 64, r, &full[s]);
            }
        }
    } else if (warp == 1) {
        // D = f32, A = B = bf16, BOTH MN-major (bits 15, 16), N>>3 at bit 17, M>>4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = static_cast<uint32_t>(kb / STAGES) & 1u;
            mbar_wait(&full[s], ph);
            tcgen05_fence_after();
            if (lane == 0) {
                const unsigned char* sa = smem + s * kStageBytes;
                const uint64_t da = umma_desc_mn_sw128(sa);
                const uint64_t db = umma_desc_mn_sw128(sa + kABytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)      // 16 rows = 2048 B per MMA
                    umma_bf16(tmem_base, da + static_cast<uint64_t>(k * 128), db + static_cast<uint64_t>(k * 128), idesc,
                              (kb | k) != 0 ? 1u : 0u);
                tcgen05_commit(&empty[s]);
                if (kb == num_kb - 1) tcgen05_commit(tmem_full);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && num_kb > 0) {
        const int q = warp & 3;
        const int row = q * 32 + lane;                         // output row = channel co0 + row
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        float* out = p.dW + static_cast<size_t>(co0 + row) * p.Ci + ci0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(out + c + j, __uint_as_float(v[j]));
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(static_cast<uint32_t>(BN)));
    }
}

// ------------------------------------------------------------------------------------------------
// 2-D bf16 row-major [rows, K] tensor, box = 64 (contiguous dim) x box_rows, 128-byte swizzle
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
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e =
            cudaFuncSetAttribute(fepe_mlp_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured[dev & 63] = true;
    }
    dim3 grid(p.Co / BN, p.M / kGemmBM);
    fepe_mlp_gemm_kernel<BN, STAGES><<<grid, kGemmThreads, smem, stream>>>(ma, mw, p);
    return static_cast<int>(cudaGetLastError());
}

template <int BN, int STAGES, int FUSE>
static int launch_gemm_persist(const void* X, const void* W, const GemmParams& p, cudaStream_t stream) {
    CUtensorMap ma, mw;
    if (!make_map(&ma, X, p.M, p.K, kGemmBM) || !make_map(&mw, W, p.Co, p.K, BN)) return FEPE_E_NODEVICE;
    constexpr int smem = STAGES * (kGemmBM * kGemmBK * 2 + BN * kGemmBK * 2) + 2 * kEpiTileBytes + 256 + 1024;
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sms[dev & 63] == 0) {
        cudaError_t e = cudaFuncSetAttribute(fepe_mlp_gemm_persist_kernel<BN, STAGES, FUSE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || n <= 0) return FEPE_E_NODEVICE;
        sms[dev & 63] = n;
    }
    const int tiles = (p.M / kGemmBM) * (p.Co / BN);
    const int grid = tiles < sms[dev & 63] ? tiles : sms[dev & 63];
    fepe_mlp_gemm_persist_kernel<BN, STAGES, FUSE><<<grid, FUSE ? kPersistFuseThreads : kPersistThreads, smem, stream>>>(ma, mw, p);
    return static_cast<int>(cudaGetLastError());
}

template <int BN>
static int launch_wgrad(const void* dY, const void* X, const WgradParams& p, int slabs, cudaStream_t stream) {
    CUtensorMap my, mx;
    if (!make_map(&my, dY, p.M, p.Co, 64) || !make_map(&mx, X, p.M, p.Ci, 64)) return FEPE_E_NODEVICE;
    constexpr int smem = 3 * (64 * 128 * 2 + 64 * BN * 2) + 256 + 1024;
    static bool configured[64] = {false};                 // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fepe_mlp_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
        configured[dev & 63] = true;
    }
    dim3 grid(p.Ci / BN, p.Co / 128, slabs);
    fepe_mlp_wgrad_kernel<BN><<<grid, kGemmThreads, smem, stream>>>(my, mx, p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace fepe

extern "C" {

int fepe_mlp_normbwd(const void* dX, const void* Xp, const void* Y, const float* stats, const float* gamma, float* A,
                     void* dY, int B, int Npad, int Nvalid, int Co, float eps, float slope, void* stream) {
    if (!dX || !Xp || !Y || !stats || !gamma || !A || !dY || B <= 0 || (Co % 64) != 0 || (Npad % 128) != 0)
        return FEPE_E_BADARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(Npad / 128, B);
    fepe::fepe_mlp_normbwd_reduce_kernel<<<grid, 256, 4 * Co * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(dX), static_cast<const __nv_bfloat16*>(Xp), static_cast<const __nv_bfloat16*>(Y),
        stats, A, Co, Npad, Nvalid, eps, slope);
    fepe::fepe_mlp_normbwd_apply_kernel<<<grid, 256, 5 * Co * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(dX), static_cast<const __nv_bfloat16*>(Xp), static_cast<const __nv_bfloat16*>(Y),
        stats, A, gamma, static_cast<__nv_bfloat16*>(dY), Co, Npad, Nvalid, eps, slope);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp_last_bwd(const float* dlogits, const void* X, const float* W, void* dX, float* dW, float* db, int B, int N,
                      int Npad, int Ci, void* stream) {
    if (!dlogits || !X || !W || !dX || !dW || !db || B <= 0 || (Npad % 128) != 0) return FEPE_E_BADARG;
    dim3 grid(Npad / 128, B);
    fepe::fepe_mlp_last_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dlogits, static_cast<const __nv_bfloat16*>(X), W, static_cast<__nv_bfloat16*>(dX), dW, db, N, Npad, Ci);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp_first_bwd(const void* dY, const float* X0, const float* W, float* dX0, float* dW, int B, int N, int Npad,
                       int Ci, int Co, void* stream) {
    if (!dY || !X0 || !W || !dX0 || !dW || B <= 0 || Ci <= 0 || Ci > 8 || Co != 64 || (Npad % 128) != 0) return FEPE_E_BADARG;
    dim3 grid(Npad / 128, B);
    fepe::fepe_mlp_first_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dY), X0, W, dX0, dW, N, Npad, Ci, Co);
    return static_cast<int>(cudaGetLastError());
}

// dW[Co,Ci] += dY[M,Co]^T X[M,Ci]  (bf16 operands, fp32 accumulation; dW zeroed by the caller)
int fepe_mlp_wgrad(const void* dY, const void* X, float* dW, int M, int Co, int Ci, void* stream) {
    if (!dY || !X || !dW || M <= 0 || (M % 64) != 0 || (Co % 128) != 0 || (Ci % 64) != 0) return FEPE_E_BADARG;
    const int bn = (Ci % 128 == 0) ? 128 : 64;
    const int tiles = (Co / 128) * (Ci / bn);
    int slabs = (592 + tiles - 1) / tiles;                    // ~4 waves of 148 SMs
    const int max_slabs = M / 64;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    int rows_per_slab = ((M + slabs - 1) / slabs + 63) / 64 * 64;
    slabs = (M + rows_per_slab - 1) / rows_per_slab;
    fepe::WgradParams p{M, Co, Ci, rows_per_slab, dW};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return bn == 128 ? fepe::launch_wgrad<128>(dY, X, p, slabs, st) : fepe::launch_wgrad<64>(dY, X, p, slabs, st);
}

// Y = X W^T + b (bf16 in / out, fp32 accumulate on tcgen05) with per-(pair, channel) statistics.
int fepe_mlp_gemm(const void* X, const void* W, const float* bias, void* Y, float* stats, int B, int Npad,
                  int Nvalid, int K, int Co, void* stream) {
    if (!X || !W || !Y || B <= 0 || Npad <= 0 || (Npad % fepe::kGemmBM) != 0 || Nvalid > Npad ||
        (K % fepe::kGemmBK) != 0 || (Co % 64) != 0)
        return FEPE_E_BADARG;
    if (reinterpret_cast<uintptr_t>(bias) & 15u) return FEPE_E_BADARG;   // read as 16-byte vectors
    fepe::GemmParams p{B * Npad, K, Co, Npad, Nvalid, bias, static_cast<__nv_bfloat16*>(Y), stats, nullptr, 0.f};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // 2 stages / 2 CTAs per SM: the epilogue of one CTA overlaps the main loop of the other.  Measured faster
    // than 4 stages / 1 CTA per SM on every layer (profiles/r1_mlp_timing.txt).  
    // Default for Co % 128 == 0: the persistent kernel (double-buffered TMEM accumulator, 128 x 256 tiles when Co
    // allows).  fepe_set_dispatch(FEPE_DISPATCH_MLP_GEMM, 1) selects the one-tile-per-CTA kernel, 2 forces BN = 128.
    const int mode = fepe::dispatch_get(FEPE_DISPATCH_MLP_GEMM);   // 0 = automatic; 1 = one tile per CTA; 2 = persistent, BN = 128
    const bool tile_mode = mode == 1;
    if (!tile_mode && Co % 128 == 0) {
        const bool force128 = mode == 2;
        if (Co % 256 == 0 && !force128) return fepe::launch_gemm_persist<256, 4, 0>(X, W, p, st);
        return fepe::launch_gemm_persist<128, 6, 0>(X, W, p, st);
    }
    const int stages = 2;
    if (Co % 128 == 0)
        return stages == 2 ? fepe::launch_gemm<128, 2>(X, W, p, st) : fepe::launch_gemm<128, 4>(X, W, p, st);
    return stages == 2 ? fepe::launch_gemm<64, 2>(X, W, p, st) : fepe::launch_gemm<64, 4>(X, W, p, st);
}

// Fused variant for inference: the A operand is the previous layer's PRE-norm output Yprev [B*Npad, K] and the
// InstanceNorm + LeakyReLU of that layer is applied to the operand tiles in shared memory (ss from fepe_mlp_scale_shift).
int fepe_mlp_gemm_norm(const void* Yprev, const float* ss, float slope, const void* W, const float* bias, void* Y,
                       float* stats, int B, int Npad, int Nvalid, int K, int Co, void* stream) {
    if (!Yprev || !ss || !W || !Y || B <= 0 || Npad <= 0 || (Npad % fepe::kGemmBM) != 0 || Nvalid > Npad ||
        (K % fepe::kGemmBK) != 0 || (Co % 128) != 0 || (reinterpret_cast<uintptr_t>(ss) & 15u) ||
        (reinterpret_cast<uintptr_t>(bias) & 15u) || !(slope > 0.f && slope < 1.f))
        return FEPE_E_BADARG;
    fepe::GemmParams p{B * Npad, K, Co, Npad, Nvalid, bias, static_cast<__nv_bfloat16*>(Y), stats, ss, slope};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Default: variant 1 (8 epilogue + 4 transform warps, (a, d) prefetched from global memory).  fepe_set_dispatch(FEPE_DISPATCH_MLP_FUSE, 2)
    // selects variant 2 (4 epilogue + 8 transform warps, (a, d) through a shared-memory slot of the stage): measured
    // equal or slower on every layer (profiles/r1_mlp_fused_norm.md) -- the fused layers are bound by the bytes in flight
    // through the L2 and by shared-memory bandwidth, not by the transform's latency -- kept as the tested alternative.
    if (fepe::dispatch_get(FEPE_DISPATCH_MLP_FUSE) == 2) {
        if (Co % 256 == 0) return fepe::launch_gemm_persist<256, 4, 2>(Yprev, W, p, st);
        return fepe::launch_gemm_persist<128, 6, 2>(Yprev, W, p, st);
    }
    if (Co % 256 == 0) return fepe::launch_gemm_persist<256, 4, 1>(Yprev, W, p, st);
    return fepe::launch_gemm_persist<128, 6, 1>(Yprev, W, p, st);
}

int fepe_mlp_scale_shift(float* stats, const float* gamma, const float* beta, float* ss, int B, int Co, int Nvalid,
                         float eps, int clear_stats, void* stream) {
    if (!stats || !gamma || !beta || !ss || B <= 0 || Co <= 0 || (Co & 1) || Nvalid <= 0 ||
        (reinterpret_cast<uintptr_t>(stats) & 15u) || (reinterpret_cast<uintptr_t>(ss) & 15u))
        return FEPE_E_BADARG;
    const int total = B * (Co / 2);
    fepe::fepe_mlp_scale_shift_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        stats, gamma, beta, reinterpret_cast<float4*>(ss), total, Co, Nvalid, eps, clear_stats);
    return static_cast<int>(cudaGetLastError());
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
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* y = static_cast<__nv_bfloat16*>(Y);
    if (Ci == 4) fepe::fepe_mlp_first_kernel<4><<<grid, 128, 0, st>>>(X0, W, bias, y, stats, B, N, Npad, Ci, Co);
    else if (Ci == 7) fepe::fepe_mlp_first_kernel<7><<<grid, 128, 0, st>>>(X0, W, bias, y, stats, B, N, Npad, Ci, Co);
    else fepe::fepe_mlp_first_kernel<0><<<grid, 128, 0, st>>>(X0, W, bias, y, stats, B, N, Npad, Ci, Co);
    return static_cast<int>(cudaGetLastError());
}

int fepe_mlp_last(const void* X, const float* W, float bias, float* logits, float* weights, int B, int N, int Npad,
                  int Ci, void* stream) {
    if (!X || !W || !logits || !weights || B <= 0 || (Ci & 7)) return FEPE_E_BADARG;
    fepe::fepe_mlp_last_kernel<false><<<B, 256, Npad * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(X), W, bias, logits, weights, N, Npad, Ci, nullptr, 0.f);
    return static_cast<int>(cudaGetLastError());
}

// fepe_mlp_last on the PRE-norm output Y of the last block, with that block's InstanceNorm + LeakyReLU fused in.
int fepe_mlp_last_norm(const void* Y, const float* ss, float slope, const float* W, float bias, float* logits,
                       float* weights, int B, int N, int Npad, int Ci, void* stream) {
    if (!Y || !ss || !W || !logits || !weights || B <= 0 || (Ci & 7) || (reinterpret_cast<uintptr_t>(ss) & 15u) ||
        !(slope > 0.f && slope < 1.f))
        return FEPE_E_BADARG;
    fepe::fepe_mlp_last_kernel<true><<<B, 256, Npad * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(Y), W, bias, logits, weights, N, Npad, Ci, reinterpret_cast<const float4*>(ss),
        slope);
    return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
