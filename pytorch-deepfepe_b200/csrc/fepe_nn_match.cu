// fepe_nn_match: mutual-nearest-neighbour descriptor matching on the device -- the step that
// builds `matches_xy_ori` right before the hot path when the keypoint front-end is used.
//
// Replaces, per sample, the host call at deepFEPE/train_good_utils.py:683-691
//     matching_mask = SP_tracker.nn_match_two_way(desc1.T, desc2.T, nn_thresh)      # numpy, after a D2H of both sets
// (PointTracker.nn_match_two_way of the un-vendored `superpoint` package, identical to MagicLeap's published demo code):
//     dmat = sqrt(2 - 2 clip(desc1^T desc2, -1, 1)); row-wise argmin + threshold + the
// column-wise argmin must point back (mutual); matches ordered by the first index.
//
// K1 fepe_nn_dist_kernel   fp32 CUDA-core tile GEMM (128 x 128 x 16, 8 x 8 per thread) over the [N1 x N2] dot
//                          products of one pair; the distance matrix is never written: each tile folds its entries into
//                          a per-row and a per-column running minimum, a 64-bit key (distance bits << 32 | index) so
//                          that atomicMin reproduces argmin's first-occurrence rule exactly.
// K2 fepe_nn_select_kernel threshold + mutual test + ORDERED compaction (block scan) into (idx1, idx2, score, count).
// K1 exists twice.  fepe_nn_dist_tc_kernel (default when D % 64 == 0): the dot products on tcgen05 at fp32 accuracy with
// the split-fp16 arithmetic of the weight MLP (fepe_mlp32.cu): both descriptor tiles arrive as fp32 TMA boxes, are scaled
// by 2^8 and split in place into fp16 hi / lo tiles, a product is three kind::f16 MMAs accumulated in fp32 in tensor memory
// (bf16 / tf32 operands alone would change nearest-neighbour decisions between near-ties).  The tile is computed TWICE,
// as D = A B^T and as D^T = B A^T (the tensor pipe is idle otherwise): one epilogue thread per row of D folds its 128
// entries into the row minimum, one per row of D^T into the column minimum -- no cross-thread reduction.
// fepe_nn_dist_kernel: the same on CUDA cores in plain fp32 (any D % 16 == 0; FEPE_DISPATCH_NN_DIST = 1).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_common.cuh"
#include "fepe_dispatch.cuh"
#include "fepe_split.cuh"
#include "fepe_umma.cuh"

namespace fepe {

constexpr int kNNTile = 128;
constexpr int kNNK = 16;
constexpr int kNNThreads = 256;
constexpr int kNNPad = 4;

struct NNParams {
    const float* d1;       // [B,N1,D]
    const float* d2;       // [B,N2,D]
    const int* n1;         // [B] or null
    const int* n2;         // [B] or null
    int B, N1, N2, D;
    float thresh;
    unsigned long long* rowbest;   // [B,N1]
    unsigned long long* colbest;   // [B,N2]
    int* idx1;             // [B,N1]
    int* idx2;             // [B,N1]
    float* score;          // [B,N1]
    int* count;            // [B]
};

__device__ __forceinline__ unsigned long long nn_key(float dot, int idx) {
    // the reference's arithmetic, in its precision (float32 numpy): sqrt(2 - 2 clip(dot, -1, 1))
    const float c = fminf(fmaxf(dot, -1.0f), 1.0f);
    const float dist = sqrtf(2.0f - 2.0f * c);                      // >= 0: its bit pattern orders like the value
    return (static_cast<unsigned long long>(__float_as_uint(dist)) << 32) | static_cast<unsigned>(idx);
}

__global__ void __launch_bounds__(kNNThreads) fepe_nn_dist_kernel(const NNParams p) {
    __shared__ __align__(16) float As[kNNK][kNNTile + kNNPad];
    __shared__ __align__(16) float Bs[kNNK][kNNTile + kNNPad];
    __shared__ unsigned long long rowmin[kNNTile];
    __shared__ unsigned long long colmin[kNNTile];
    const int b = blockIdx.z;
    const int row0 = blockIdx.y * kNNTile, col0 = blockIdx.x * kNNTile;
    const int n1 = p.n1 ? min(max(p.n1[b], 0), p.N1) : p.N1;
    const int n2 = p.n2 ? min(max(p.n2[b], 0), p.N2) : p.N2;
    if (row0 >= n1 || col0 >= n2) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    if (tid < kNNTile) { rowmin[tid] = ~0ull; colmin[tid] = ~0ull; }

    const float* A = p.d1 + static_cast<size_t>(b) * p.N1 * p.D;
    const float* Bm = p.d2 + static_cast<size_t>(b) * p.N2 * p.D;
    // global -> shared: thread t moves row (t >> 2) and (t >> 2) + 64, k-quad (t & 3), of both operands
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = row0 + lr + 64 * h, c = col0 + lr + 64 * h;
            ra[h] = (r < n1) ? __ldg(reinterpret_cast<const float4*>(A + static_cast<size_t>(r) * p.D + k0 + lk))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = (c < n2) ? __ldg(reinterpret_cast<const float4*>(Bm + static_cast<size_t>(c) * p.D + k0 + lk))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + 64 * h;
            As[lk + 0][r] = ra[h].x; As[lk + 1][r] = ra[h].y; As[lk + 2][r] = ra[h].z; As[lk + 3][r] = ra[h].w;
            Bs[lk + 0][r] = rb[h].x; Bs[lk + 1][r] = rb[h].y; Bs[lk + 2][r] = rb[h].z; Bs[lk + 3][r] = rb[h].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    gload(0);
    for (int k0 = 0; k0 < p.D; k0 += kNNK) {
        __syncthreads();                 // previous tile fully consumed
        sstore();
        __syncthreads();
        if (k0 + kNNK < p.D) gload(k0 + kNNK);      // next tile in flight during the math
#pragma unroll
        for (int k = 0; k < kNNK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }

    // fold the tile into the running row / column minima
    unsigned long long rbest[8], cbest[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { rbest[i] = ~0ull; cbest[i] = ~0ull; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (r < n1 && c < n2) {
                const unsigned long long kr = nn_key(acc[i][j], c);
                const unsigned long long kc = (kr & 0xffffffff00000000ull) | static_cast<unsigned>(r);
                rbest[i] = kr < rbest[i] ? kr : rbest[i];
                cbest[j] = kc < cbest[j] ? kc : cbest[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int lr_ = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
        const int lc_ = (i < 4) ? tx * 4 + i : 64 + tx * 4 + (i - 4);
        if (rbest[i] != ~0ull) atomicMin(&rowmin[lr_], rbest[i]);
        if (cbest[i] != ~0ull) atomicMin(&colmin[lc_], cbest[i]);
    }
    __syncthreads();
    if (tid < kNNTile) {
        if (row0 + tid < n1 && rowmin[tid] != ~0ull)
            atomicMin(p.rowbest + static_cast<size_t>(b) * p.N1 + row0 + tid, rowmin[tid]);
    } else {
        const int c = tid - kNNTile;
        if (col0 + c < n2 && colmin[c] != ~0ull)
            atomicMin(p.colbest + static_cast<size_t>(b) * p.N2 + col0 + c, colmin[c]);
    }
}

// ------------------------------------------------------------------------------------------------
// K1 on tensor cores.  One CTA per 128 x 128 tile of one pair.  16 warps: warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4-7 = epilogue of D (row minima), warps 8-15 = operand transform (one thread per row
// of the two 128 x 64 fp32 tiles of a k-block); warps 8-11 then turn into the epilogue of D^T (column minima).
// Stage = [A box 0 | A box 1 | B box 0 | B box 1], each 128 rows x 32 channels fp32 = 16 KB (128-byte swizzle); after
// the transform [A hi | A lo | B hi | B lo] as 128 rows x 64 fp16 (same geometry: see fepe_mlp32.cu).
// ------------------------------------------------------------------------------------------------
constexpr int kTcThreads = 512;
constexpr int kTcStages = 3;
constexpr int kTcBox = 128 * 128;                    // 16 KB
constexpr int kTcStageBytes = 4 * kTcBox;            // 64 KB
constexpr float kTcScale = 256.f;                    // per operand: unit descriptors -> fp16 hi / lo well inside the normal range

// Row of a tile -> its best entry.  The distance sqrt(2 - 2 clip(dot)) is non-increasing in the dot product, so the row is
// scanned for the LARGEST clipped dot product (first index on equal values: 6 instructions per entry instead of a
// square root and a 64-bit compare) and only the winner becomes a key; keys of different tiles still meet in atomicMin,
// where equal distances resolve to the smaller index as numpy's argmin does.
__device__ __forceinline__ unsigned long long nn_fold_rows(uint32_t tmem_addr, float inv_scale, int first_other, int n_other) {
    float best = -3.0e38f;
    int best_o = -1;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        uint32_t v[32];
        tmem_ld32(tmem_addr + static_cast<uint32_t>(pass * 32), v);
        const int o0 = first_other + pass * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float c = fminf(__uint_as_float(v[j]) * inv_scale, 1.0f);
            const bool take = (o0 + j < n_other) && (c > best);
            best = take ? c : best;
            best_o = take ? o0 + j : best_o;
        }
    }
    return best_o >= 0 ? nn_key(best, best_o) : ~0ull;
}

__global__ void __launch_bounds__(kTcThreads, 1)
fepe_nn_dist_tc_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, const NNParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int b = blockIdx.z;
    const int row0 = blockIdx.y * kNNTile, col0 = blockIdx.x * kNNTile;
    const int n1 = p.n1 ? min(max(p.n1[b], 0), p.N1) : p.N1;
    const int n2 = p.n2 ? min(max(p.n2[b], 0), p.N2) : p.N2;
    if (row0 >= n1 || col0 >= n2) return;            // uniform over the CTA, before any barrier or allocation
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
    uint64_t* empty = full + kTcStages;
    uint64_t* ready = empty + kTcStages;
    uint64_t* tmem_full = ready + kTcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = p.D / 64;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&ready[s], 8); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr float inv_scale = 1.f / (kTcScale * kTcScale);

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                const uint32_t ph = static_cast<uint32_t>(kb / kTcStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                unsigned char* sa = smem + s * kTcStageBytes;
                mbar_arrive_expect_tx(&full[s], kTcStageBytes);
                tma_load_2d(sa, &map1, kb * 64, b * p.N1 + row0, &full[s]);
                tma_load_2d(sa + kTcBox, &map1, kb * 64 + 32, b * p.N1 + row0, &full[s]);
                tma_load_2d(sa + 2 * kTcBox, &map2, kb * 64, b * p.N2 + col0, &full[s]);
                tma_load_2d(sa + 3 * kTcBox, &map2, kb * 64 + 32, b * p.N2 + col0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // D = f32 (bit 4), A = B = f16, both K-major, N = 128 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
        constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(128 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kTcStages;
            const uint32_t ph = static_cast<uint32_t>(kb / kTcStages) & 1u;
            mbar_wait(&ready[s], ph);
            tcgen05_fence_after();
            if (lane == 0) {
                const unsigned char* sa = smem + s * kTcStageBytes;
                const uint64_t ah = umma_desc_k_sw128(sa), al = umma_desc_k_sw128(sa + kTcBox);
                const uint64_t bh = umma_desc_k_sw128(sa + 2 * kTcBox), bl = umma_desc_k_sw128(sa + 3 * kTcBox);
#pragma unroll
                for (int k = 0; k < 4; ++k) {                       // 16 fp16 channels = 32 B per step
                    const uint64_t o = static_cast<uint64_t>(k * 2);
                    const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                    umma_bf16(tmem_base, ah + o, bh + o, idesc, acc);            // D   = desc1 x desc2^T
                    umma_bf16(tmem_base, al + o, bh + o, idesc, 1u);
                    umma_bf16(tmem_base, ah + o, bl + o, idesc, 1u);
                    umma_bf16(tmem_base + 128u, bh + o, ah + o, idesc, acc);     // D^T = desc2 x desc1^T
                    umma_bf16(tmem_base + 128u, bl + o, ah + o, idesc, 1u);
                    umma_bf16(tmem_base + 128u, bh + o, al + o, idesc, 1u);
                }
                tcgen05_commit(&empty[s]);
                if (kb == num_kb - 1) tcgen05_commit(tmem_full);
            }
            __syncwarp();
        }
    } else if (warp >= 8) {
        // ---------------- transform: thread tt = (operand, row) ----------------
        const int tt = static_cast<int>(threadIdx.x) - 256;
        const int row = tt & 127, sw = row & 7;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kTcStages;
            const uint32_t ph = static_cast<uint32_t>(kb / kTcStages) & 1u;
            mbar_wait(&full[s], ph);
            unsigned char* a0 = smem + s * kTcStageBytes + (tt >> 7) * 2 * kTcBox + row * 128;
            unsigned char* a1 = a0 + kTcBox;
            float v[64];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 t0 = *reinterpret_cast<const float4*>(a0 + ((c ^ sw) << 4));
                const float4 t1 = *reinterpret_cast<const float4*>(a1 + ((c ^ sw) << 4));
                v[4 * c + 0] = t0.x * kTcScale; v[4 * c + 1] = t0.y * kTcScale; v[4 * c + 2] = t0.z * kTcScale; v[4 * c + 3] = t0.w * kTcScale;
                v[32 + 4 * c + 0] = t1.x * kTcScale; v[32 + 4 * c + 1] = t1.y * kTcScale;
                v[32 + 4 * c + 2] = t1.z * kTcScale; v[32 + 4 * c + 3] = t1.w * kTcScale;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) m32::split2(v[8 * c + 2 * q], v[8 * c + 2 * q + 1], h[q], l[q]);
                *reinterpret_cast<uint4*>(a0 + ((c ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(a1 + ((c ^ sw) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
        }
        if (warp < 12) {
            // ---------------- epilogue of D^T: thread = column of the tile (a row of desc2) ----------------
            const int c = (warp & 3) * 32 + lane;
            mbar_wait(tmem_full, 0);
            tcgen05_fence_after();
            const unsigned long long best =
                nn_fold_rows(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 128u, inv_scale, row0, n1);
            if (col0 + c < n2 && best != ~0ull) atomicMin(p.colbest + static_cast<size_t>(b) * p.N2 + col0 + c, best);
            tcgen05_fence_before();
        }
    } else if (warp >= 4) {
        // ---------------- epilogue of D: thread = row of the tile (a row of desc1) ----------------
        const int r = (warp & 3) * 32 + lane;
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        const unsigned long long best = nn_fold_rows(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16), inv_scale, col0, n2);
        if (row0 + r < n1 && best != ~0ull) atomicMin(p.rowbest + static_cast<size_t>(b) * p.N1 + row0 + r, best);
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}

// threshold + mutual test + ordered compaction; one CTA per pair
__global__ void __launch_bounds__(kNNThreads) fepe_nn_select_kernel(const NNParams p) {
    __shared__ int warp_tot[kNNThreads / 32];
    __shared__ int base_s;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n1 = p.n1 ? min(max(p.n1[b], 0), p.N1) : p.N1;
    const int n2 = p.n2 ? min(max(p.n2[b], 0), p.N2) : p.N2;
    const unsigned long long* rb = p.rowbest + static_cast<size_t>(b) * p.N1;
    const unsigned long long* cb = p.colbest + static_cast<size_t>(b) * p.N2;
    int* o1 = p.idx1 + static_cast<size_t>(b) * p.N1;
    int* o2 = p.idx2 + static_cast<size_t>(b) * p.N1;
    float* os = p.score + static_cast<size_t>(b) * p.N1;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n1; i0 += kNNThreads) {
        const int i = i0 + tid;
        bool keep = false;
        int j = 0;
        float dist = 0.f;
        if (i < n1 && n2 > 0) {
            const unsigned long long k = rb[i];
            j = static_cast<int>(k & 0xffffffffu);
            dist = __uint_as_float(static_cast<unsigned>(k >> 32));
            keep = (dist < p.thresh) && (static_cast<int>(cb[j] & 0xffffffffu) == i);
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int before = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (keep) { o1[off + before] = i; o2[off + before] = j; os[off + before] = dist; }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < kNNThreads / 32; ++w) t += warp_tot[w];
            base_s += t;
        }
        __syncthreads();
    }
    if (tid == 0) p.count[b] = base_s;
}

}  // namespace fepe

extern "C" size_t fepe_nn_match_workspace_bytes(int B, int N1, int N2) {
    if (B <= 0 || N1 <= 0 || N2 <= 0) return 0;
    return static_cast<size_t>(B) * (static_cast<size_t>(N1) + static_cast<size_t>(N2)) * sizeof(unsigned long long);
}

extern "C" int fepe_nn_match(const float* desc1, const float* desc2, const int* n1, const int* n2, int B, int N1, int N2,
                             int D, float nn_thresh, void* workspace, int* idx1, int* idx2, float* score, int* count,
                             void* stream) {
    if (B == 0) return 0;
    if (!desc1 || !desc2 || !workspace || !idx1 || !idx2 || !score || !count || B < 0 || N1 <= 0 || N2 <= 0 || D <= 0)
        return FEPE_E_BADARG;
    if ((D % fepe::kNNK) != 0) return FEPE_E_BADARG;                 // 16-float k-steps (SuperPoint: D = 256)
    if ((reinterpret_cast<uintptr_t>(desc1) & 15u) || (reinterpret_cast<uintptr_t>(desc2) & 15u) ||
        (reinterpret_cast<uintptr_t>(workspace) & 7u))
        return FEPE_E_BADARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    fepe::NNParams p{};
    p.d1 = desc1; p.d2 = desc2; p.n1 = n1; p.n2 = n2; p.B = B; p.N1 = N1; p.N2 = N2; p.D = D; p.thresh = nn_thresh;
    p.rowbest = static_cast<unsigned long long*>(workspace);
    p.colbest = p.rowbest + static_cast<size_t>(B) * N1;
    p.idx1 = idx1; p.idx2 = idx2; p.score = score; p.count = count;
    cudaError_t e = cudaMemsetAsync(workspace, 0xff, fepe_nn_match_workspace_bytes(B, N1, N2), s);
    if (e != cudaSuccess) return static_cast<int>(e);
    const dim3 grid((N2 + fepe::kNNTile - 1) / fepe::kNNTile, (N1 + fepe::kNNTile - 1) / fepe::kNNTile, B);
    const int force = fepe::dispatch_get(FEPE_DISPATCH_NN_DIST);        // 0 by shape | 1 CUDA cores | 2 tensor cores
    bool tc = (D % 64) == 0 && force != 1 && static_cast<long long>(B) * N1 < (1ll << 31) && static_cast<long long>(B) * N2 < (1ll << 31);
    CUtensorMap m1, m2;
    if (tc) tc = fepe::make_map_2d(&m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, desc1, B * N1, D, 32, 128) &&
                 fepe::make_map_2d(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, desc2, B * N2, D, 32, 128);
    if (tc) {
        constexpr int smem = fepe::kTcStages * fepe::kTcStageBytes + 256 + 1024;
        static bool configured[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63]) {
            e = cudaFuncSetAttribute(fepe::fepe_nn_dist_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return static_cast<int>(e);
            configured[dev & 63] = true;
        }
        fepe::fepe_nn_dist_tc_kernel<<<grid, fepe::kTcThreads, smem, s>>>(m1, m2, p);
    } else {
        fepe::fepe_nn_dist_kernel<<<grid, fepe::kNNThreads, 0, s>>>(p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    fepe::fepe_nn_select_kernel<<<B, fepe::kNNThreads, 0, s>>>(p);
    return static_cast<int>(cudaGetLastError());
}
