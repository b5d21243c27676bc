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
// fp32 on CUDA cores, not tensor cores: nearest-neighbour decisions between near-ties would change under bf16 / tf32
// rounding of the operands, and the whole step is ~0.6 GFLOP per pair.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"

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
    fepe::fepe_nn_dist_kernel<<<grid, fepe::kNNThreads, 0, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    fepe::fepe_nn_select_kernel<<<B, fepe::kNNThreads, 0, s>>>(p);
    return static_cast<int>(cudaGetLastError());
}
