// fepe_recover_pose: validation pose recovery on the device.
//
// Replaces, per (layer, pair), what the reference does on the host for every validation sample through a pebble
// process pool (deepFEPE/Train_model_pipeline.py:954-964,1048-1061 -> train_good_utils.py:553-646 val_rt ->
// dsac_tools/utils_F.py:909-954 goodCorr_eval_nondecompose):
//     num_inlier, R, t, mask = cv2.recoverPose(E_hat, p1s, p2s, focal=K[0,0], pp=(K[0,2], K[1,2]))   (:936)
//     R_cam, t_cam = invert_Rt(R, t); err_q = rot12_to_angle_error(..); err_t = vector_angle(..)        (:938-940)
// One CTA per (layer, pair).  Work item = (correspondence, candidate): 4 N independent 4x4 symmetric eigenproblems in
// fp64 (fepe_recover.cuh), one per thread and trip; counts through warp ballots, the per-point result bits of all four
// candidates parked in shared memory so the mask of the winner is written without a second triangulation.
// Traffic is 16 B per correspondence in, 1 B out: the kernel is fp64-latency bound and exists to keep the host (and
// its D2H copy of the matches) out of the validation loop.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_recover.cuh"

namespace fepe {

constexpr int kRecoverThreads = 256;

struct RecoverParams {
    const float* E;          // [L,B,9]
    const float* K;          // [B,9]
    const float* matches;    // [B,N,4] pixels
    const int* n_valid;      // [B] or null
    const float* Rt;         // [B,16] scene motion or null
    int L, B, N;
    float thresh;
    float* out;              // [L,B,FEPE_RECOVER_OUT_FLOATS]
    unsigned char* mask;     // [L,B,N] or null
};

__global__ void __launch_bounds__(kRecoverThreads) fepe_recover_pose_kernel(const RecoverParams p) {
    extern __shared__ uint32_t bits[];              // one byte per correspondence: bit k = candidate k is good
    __shared__ int counts[4];
    const int item = blockIdx.x;                    // layer * B + pair
    const int pair = item % p.B;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    int Nv = p.N;
    if (p.n_valid != nullptr) { const int v = p.n_valid[pair]; Nv = v < 0 ? 0 : (v < p.N ? v : p.N); }
    const int words = (p.N + 3) / 4;
    for (int i = tid; i < words; i += kRecoverThreads) bits[i] = 0u;
    if (tid < 4) counts[tid] = 0;

    // warp-uniform: decomposition of E (every thread; a few hundred fp64 operations)
    double E[9], R1[9], R2[9], t[3];
    {
        double U[9], S[3], V[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] = static_cast<double>(__ldg(p.E + static_cast<size_t>(item) * 9 + i));
        essential_decompose(E, R1, R2, t, U, S, V);
    }
    const double focal = static_cast<double>(__ldg(p.K + pair * 9));          // cv2 call site: focal = K[0,0] for both axes
    const double ppx = static_cast<double>(__ldg(p.K + pair * 9 + 2)), ppy = static_cast<double>(__ldg(p.K + pair * 9 + 5));
    const double inv_f = 1.0 / focal;
    const int cand = tid & 3;                       // kRecoverThreads % 4 == 0: a thread keeps its candidate
    double P[12];
    recover_candidate(cand, R1, R2, t, P);
    __syncthreads();

    const float4* gm = reinterpret_cast<const float4*>(p.matches) + static_cast<size_t>(pair) * p.N;
    const double thresh = static_cast<double>(p.thresh);
    int mine = 0;
    const int n_items = 4 * Nv;
    for (int base = 0; base < n_items; base += kRecoverThreads) {
        const int idx = base + tid;
        bool ok = false;
        if (idx < n_items) {
            const int pt = idx >> 2;
            const float4 q = __ldg(gm + pt);
            // cv2 converts the points to double first, then (x - pp) / focal
            const double x1 = (static_cast<double>(q.x) - ppx) * inv_f, y1 = (static_cast<double>(q.y) - ppy) * inv_f;
            const double x2 = (static_cast<double>(q.z) - ppx) * inv_f, y2 = (static_cast<double>(q.w) - ppy) * inv_f;
            double X[4];
            triangulate_dlt(x1, y1, x2, y2, P, X);
            ok = cheirality_ok(X, P, thresh);
            if (ok) atomicOr(&bits[pt >> 2], 1u << (((pt & 3) << 3) + cand));
        }
        mine += ok ? 1 : 0;
    }
    // lanes with equal (lane & 3) share a candidate
    mine += __shfl_xor_sync(0xffffffffu, mine, 4);
    mine += __shfl_xor_sync(0xffffffffu, mine, 8);
    mine += __shfl_xor_sync(0xffffffffu, mine, 16);
    if (lane < 4) atomicAdd(&counts[lane], mine);
    __syncthreads();

    const int g1 = counts[0], g2 = counts[1], g3 = counts[2], g4 = counts[3];
    const int best = recover_select(g1, g2, g3, g4);
    if (p.mask != nullptr) {
        unsigned char* mo = p.mask + static_cast<size_t>(item) * p.N;
        const unsigned char* by = reinterpret_cast<const unsigned char*>(bits);
        for (int i = tid; i < p.N; i += kRecoverThreads) mo[i] = ((by[i] >> best) & 1u) ? 255 : 0;   // cv2 masks hold 0 / 255
    }
    if (tid == 0) {
        double Pb[12];
        recover_candidate(best, R1, R2, t, Pb);
        double R[9], tt[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            R[3 * r] = Pb[4 * r]; R[3 * r + 1] = Pb[4 * r + 1]; R[3 * r + 2] = Pb[4 * r + 2];
            tt[r] = Pb[4 * r + 3];
        }
        if (Nv < 5) {                       // fewer than 5 correspondences: R = I, t = 0 (utils_F.py:948-952)
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
            tt[0] = 0.0; tt[1] = 0.0; tt[2] = 0.0;
        }
        float* o = p.out + static_cast<size_t>(item) * FEPE_RECOVER_OUT_FLOATS;
#pragma unroll
        for (int i = 0; i < 9; ++i) o[i] = static_cast<float>(R[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) o[9 + i] = static_cast<float>(tt[i]);
        o[12] = static_cast<float>(best == 0 ? g1 : best == 1 ? g2 : best == 2 ? g3 : g4);
        o[13] = static_cast<float>(best);
        o[14] = static_cast<float>(g1); o[15] = static_cast<float>(g2); o[16] = static_cast<float>(g3); o[17] = static_cast<float>(g4);
        double eq = 180.0, et = 90.0;      // the reference's values when recovery is impossible (utils_F.py:946-950)
        if (p.Rt != nullptr && Nv >= 5) {
            double Rs[9], ts[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) Rs[3 * r + c] = static_cast<double>(p.Rt[pair * 16 + 4 * r + c]);
                ts[r] = static_cast<double>(p.Rt[pair * 16 + 4 * r + 3]);
            }
            recover_errors(R, tt, Rs, ts, eq, et);
        }
        o[18] = static_cast<float>(eq);
        o[19] = static_cast<float>(et);
        o[20] = static_cast<float>(Nv);
        o[21] = 0.f; o[22] = 0.f; o[23] = 0.f;
    }
}

}  // namespace fepe

extern "C" int fepe_recover_pose(const float* E, const float* K, const float* matches, const int* n_valid, int L, int B,
                                 int N, float distance_thresh, const float* Rt_scene, float* out, unsigned char* mask,
                                 void* stream) {
    if (L == 0 || B == 0) return 0;
    if (!E || !K || !matches || !out || L < 0 || B < 0 || N <= 0) return FEPE_E_BADARG;
    if (reinterpret_cast<uintptr_t>(matches) & 15u) return FEPE_E_BADARG;
    const size_t smem = static_cast<size_t>((N + 3) / 4) * 4;
    if (smem > 200 * 1024) return FEPE_E_TOOLARGE;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fepe::fepe_recover_pose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    fepe::RecoverParams p{E, K, matches, n_valid, Rt_scene, L, B, N, distance_thresh, out, mask};
    fepe::fepe_recover_pose_kernel<<<L * B, fepe::kRecoverThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}
