// Shared pieces of the fused forward / backward streaming kernels: launch parameters, the per-pair
// Hartley state, the shared-memory ring geometry and the host-side device bookkeeping.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_common.cuh"
#include "fepe_math.cuh"

namespace fepe {

#ifndef FEPE_WARPS
#define FEPE_WARPS 8      // 7 consumers + 1 producer: 256 threads => 255 registers, the fp64 solver does not spill
#endif
constexpr int kMaxWarps = FEPE_WARPS;        // consumers + 1 producer
constexpr int kThreads = kMaxWarps * 32;
constexpr int kMaxStages = 16;
constexpr int kScratchDoubles = 64;          // per consumer warp: Gram entries in, solution out (fepe_fit.cu)

struct FitParams {
    const float* matches;   // [B,N,4]
    const float* weights;   // [B,N]
    int B, N;
    float ax, bx, ay, by;
    float clamp_at;
    float* F_out;           // [B,9]
    float* resid;           // [B,N]
    float* epi;             // [B,N] or null
    double* saved;          // [B,FEPE_SAVED_DOUBLES] or null
    // backward only
    const float* gF;
    const float* gresid;
    const float* gepi;
    float* gweights;
    float* gmatches;        // [B,N,4] or null (coordinate gradient)
    RingLayout ring;
};

__device__ __forceinline__ float approx_sqrt(float x) {   // MUFU-based, <= 2 ulp: ample for distances
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float approx_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Per-pair affine maps derived from the Hartley transforms (all warp-uniform registers).
struct PairNorm {
    float m1x, m1y, m2x, m2y;   // raw means
    float c1x, c1y, c2x, c2y;   // primed centroids  (ax*m+bx)
    float s1, s2;               // Hartley scales 1.4142/meandist (literal as in DeepFNet.py:168)
};


// Warp driver of the multi-shift eigen-solver (scalar pieces and rationale in fepe_math.cuh), on the tridiagonal form:
// every lane reduces G = Q T Q^T (uniform work), a lane's shift is then a tridiagonal LDL^T.  The 28 reflector entries
// wait in `hv_s` (shared memory the caller does not need during the solve) instead of 56 registers.
// All 32 lanes call it with the same g36; returns the number of rounds.
__device__ __forceinline__ int eig9_smallest_warp(const double* __restrict__ g36, double (&f)[9], double& lambda,
                                                  int lane, double* hv_s) {
    double ta[9], tb[8], htau[7];
    {
        double hv[28];
        tridiag9(g36, ta, tb, hv, htau);
        __syncwarp();          // whatever the caller kept in hv_s has been read by every lane
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 28; ++i) hv_s[i] = hv[i];
        }
        __syncwarp();
    }
    Eig9Bracket b;
    const double tr_g = tri9_normalise(ta, tb);
    if (!(tr_g > 0.0) || !tri9_bracket_init(ta, b)) {
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
        lambda = 0.0;
        return 0;
    }
    const double tiny = 1e-18 * b.tr;
    {   // Sturm probes (fepe_math.cuh): 32 shifts each, one ballot, no exchange
        double tb2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tb2[i] = tb[i] * tb[i];
        tri9_probe_begin(b, 32);
#pragma unroll 1
        for (int sub = 0; sub < 5; ++sub) {
            const int cnt = tri9_sturm_count(ta, tb2, tri9_probe_shift(b, lane, 32, sub));
            const unsigned bad = ~__ballot_sync(0xffffffffu, cnt == 0);
            tri9_probe_update(b, bad ? (__ffs(bad) - 1) : 32, 32, sub);
        }
        tri9_probe_finish(b);
    }
    double x[9];
    eig9_start_vector(x);
    double rho = 0.0;
    int rounds = 0;
    while (rounds < 10) {
        const double mu = eig9_lane_shift(b, lane);
        double xl[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) xl[i] = x[i];
        int nneg;
        double rho_l, r_l, c_l;
        tri9_lane_round(ta, tb, mu, tiny, 2, xl, nneg, rho_l, r_l, c_l);
        ++rounds;
        const unsigned ok = __ballot_sync(0xffffffffu, nneg == 0);
        const unsigned bad = ~ok;
        const int first_fail = bad ? (__ffs(bad) - 1) : 32;      // shifts ascend with the lane index
        const int best = first_fail - 1;
        if (best < 0) {        // even the safe shift failed (G indefinite to rounding): move it further down
            b.lo = b.lo * 64.0 - 1e-13 * b.tr;
            b.lo_heur = b.lo;
            continue;
        }
        const double mu_best = __shfl_sync(0xffffffffu, mu, best);
        const double mu_fail = (first_fail < 32) ? __shfl_sync(0xffffffffu, mu, first_fail & 31) : -1.0;
        rho = __shfl_sync(0xffffffffu, rho_l, best);
        const double r = __shfl_sync(0xffffffffu, r_l, best);
        const double c = __shfl_sync(0xffffffffu, c_l, best);
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = __shfl_sync(0xffffffffu, xl[i], best);
        if (eig9_bracket_update(b, mu_best, mu_fail, rho, r, c)) break;
    }
    tridiag9_back(hv_s, htau, x);
    __syncwarp();              // every lane has read the reflectors: the caller may reuse hv_s
    canonical_sign9(x, f);
    lambda = rho * tr_g;
    return rounds;
}

// out = T2^T F2 T1 with T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]] (DeepFNet.py:256), fp64 in, fp32 out.
__device__ __forceinline__ void denormalise_F(const double (&F2)[9], const PairNorm& h, float (&Fo)[9]) {
    const double s1 = h.s1, s2 = h.s2;
    const double t1x = -s1 * h.c1x, t1y = -s1 * h.c1y, t2x = -s2 * h.c2x, t2y = -s2 * h.c2y;
    double A[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        A[3 * r] = F2[3 * r] * s1;
        A[3 * r + 1] = F2[3 * r + 1] * s1;
        A[3 * r + 2] = fma(F2[3 * r], t1x, fma(F2[3 * r + 1], t1y, F2[3 * r + 2]));
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Fo[k] = static_cast<float>(s2 * A[k]);
        Fo[3 + k] = static_cast<float>(s2 * A[3 + k]);
        Fo[6 + k] = static_cast<float>(fma(t2x, A[k], fma(t2y, A[3 + k], A[6 + k])));
    }
}

// Pair-ring data movement.  `n_arrays` per-pair arrays travel into a stage: array 0 = the [N,4]
// coordinates (16 B per correspondence, always 16-byte aligned), arrays 1.. = [N] fp32 rows.  Rows whose
// global address is 16-byte aligned (N % 4 == 0) go by bulk-async copy from the PRODUCER warp; ragged rows
// are copied by the CONSUMER warp that owns the pair, after its full-barrier wait (ring_fill_ragged) --
// writer and reader are then the same warp, ordered by __syncwarp, and no generic-proxy write ever has to
// be published to another warp through an mbarrier (compute-sanitizer racecheck clean).
struct RingSources {
    const float* ptr[4];   // ptr[0]: [B,N,4]; ptr[1..]: [B,N] (null: that slot is never read)
    int n_arrays;
};

__device__ __forceinline__ bool ring_row_is_bulk(const float* base, size_t pair, int N) {
    return base != nullptr && ((reinterpret_cast<uintptr_t>(base + pair * static_cast<size_t>(N)) & 15u) == 0) &&
           ((N & 3) == 0);
}

__device__ __forceinline__ void ring_producer(unsigned char* smem, const RingLayout& ring, uint64_t* full,
                                              uint64_t* empty, const RingSources& src, int N, int n_local, int lane) {
    if (lane != 0) return;
    const int S = ring.stages;
    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;
    const uint32_t row_bytes = static_cast<uint32_t>(N) * 4u;
    for (int j = 0; j < n_local; ++j) {
        const int stage = j % S;
        const uint32_t phase = static_cast<uint32_t>(j / S) & 1u;
        mbar_wait(&empty[stage], phase ^ 1u);
        const size_t pair = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        unsigned char* sb = smem + static_cast<size_t>(stage) * ring.stage_bytes;
        uint32_t tx = pts_bytes;
        for (int a = 1; a < src.n_arrays; ++a)
            if (ring_row_is_bulk(src.ptr[a], pair, N)) tx += row_bytes;
        mbar_arrive_expect_tx(&full[stage], tx);
        bulk_g2s(sb, src.ptr[0] + pair * static_cast<size_t>(N) * 4, pts_bytes, &full[stage]);
        for (int a = 1; a < src.n_arrays; ++a)
            if (ring_row_is_bulk(src.ptr[a], pair, N))
                bulk_g2s(sb + pts_bytes + static_cast<uint32_t>(a - 1) * row_bytes,
                         src.ptr[a] + pair * static_cast<size_t>(N), row_bytes, &full[stage]);
    }
}

// Consumer side of the ragged case: called by the owning warp after mbar_wait(full).
__device__ __forceinline__ void ring_fill_ragged(unsigned char* sb, const RingSources& src, size_t pair, int N,
                                                 int lane) {
    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;
    const uint32_t row_bytes = static_cast<uint32_t>(N) * 4u;
    bool any = false;
    for (int a = 1; a < src.n_arrays; ++a) {
        if (src.ptr[a] == nullptr || ring_row_is_bulk(src.ptr[a], pair, N)) continue;
        float* dst = reinterpret_cast<float*>(sb + pts_bytes + static_cast<uint32_t>(a - 1) * row_bytes);
        const float* g = src.ptr[a] + pair * static_cast<size_t>(N);
        for (int i = lane; i < N; i += 32) dst[i] = __ldg(g + i);
        any = true;
    }
    if (any) __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// host side (inline: one copy per translation unit is fine, the state is per process via statics
// inside device_info())
struct DeviceInfo {
    int ok = 0;
    int sms = 0;
    int smem_optin = 0;
    int fwd_configured = 0;
    int bwd_configured = 0;
    int small_configured = 0;
    int small_pose_configured = 0;
};

inline DeviceInfo& device_info() {
    static DeviceInfo info[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        static DeviceInfo bad;
        return bad;
    }
    DeviceInfo& d = info[dev];
    if (!d.ok) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        d.ok = (major == 10 && d.sms > 0) ? 1 : -1;
    }
    return d;
}

inline bool make_ring(int N, int bytes_per_corr, int smem_limit, RingLayout& r) {
    const int stage = ((N * bytes_per_corr + 127) / 128) * 128;
    const int fixed = 2 * kMaxStages * 8 + (kMaxWarps - 1) * kScratchDoubles * 8 + 256;
    int S = (smem_limit - fixed) / stage;
    if (S > kMaxStages) S = kMaxStages;
    if (S < 2) return false;
    int C = S - 1;                     // keep at least one stage of prefetch
    if (S >= 6) C = S - 2;
    if (C > kMaxWarps - 1) C = kMaxWarps - 1;
    r.stages = S;
    r.consumers = C;
    r.stage_bytes = stage;
    r.bar_off = S * stage;
    r.scratch_off = r.bar_off + 2 * kMaxStages * 8;
    r.total_bytes = r.scratch_off + (kMaxWarps - 1) * kScratchDoubles * 8;
    return true;
}

// Split throughput pipeline (fepe_fit_split.cu); taken from kSplitMinPairsPerSM pairs per SM upwards (measured crossover)
constexpr int kSplitMinPairsPerSM = 16;
bool split_path_supported(const FitParams& p, const DeviceInfo& d);
int launch_split(FitParams p, const DeviceInfo& d, cudaStream_t stream);

}  // namespace fepe
