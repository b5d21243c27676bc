// Throughput path of the weighted 8-point forward for batches that fill the machine: the fused
// one-warp-per-pair kernel (fepe_fit.cu) keeps a pair resident in shared memory through its eigen-solve,
// a ~16 k-cycle serial fp64 chain that only one lane's worth of a warp can use -- with 11 stages per SM
// that residency, not HBM, bounds the launch.  Here the pair leaves shared memory before the solve:
//
//   K1  fepe_gram_kernel    persistent, S shared-memory stages of one pair each filled by TMA bulk copies
//                           (mbarrier completion), a TEAM of T warps per pair: Hartley passes + fp64 Gram pass
//                           out of shared memory; writes the pair's state (Hartley transforms + the 36
//                           distinct Gram entries, 42 doubles).  Self-refilling: the team that finishes with a
//                           stage issues the copy of the pair that will use it next, so there is no producer
//                           warp and no `empty` barrier.  128 registers, 16 working warps per SM instead of 7.
//   K2  fepe_solve_kernel   one LANE per pair (32 different eigenproblems per warp, Gram entries transposed
//                           through shared memory): serial shifted inverse iteration (fepe_math.cuh
//                           eig9_smallest), rank 2, de-normalisation.  No lane does redundant work.
//   K3  fepe_resid_kernel   one CTA per pair, streams the correspondences a second time (from L2 when the
//                           launch is chunked to fit it, else from HBM) and writes both residual rows.
//
// The state lives in the caller's `saved` rows when they are requested (training) and otherwise in the head
// of the pair's own `residual` output row (512 B; K3 reads it before it overwrites the row), so the library
// still allocates nothing.  Reference being replaced: the same as fepe_fit.cu (deepFEPE/models/DeepFNet.py
// :148-179, :181-257, dsac_tools/utils_F.py:400-413).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "fepe_dispatch.cuh"
#include "fepe_fit.cuh"
#include "fepe_fit_passes.cuh"

namespace fepe {

constexpr int kGramConsumerWarps = 16;
constexpr int kGramThreads = kGramConsumerWarps * 32;   // 512 threads => 128 registers per thread
#ifndef FEPE_HTILE
#define FEPE_HTILE 16
#endif
constexpr int kHartleyTile = FEPE_HTILE;
constexpr int kGramFixedBytes = 2 * kMaxStages * 8 + kGramConsumerWarps * (36 * 8 + 8 * 4) + 128;

__device__ __forceinline__ double* pair_state(const FitParams& p, size_t pair) {
    return (p.saved != nullptr) ? p.saved + pair * FEPE_SAVED_DOUBLES
                                : reinterpret_cast<double*>(p.resid + pair * static_cast<size_t>(p.N));
}

// Diagnostics (fepe_debug_trace): one record {kernel id, SM, start ns, end ns} per CTA of the three kernels, to see on
// which SMs and when they run (and, for experiments with several streams, whether kernels really share an SM).  Costs one
// global load per CTA while no buffer is set.
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned int g_trace_cap = 0;
__device__ unsigned int g_trace_n = 0;
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_cta(unsigned long long kid, unsigned long long t0) {
    unsigned long long* buf = g_trace;
    if (buf == nullptr) return;
    const unsigned int i = atomicAdd(&g_trace_n, 1u);
    if (i >= g_trace_cap) return;
    unsigned int sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    buf[i * 4 + 0] = kid; buf[i * 4 + 1] = sm; buf[i * 4 + 2] = t0; buf[i * 4 + 3] = trace_now();
}

template <int T>
__device__ __forceinline__ void team_sync(int team) {
    if constexpr (T == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(T * 32) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// K1: Hartley + Gram.  T warps per pair.
// ------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kGramThreads, 1) fepe_gram_kernel(const FitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.ring.stages;
    constexpr int G = kGramConsumerWarps / T;              // teams (= ring consumers)
    const unsigned long long t_begin = trace_now();
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.ring.bar_off);
    // use counter per stage (second half of the barrier block): the refilling team announces "use k of this stage is on
    // its way" BEFORE the waiter may trust the barrier's parity -- a parity wait alone cannot tell use k from use k-2, and
    // nothing else stops a fast team from reaching a stage two uses early (tests/test_gram_ring_protocol.py)
    volatile int* use_cnt = reinterpret_cast<volatile int*>(full + kMaxStages);
    double* gram_w = reinterpret_cast<double*>(smem + p.ring.scratch_off);             // [16][36]
    float* red = reinterpret_cast<float*>(gram_w + kGramConsumerWarps * 36);           // [16][8]
    const int N = p.N;
    const int n_local = (p.B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);

    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;
    // copy of local pair j into its stage (one thread)
    auto fill = [&](int j) {
        const int stage = j % S;
        const size_t pair = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        unsigned char* sb = smem + static_cast<size_t>(stage) * p.ring.stage_bytes;
        const bool wb = ring_row_is_bulk(p.weights, pair, N);
        use_cnt[stage] = j / S + 1;
        mbar_arrive_expect_tx(&full[stage], pts_bytes + (wb ? static_cast<uint32_t>(N) * 4u : 0u));
        bulk_g2s(sb, p.matches + pair * static_cast<size_t>(N) * 4, pts_bytes, &full[stage]);
        if (wb) bulk_g2s(sb + pts_bytes, p.weights + pair * static_cast<size_t>(N), static_cast<uint32_t>(N) * 4u, &full[stage]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); use_cnt[s] = 0; }
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 0; j < S && j < n_local; ++j) fill(j);
    }

    const int team = warp / T;
    const int wt = warp % T;
    const int tt = wt * 32 + lane;
    constexpr int TS = T * 32;
    const float ax = p.ax, bx = p.bx, ay = p.ay, by = p.by;
    float* red_t = red + team * T * 8;                      // [T][8]: 0..3 sums, 4..5 distances
    double* gram_t = gram_w + team * T * 36;                // [T][36]
    constexpr int kRedStride = 33;                          // doubles per (warp, entry) row: 32 lanes + 1 of padding
    const bool reduce_in_stage = p.ring.stage_bytes >= T * 36 * kRedStride * 8;

    for (int j = team; j < n_local; j += G) {
        const int stage = j % S;
        const uint32_t phase = static_cast<uint32_t>(j / S) & 1u;
        const size_t pair = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        unsigned char* sb = smem + static_cast<size_t>(stage) * p.ring.stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(sb);
        float* sw = reinterpret_cast<float*>(sb + pts_bytes);
        const long long tc0 = clock64();
        while (use_cnt[stage] != j / S + 1) {
        }
        mbar_wait(&full[stage], phase);
        const long long tc1 = clock64();
        if (!ring_row_is_bulk(p.weights, pair, N)) {        // ragged weight row: the team copies it itself
            const float* g = p.weights + pair * static_cast<size_t>(N);
            for (int i = tt; i < N; i += TS) sw[i] = __ldg(g + i);
            team_sync<T>(team);
        }

        PairNorm h;
        float sums[4], dist[2] = {0.f, 0.f};
        const bool tiled = N <= kHartleyTile * TS;           // every thread owns <= kHartleyTile correspondences
        const bool last_only = tiled && N >= (kHartleyTile - 1) * TS;   // only a thread's last element can be beyond N
        if (last_only) {
            float4 q[kHartleyTile];
            tile_load<kHartleyTile, true>(sp, N, tt, TS, q);
            tile_sums(q, sums);
        } else if (tiled) {
            float4 q[kHartleyTile];
            tile_load(sp, N, tt, TS, q);
            tile_sums(q, sums);
        } else {
            pass_sums(sp, N, tt, TS, sums);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sums[k] = warp_sum(sums[k]);
        if constexpr (T > 1) {
            if (lane == 0) { red_t[wt * 8 + 0] = sums[0]; red_t[wt * 8 + 1] = sums[1]; red_t[wt * 8 + 2] = sums[2]; red_t[wt * 8 + 3] = sums[3]; }
            team_sync<T>(team);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < T; ++w) t += red_t[w * 8 + k];
                sums[k] = t;
            }
        }
        finish_norm(h, sums, dist, N, ax, bx, ay, by, false);
        if (last_only) {
            float4 q[kHartleyTile];                          // second round of loads: cheaper than 64 live registers
            tile_load<kHartleyTile, true>(sp, N, tt, TS, q);
            tile_dist<kHartleyTile, true>(q, N, tt, TS, ax, ay, h, dist);
        } else if (tiled) {
            float4 q[kHartleyTile];
            tile_load(sp, N, tt, TS, q);
            tile_dist(q, N, tt, TS, ax, ay, h, dist);
        } else {
            pass_dist(sp, N, tt, TS, ax, ay, h, dist);
        }
        dist[0] = warp_sum(dist[0]);
        dist[1] = warp_sum(dist[1]);
        if constexpr (T > 1) {
            if (lane == 0) { red_t[wt * 8 + 4] = dist[0]; red_t[wt * 8 + 5] = dist[1]; }
            team_sync<T>(team);
            dist[0] = 0.f; dist[1] = 0.f;
#pragma unroll
            for (int w = 0; w < T; ++w) { dist[0] += red_t[w * 8 + 4]; dist[1] += red_t[w * 8 + 5]; }
        }
        finish_norm(h, sums, dist, N, ax, bx, ay, by, true);
        const PairMap m = make_map(h, ax, ay);

        double* st = pair_state(p, pair);
        double acc[36];
        const long long tc2 = clock64();
        pass_gram(sp, sw, N, tt, TS, m, acc);
        const long long tc3 = clock64();
        if (reduce_in_stage) {
            // Cross-lane reduction through the pair's own stage (its correspondences are dead now): every lane
            // stores its 36 partial sums, thread e of the team adds up the T*32 partials of entry e.  A third of
            // the instructions of the shuffle reduce-scatter (no selects), and no per-warp second level.  Rows are
            // padded to 33 doubles: the 16 lanes of a half warp (consecutive entries, same column) then read 16
            // different 8-byte banks with plain immediate offsets -- no rotation arithmetic in front of the loads.
            team_sync<T>(team);                              // all reads of the stage are done
            double* tb = reinterpret_cast<double*>(sb);      // [T*36][33]
#pragma unroll
            for (int e = 0; e < 36; ++e) tb[(wt * 36 + e) * kRedStride + lane] = acc[e];
            team_sync<T>(team);
            const int e = (T == 1) ? lane : tt;              // T == 1: entries 32..35 in a second trip below
            if (e < 36) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < T; ++w) {
                    const double* row = tb + (w * 36 + e) * kRedStride;
#pragma unroll
                    for (int k = 0; k < 32; k += 16) {       // 16 loads in flight, pairwise tree: depth 4 instead of 16
                        double v[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) v[u] = row[k + u];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] += v[u + 8];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] += v[u + 4];
                        tot += (v[0] + v[2]) + (v[1] + v[3]);
                    }
                }
                st[16 + e] = tot;
            }
            if constexpr (T == 1) {
                if (lane < 4) {
                    const double* row = tb + (32 + lane) * kRedStride;
                    double t0 = 0.0, t1 = 0.0;
#pragma unroll
                    for (int k = 0; k < 32; k += 2) { t0 += row[k]; t1 += row[k + 1]; }
                    st[16 + 32 + lane] = t0 + t1;
                }
            }
            if (tt == TS - 1) { st[0] = h.m1x; st[1] = h.m1y; st[2] = h.s1; st[3] = h.m2x; st[4] = h.m2y; st[5] = h.s2; }
            team_sync<T>(team);                              // the stage may be refilled
        } else {
            int base = 0, cnt = 36;
            ReduceScatter<36, 16>::run(acc, lane, base, cnt);
            if (cnt > 0) gram_t[wt * 36 + base] = acc[0];
            if (cnt > 1) gram_t[wt * 36 + base + 1] = acc[1];
            team_sync<T>(team);   // every read of the stage is done, every partial Gram is in shared memory
            if (tt < 36) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < T; ++w) t += gram_t[w * 36 + tt];
                st[16 + tt] = t;
            } else if (tt == 36) {
                st[0] = h.m1x; st[1] = h.m1y; st[2] = h.s1; st[3] = h.m2x; st[4] = h.m2y; st[5] = h.s2;
            }
            if constexpr (T == 1) {
                if (lane < 4) st[16 + 32 + lane] = gram_t[32 + lane];
                if (lane == 4) { st[0] = h.m1x; st[1] = h.m1y; st[2] = h.s1; st[3] = h.m2x; st[4] = h.m2y; st[5] = h.s2; }
                __syncwarp();     // gram_t is rewritten by this warp's next pair
            }
        }
        if (tt == 0 && j + S < n_local) {
            fence_proxy_async();  // the team's generic-proxy accesses to the stage precede the async-proxy refill
            fill(j + S);
        }
        if (tt == 0 && p.saved != nullptr) {   // per-phase SM cycles of this pair (diagnostics, bench.py --phases)
            const long long tc4 = clock64();
            st[56] = static_cast<double>(tc1 - tc0);   // waiting for the bulk copy
            st[57] = static_cast<double>(tc2 - tc1);   // Hartley passes
            st[58] = static_cast<double>(tc3 - tc2);   // Gram pass
            st[59] = static_cast<double>(tc4 - tc3);   // cross-lane reduction + state store
            st[60] = 0.0; st[61] = 0.0; st[62] = 0.0;
        }
    }
    if (g_trace != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) trace_cta(1, t_begin);
    }
}

// ------------------------------------------------------------------------------------------------
// K2: one lane per pair.
// ------------------------------------------------------------------------------------------------
constexpr int kSolveStride = 33;

__global__ void __launch_bounds__(32) fepe_solve_kernel(const FitParams p) {
    __shared__ double g[36 * kSolveStride];
    const unsigned long long t_begin = trace_now();
    const int lane = threadIdx.x;
    const size_t pair0 = static_cast<size_t>(blockIdx.x) * 32;
    const int n_here = min(32, p.B - static_cast<int>(pair0));
    {   // transpose the 32 Gram matrices into [entry][pair]; all 36 loads of a lane are in flight together
        double tmp[36];
#pragma unroll
        for (int it = 0; it < 36; ++it) {
            const int k = it * 32 + lane;
            const int pr = k / 36, e = k - pr * 36;
            tmp[it] = (pr < n_here) ? pair_state(p, pair0 + pr)[16 + e] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < 36; ++it) {
            const int k = it * 32 + lane;
            const int pr = k / 36, e = k - pr * 36;
            g[e * kSolveStride + pr] = tmp[it];
        }
    }
    __syncwarp();
    if (lane >= n_here) return;
    const size_t pair = pair0 + lane;
    double* st = pair_state(p, pair);
    PairNorm h;
    h.m1x = static_cast<float>(st[0]); h.m1y = static_cast<float>(st[1]); h.s1 = static_cast<float>(st[2]);
    h.m2x = static_cast<float>(st[3]); h.m2y = static_cast<float>(st[4]); h.s2 = static_cast<float>(st[5]);
    h.c1x = fmaf(p.ax, h.m1x, p.bx); h.c1y = fmaf(p.ay, h.m1y, p.by);
    h.c2x = fmaf(p.ax, h.m2x, p.bx); h.c2y = fmaf(p.ay, h.m2y, p.by);

    double f[9], lambda;
    const StridedG36<kSolveStride> view{g + lane};
    const int its = eig9_smallest_tri(view, f, lambda);
    double F2[9], v3[3], sigma3;
    rank2_project(f, F2, v3, sigma3);
    float Fo[9];
    denormalise_F(F2, h, Fo);
#pragma unroll
    for (int i = 0; i < 9; ++i) p.F_out[pair * 9 + i] = Fo[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) st[6 + i] = f[i];
    st[15] = lambda;
    st[52] = static_cast<double>(its);
    st[53] = v3[0]; st[54] = v3[1]; st[55] = v3[2];
    st[63] = sigma3;
    if (lane == 0) trace_cta(2, t_begin);
}

// ------------------------------------------------------------------------------------------------
// K3: residual rows.  One CTA per pair.
// ------------------------------------------------------------------------------------------------
constexpr int kResidThreads = 256;
constexpr int kResidPrefetch = 4;

__global__ void __launch_bounds__(kResidThreads) fepe_resid_kernel(const FitParams p) {
    pdl_launch_dependents();
    const unsigned long long t_begin = trace_now();
    // Highest pair first: K1 walks the batch upwards, so the pairs it read last are the ones still in the 126 MB L2.
    const size_t pair = gridDim.x - 1 - blockIdx.x;
    const int N = p.N;
    const int tid = threadIdx.x;
    const float4* gp = reinterpret_cast<const float4*>(p.matches) + pair * static_cast<size_t>(N);
    const float* gw = p.weights + pair * static_cast<size_t>(N);
    // The first kResidPrefetch correspondences of every thread are requested BEFORE the pair's parameters: a CTA
    // lives for a few microseconds, so a dependent parameter round trip in front of the stream would idle it.
    float4 q[kResidPrefetch];
    float wv[kResidPrefetch];
#pragma unroll
    for (int u = 0; u < kResidPrefetch; ++u) {
        const int i = tid + u * kResidThreads;
        if (i < N) { q[u] = __ldcs(gp + i); wv[u] = __ldcs(gw + i); }
    }
    const double* st = pair_state(p, pair);
    PairNorm h;
    h.m1x = static_cast<float>(st[0]); h.m1y = static_cast<float>(st[1]); h.s1 = static_cast<float>(st[2]);
    h.m2x = static_cast<float>(st[3]); h.m2y = static_cast<float>(st[4]); h.s2 = static_cast<float>(st[5]);
    float ff[9], Fo[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { ff[i] = static_cast<float>(st[6 + i]); Fo[i] = p.F_out[pair * 9 + i]; }
    if (p.saved == nullptr) __syncthreads();   // the state sits in the row this CTA is about to overwrite
    const PairMap m = make_map(h, p.ax, p.ay);
    float* r_out = p.resid + pair * static_cast<size_t>(N);
    float* e_out = (p.epi != nullptr) ? p.epi + pair * static_cast<size_t>(N) : nullptr;
#pragma unroll
    for (int u = 0; u < kResidPrefetch; ++u) {
        const int i = tid + u * kResidThreads;
        if (i < N) {
            __stcs(r_out + i, resid_one(q[u], wv[u], m, ff));
            if (e_out != nullptr) __stcs(e_out + i, epi_one(q[u], Fo, p.ax, p.bx, p.ay, p.by, p.clamp_at));
        }
    }
    pass_resid(gp, gw, N, tid + kResidPrefetch * kResidThreads, kResidThreads, m, ff, Fo, p.ax, p.bx, p.ay, p.by,
               p.clamp_at, r_out, e_out);
    if (tid == 0) trace_cta(3, t_begin);
}

template <int T>
static cudaError_t launch_gram(const FitParams& p, int grid, cudaStream_t stream) {
    static int configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(fepe_gram_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             device_info().smem_optin);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = 1;
    }
    fepe_gram_kernel<T><<<grid, kGramThreads, p.ring.total_bytes, stream>>>(p);
    return cudaGetLastError();
}

// Can the split pipeline run this problem?  (state rows need 512 aligned bytes; K1 needs S > teams.)
bool split_path_supported(const FitParams& p, const DeviceInfo& d) {
    if (p.saved == nullptr) {
        if (p.N < 128 || (p.N & 1) || (reinterpret_cast<uintptr_t>(p.resid) & 7u)) return false;
    }
    const int stage = ((p.N * 20 + 127) / 128) * 128;
    return (d.smem_optin - kGramFixedBytes) / stage >= 5;     // 4 teams of 4 warps + one stage of prefetch
}

int launch_split(FitParams p, const DeviceInfo& d, cudaStream_t stream) {
    const int stage = ((p.N * 20 + 127) / 128) * 128;
    int S = (d.smem_optin - kGramFixedBytes) / stage;
    if (S > kMaxStages) S = kMaxStages;
    if (S < 5) return FEPE_E_TOOLARGE;
    // largest number of teams that still leaves >= 2 stages of prefetch
    int T = 1;
    while (T < 4 && kGramConsumerWarps / T > S - 2) T *= 2;
    if (kGramConsumerWarps / T > S - 1) return FEPE_E_TOOLARGE;
    const int force = dispatch_get(FEPE_DISPATCH_GRAM_TEAM);     // 0 = automatic; 1 / 2 / 3 = teams of 1 / 2 / 4 warps
    if (force != 0) {
        const int t = (force == 3) ? 4 : force;
        if (kGramConsumerWarps / t <= S - 1) T = t;
    }
    p.ring.stages = S;
    p.ring.consumers = kGramConsumerWarps / T;
    p.ring.stage_bytes = stage;
    p.ring.bar_off = S * stage;
    p.ring.scratch_off = p.ring.bar_off + 2 * kMaxStages * 8;
    p.ring.total_bytes = p.ring.scratch_off + kGramConsumerWarps * (36 * 8 + 8 * 4);
    const int grid = p.B < d.sms ? p.B : d.sms;
    cudaError_t e = (T == 1)   ? launch_gram<1>(p, grid, stream)
                    : (T == 2) ? launch_gram<2>(p, grid, stream)
                               : launch_gram<4>(p, grid, stream);
    if (e != cudaSuccess) return static_cast<int>(e);
    fepe_solve_kernel<<<(p.B + 31) / 32, 32, 0, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    fepe_resid_kernel<<<p.B, kResidThreads, 0, stream>>>(p);
    return static_cast<int>(cudaGetLastError());
}

}  // namespace fepe

extern "C" int fepe_debug_trace(void* buf, int capacity_records_in) {
    if (capacity_records_in < 0) return FEPE_E_BADARG;
    const unsigned int capacity_records = (buf != nullptr) ? static_cast<unsigned int>(capacity_records_in) : 0u;
    unsigned long long* b = static_cast<unsigned long long*>(buf);
    unsigned int zero = 0;
    cudaError_t e = cudaMemcpyToSymbol(fepe::g_trace, &b, sizeof(b));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(fepe::g_trace_cap, &capacity_records, sizeof(capacity_records));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(fepe::g_trace_n, &zero, sizeof(zero));
    return static_cast<int>(e);
}
extern "C" int fepe_debug_trace_count(void) {
    unsigned int n = 0;
    if (cudaMemcpyFromSymbol(&n, fepe::g_trace_n, sizeof(n)) != cudaSuccess) return -1;
    return static_cast<int>(n);
}
