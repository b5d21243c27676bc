// Fused weighted 8-point forward:  Hartley normalisation -> constraint rows -> 9x9 Gram ->
// smallest eigenvector -> rank-2 projection -> de-normalisation -> algebraic + epipolar residuals.
// ONE kernel, one warp per image pair, correspondences read from HBM exactly once (bulk-async copy
// into shared memory), every later pass out of shared memory.
//
// Reference being replaced (deepFEPE/models/DeepFNet.py): Fit.normalize :148-179 (x2),
// Fit.weighted_svd :181-257 (a Python loop of 2 torch.svd per pair), and
// deepFEPE/dsac_tools/utils_F.py:400-413 compute_epi_residual.  See include/fepe_b200.h.
//
// Numerics.  The reference takes the SVD of the N x 9 matrix X in fp32.  Forming G = X^T X squares
// the condition number, which fp32 cannot afford, so the 36 distinct entries of G
// are accumulated in fp64: the per-correspondence factors are computed in fp32 (same rounding level
// as the reference's X), converted once (5 conversions), and all products / sums are fp64 (exact
// products of fp32 values).  The 9x9 eigenproblem and the 3x3 SVD are fp64 (fepe_math.cuh).  The
// streaming passes (Hartley sums, residuals, epipolar distances) are fp32 like the reference.
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdlib.h>

#include "fepe_dispatch.cuh"
#include "fepe_fit.cuh"
#include "fepe_fit_passes.cuh"
#include "fepe_pose_head.cuh"

namespace fepe {

// Per-warp scratch in shared memory (kScratchDoubles doubles): the Gram entries on the way in, the
// solution on the way out.  Going through shared memory instead of a local-memory struct matters: with
// ~227 KB of shared memory per CTA only ~28 KB of L1 is left and local loads miss it half of the time.
//   [0..35] g36   [36..44] f   [45] lambda   [46..48] v3   [49] sigma3   [50] rounds   [51] eig cycles
//   [52..56] Fo as 9 floats (+pad)   [57..61] Hartley state as 10 floats
constexpr int kSolF = 36, kSolLambda = 45, kSolV3 = 46, kSolSigma3 = 49, kSolRounds = 50, kSolCycles = 51,
              kSolFo = 52, kSolNorm = 57;

// Smallest eigenvector (multi-shift, all 32 lanes cooperate), rank-2 projection, de-normalisation.
// Out of line so that its fp64 working set does not inflate the registers of the streaming loops.
__device__ __noinline__ void solve_pair(double* __restrict__ scratch, int lane) {
    PairNorm h;
    {
        const float* hn = reinterpret_cast<const float*>(scratch + kSolNorm);
        h.m1x = hn[0]; h.m1y = hn[1]; h.m2x = hn[2]; h.m2y = hn[3]; h.c1x = hn[4]; h.c1y = hn[5];
        h.c2x = hn[6]; h.c2y = hn[7]; h.s1 = hn[8]; h.s2 = hn[9];
    }
    double f[9], lambda;
    const long long t0 = clock64();
    const int rounds = eig9_smallest_warp(scratch, f, lambda, lane, scratch + kSolF);   // slots 36..63 are free until the results land
    const long long t1 = clock64();
    double F2[9], v3[3], sigma3;
    rank2_project(f, F2, v3, sigma3);
    float Fo[9];
    denormalise_F(F2, h, Fo);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) scratch[kSolF + i] = f[i];
        scratch[kSolLambda] = lambda;
        scratch[kSolV3] = v3[0]; scratch[kSolV3 + 1] = v3[1]; scratch[kSolV3 + 2] = v3[2];
        scratch[kSolSigma3] = sigma3;
        scratch[kSolRounds] = static_cast<double>(rounds);
        scratch[kSolCycles] = static_cast<double>(t1 - t0);
        float* fo = reinterpret_cast<float*>(scratch + kSolFo);
#pragma unroll
        for (int i = 0; i < 9; ++i) fo[i] = Fo[i];
    }
    __syncwarp();
}

__device__ __forceinline__ void publish_norm(double* scratch, const PairNorm& h, int lane) {
    if (lane == 0) {
        float* hn = reinterpret_cast<float*>(scratch + kSolNorm);
        hn[0] = h.m1x; hn[1] = h.m1y; hn[2] = h.m2x; hn[3] = h.m2y; hn[4] = h.c1x; hn[5] = h.c1y;
        hn[6] = h.c2x; hn[7] = h.c2y; hn[8] = h.s1; hn[9] = h.s2;
    }
}

__device__ __forceinline__ void store_pair_state(const FitParams& p, size_t pair, const PairNorm& h,
                                                 const double* scratch, int lane) {
    const float* fo = reinterpret_cast<const float*>(scratch + kSolFo);
    if (lane < 9) p.F_out[pair * 9 + lane] = fo[lane];
    if (p.saved != nullptr) {
        double* sv = p.saved + pair * FEPE_SAVED_DOUBLES;
        if (lane == 0) {
            sv[0] = h.m1x; sv[1] = h.m1y; sv[2] = h.s1; sv[3] = h.m2x; sv[4] = h.m2y; sv[5] = h.s2;
            sv[52] = scratch[kSolRounds];
            sv[63] = scratch[kSolSigma3];
        }
        if (lane < 9) sv[6 + lane] = scratch[kSolF + lane];
        if (lane == 9) sv[15] = scratch[kSolLambda];
        if (lane >= 10 && lane < 13) sv[53 + lane - 10] = scratch[kSolV3 + lane - 10];
        for (int i = lane; i < 36; i += 32) sv[16 + i] = scratch[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Throughput kernel: persistent pair ring, one warp per pair (see fepe_common.cuh).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) fepe_fit_fwd_kernel(const FitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int S = p.ring.stages;
    const int C = p.ring.consumers;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.ring.bar_off);
    uint64_t* empty = full + S;
    const int N = p.N;
    const int n_local = (p.B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                        static_cast<int>(gridDim.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    __syncthreads();

    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;

    if (warp == C) {
        // ---------------- producer warp ----------------
        ring_producer(smem, p.ring, full, empty, RingSources{{p.matches, p.weights, nullptr, nullptr}, 2}, N, n_local,
                      lane);
        return;
    }
    if (warp > C) return;

    // ---------------- consumer warps: one image pair at a time ----------------
    double* gram = reinterpret_cast<double*>(smem + p.ring.scratch_off) + warp * kScratchDoubles;
    const float ax = p.ax, bx = p.bx, ay = p.ay, by = p.by;

    for (int j = warp; j < n_local; j += C) {
        const int stage = j % S;
        const uint32_t phase = static_cast<uint32_t>(j / S) & 1u;
        const size_t pair = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        unsigned char* sb = smem + static_cast<size_t>(stage) * p.ring.stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(sb);
        const float* sw = reinterpret_cast<const float*>(sb + pts_bytes);
        const long long tc0 = clock64();
        mbar_wait(&full[stage], phase);
        ring_fill_ragged(sb, RingSources{{p.matches, p.weights, nullptr, nullptr}, 2}, pair, N, lane);
        const long long tc1 = clock64();

        // ---- passes 1+2: Hartley transforms of both images (Fit.normalize with unit weights) ----
        PairNorm h;
        float sums[4], dist[2] = {0.f, 0.f};
        pass_sums(sp, N, lane, 32, sums);
#pragma unroll
        for (int k = 0; k < 4; ++k) sums[k] = warp_sum(sums[k]);
        finish_norm(h, sums, dist, N, ax, bx, ay, by, false);
        pass_dist(sp, N, lane, 32, ax, ay, h, dist);
        dist[0] = warp_sum(dist[0]);
        dist[1] = warp_sum(dist[1]);
        finish_norm(h, sums, dist, N, ax, bx, ay, by, true);
        const PairMap m = make_map(h, ax, ay);

        const long long tc2 = clock64();
        // ---- pass 3: Gram, then warp reduce-scatter into shared memory ----
        double acc[36];
        pass_gram(sp, sw, N, lane, 32, m, acc);
        const long long tc3 = clock64();
        {
            int base = 0, cnt = 36;
            ReduceScatter<36, 16>::run(acc, lane, base, cnt);
            if (cnt > 0) gram[base] = acc[0];
            if (cnt > 1) gram[base + 1] = acc[1];
        }
        __syncwarp();
        const long long tc3b = clock64();

        // ---- eigenvector, rank 2, de-normalisation ----
        publish_norm(gram, h, lane);
        __syncwarp();
        solve_pair(gram, lane);
        const long long tc4 = clock64();
        store_pair_state(p, pair, h, gram, lane);

        // ---- pass 4: residual and clamped epipolar distance ----
        float ff[9], Fo[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            ff[i] = static_cast<float>(gram[kSolF + i]);
            Fo[i] = reinterpret_cast<const float*>(gram + kSolFo)[i];
        }
        pass_resid(sp, sw, N, lane, 32, m, ff, Fo, ax, bx, ay, by, p.clamp_at,
                   p.resid + pair * static_cast<size_t>(N),
                   (p.epi != nullptr) ? p.epi + pair * static_cast<size_t>(N) : nullptr);
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[stage]);
            if (p.saved != nullptr) {   // per-phase SM cycles of this pair (diagnostics, see bench.py --phases)
                double* sv = p.saved + pair * FEPE_SAVED_DOUBLES;
                const long long tc5 = clock64();
                sv[56] = static_cast<double>(tc1 - tc0);   // waiting for the bulk copy
                sv[57] = static_cast<double>(tc2 - tc1);   // Hartley passes
                sv[58] = static_cast<double>(tc3 - tc2);   // Gram pass
                sv[59] = static_cast<double>(tc4 - tc3);   // reduce + eigen + rank 2
                sv[60] = static_cast<double>(tc5 - tc4);   // residual pass
                sv[61] = static_cast<double>(tc3b - tc3);  // of which: warp reduce-scatter of the Gram
                sv[62] = gram[kSolCycles];                 // of which: eigen iteration
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Latency kernel for small batches (B <= ~2 pairs per SM): one CTA of kSmallWarps warps per pair, the
// correspondences split across all its threads, block-level reductions through shared memory, the
// small solve on warp 0.  A 256-pair launch is bound by the latency of ONE pair, so the streaming
// passes are spread over 4 warps instead of 1; two such CTAs fit an SM (registers), which covers
// B = 256 on 148 SMs in a single wave.
// ------------------------------------------------------------------------------------------------
// Multi-shift eigen-solver spread over ALL warps of a CTA (latency kernel): NW*32 shifts per round instead of 32,
// so lambda_min is bracketed NW*32-fold per round and fewer rounds are needed (host emulation:
// tests/host_shim.cpp shim_eig9_multishift_n).  Every thread calls it with the same g36 (shared memory) and gets the
// same f / lambda back; `okm` (NW words) and `xch` (16 doubles) are shared-memory exchange buffers.
// Tridiagonal form (fepe_math.cuh tridiag9): warp 0 reduces G = Q T Q^T once and parks T and the reflectors in `tri`
// (52 doubles of shared memory); a lane's shift then costs a tridiagonal LDL^T (8 multipliers, 9 pivots) instead of a
// dense 9x9 one, and the eigenvector of T is carried back with the reflectors at the end.
constexpr int kTriDoubles = 52;      // ta[9] tb[8] hv[28] htau[7]
template <int NW>
__device__ __forceinline__ int eig9_smallest_cta(const double* __restrict__ g36, double (&f)[9], double& lambda,
                                                 int warp, int lane, unsigned* okm, double* xch, double* tri) {
    if (warp == 0) {
        double ta[9], tb[8], hv[28], htau[7];
        tridiag9(g36, ta, tb, hv, htau);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) tri[i] = ta[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) tri[9 + i] = tb[i];
#pragma unroll
            for (int i = 0; i < 28; ++i) tri[17 + i] = hv[i];
#pragma unroll
            for (int i = 0; i < 7; ++i) tri[45 + i] = htau[i];
        }
    }
    __syncthreads();
    double ta[9], tb[8];
#pragma unroll
    for (int i = 0; i < 9; ++i) ta[i] = tri[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) tb[i] = tri[9 + i];
    Eig9Bracket b;
    const double tr_g = tri9_normalise(ta, tb);
    if (!(tr_g > 0.0) || !tri9_bracket_init(ta, b)) {
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
        lambda = 0.0;
        return 0;
    }
    const double tiny = 1e-18 * b.tr;
    const int L = warp * 32 + lane;
    // Sturm probes: NW*32 shifts per probe, nothing exchanged but the ballots (double-buffered: one barrier per probe).
    // One geometric + kProbes-1 linear probes leave lambda_min bracketed to ~1e-6 relative, so the first full round
    // below -- factorisation + two solves per lane -- already converges for any ordinary eigen-gap.
    {
        double tb2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tb2[i] = tb[i] * tb[i];
        constexpr int kProbes = (NW >= 2) ? 4 : 5;
        tri9_probe_begin(b, NW * 32);
#pragma unroll 1
        for (int sub = 0; sub < kProbes; ++sub) {
            const int cnt = tri9_sturm_count(ta, tb2, tri9_probe_shift(b, L, NW * 32, sub));
            const unsigned ok = __ballot_sync(0xffffffffu, cnt == 0);
            unsigned* okb = okm + (sub & 1) * NW;
            if (lane == 0) okb[warp] = ok;
            __syncthreads();
            int first_fail = NW * 32;
#pragma unroll
            for (int w = NW - 1; w >= 0; --w) {
                const unsigned bad = ~okb[w];
                if (bad) first_fail = w * 32 + __ffs(bad) - 1;
            }
            tri9_probe_update(b, first_fail, NW * 32, sub);
        }
        tri9_probe_finish(b);
        __syncthreads();       // the full rounds reuse okm[0..NW)
    }
    double x[9];
    eig9_start_vector(x);
    double rho = 0.0;
    int rounds = 0;
    while (rounds < 10) {
        const double mu = eig9_lane_shift(b, L, NW * 32);
        double xl[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) xl[i] = x[i];
        int nneg;
        double rho_l, r_l, c_l;
        tri9_lane_round(ta, tb, mu, tiny, 2, xl, nneg, rho_l, r_l, c_l);
        ++rounds;
        const unsigned ok = __ballot_sync(0xffffffffu, nneg == 0);
        if (lane == 0) okm[warp] = ok;
        __syncthreads();
        int first_fail = NW * 32;                           // shifts ascend with the CTA-wide lane index
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            const unsigned bad = ~okm[w];
            if (bad) first_fail = w * 32 + __ffs(bad) - 1;
        }
        const int best = first_fail - 1;
        if (best < 0) {        // even the safe shift failed (G indefinite to rounding): move it further down
            b.lo = b.lo * 64.0 - 1e-13 * b.tr;
            b.lo_heur = b.lo;
            __syncthreads();   // okm is rewritten next round
            continue;
        }
        if (L == best) {
            xch[0] = mu; xch[1] = rho_l; xch[2] = r_l; xch[3] = c_l;
#pragma unroll
            for (int i = 0; i < 9; ++i) xch[4 + i] = xl[i];
        }
        if (L == first_fail) xch[13] = mu;
        __syncthreads();
        const double mu_best = xch[0];
        const double mu_fail = (first_fail < NW * 32) ? xch[13] : -1.0;
        rho = xch[1];
        const double r = xch[2], c = xch[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = xch[4 + i];
        if (eig9_bracket_update(b, mu_best, mu_fail, rho, r, c)) break;
    }
    tridiag9_back(tri + 17, tri + 45, x);
    canonical_sign9(x, f);
    lambda = rho * tr_g;
    return rounds;
}

// 2 warps x 6 CTAs per SM (168 registers, ~200 B of spills in the solver) instead of 4 x 2 (226 registers): a single
// launch is ~9 % slower (21.9 vs 20.1 us at 256 pairs), but independent launches that overlap -- the steady state of a
// pipelined caller and of bench.py -- gain 35 % (25.8 M vs 19.1 M pairs/s): half as many lanes repeat the eigen-solve
// and three times as many CTAs hide each other's latency (profiles/r1_small_kernel_variants.txt).
#ifndef FEPE_SMALL_WARPS
#define FEPE_SMALL_WARPS 2
#endif
#ifndef FEPE_SMALL_MINBLOCKS
#define FEPE_SMALL_MINBLOCKS 6
#endif
constexpr int kSmallWarps = FEPE_SMALL_WARPS;
constexpr int kSmallThreads = kSmallWarps * 32;

// POSE: the pose / loss head of the pair (fepe_pose_head.cuh, L = 1) runs on the last warp of the CTA while the
// other warps write the residual rows -- F never leaves the SM before E -> R,t is done (fepe_fit_pose_fwd).
template <bool POSE>
__global__ void __launch_bounds__(kSmallThreads, FEPE_SMALL_MINBLOCKS) fepe_fit_fwd_small_kernel(const FitParams p,
                                                                                               const PoseParams pp) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar;
    __shared__ float red[kSmallWarps][4];
    __shared__ double gram_w[kSmallWarps][36];
    __shared__ double gram[kScratchDoubles];
    __shared__ unsigned eig_ok[2 * kSmallWarps];
    __shared__ double eig_xch[16];
    __shared__ double eig_tri[kTriDoubles];
    __shared__ float pose_in[POSE ? 32 : 1];                                // K(9) q(4) t(3) R_scene(9)
    __shared__ float pose_virt[POSE ? 2 * kVirtPerLane * 3 * 32 : 1];      // [img][u][c][lane]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tid = threadIdx.x;
    const int N = p.N;
    const size_t pair = blockIdx.x;
    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;
    const uint32_t w_bytes = static_cast<uint32_t>(N) * 4u;
    const float4* sp = reinterpret_cast<const float4*>(smem);
    float* sw = reinterpret_cast<float*>(smem + pts_bytes);
    const float* gw = p.weights + pair * static_cast<size_t>(N);
    const bool w_bulk = ((reinterpret_cast<uintptr_t>(gw) & 15u) == 0) && ((N & 3) == 0);
    const float ax = p.ax, bx = p.bx, ay = p.ay, by = p.by;

    const long long ts0 = clock64();
    if (tid == 0) {
        mbar_init(&full_bar, 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar, pts_bytes + (w_bulk ? w_bytes : 0u));
        bulk_g2s(smem, p.matches + pair * static_cast<size_t>(N) * 4, pts_bytes, &full_bar);
        if (w_bulk) bulk_g2s(smem + pts_bytes, gw, w_bytes, &full_bar);
    }
    if (!w_bulk) {
        for (int i = tid; i < N; i += kSmallThreads) sw[i] = __ldg(gw + i);
    }
    if constexpr (POSE) {
        // the pose head's inputs are fetched now, behind the bulk copy, and parked in shared memory
        if (warp == kSmallWarps - 1) {
            float v = 0.f;
            if (lane < 9) v = __ldg(pp.K + pair * 9 + lane);
            else if (lane < 13) v = __ldg(pp.q_gt + pair * 4 + (lane - 9));
            else if (lane < 16) v = __ldg(pp.t_gt + pair * 3 + (lane - 13));
            else if (lane < 25 && pp.Rt != nullptr) v = __ldg(pp.Rt + pair * 16 + 4 * ((lane - 16) / 3) + (lane - 16) % 3);
            pose_in[lane] = v;
            const bool has_virt = (pp.virt1 != nullptr) && (pp.V > 0);
#pragma unroll
            for (int u = 0; u < kVirtPerLane; ++u) {
                const int i = lane + 32 * u;
                const bool live = has_virt && i < pp.V;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    pose_virt[((0 * kVirtPerLane + u) * 3 + c) * 32 + lane] = live ? __ldg(pp.virt1 + (pair * pp.V + i) * 3 + c) : 0.f;
                    pose_virt[((1 * kVirtPerLane + u) * 3 + c) * 32 + lane] = live ? __ldg(pp.virt2 + (pair * pp.V + i) * 3 + c) : 0.f;
                }
            }
        }
    }
    mbar_wait(&full_bar, 0);
    __syncthreads();       // also orders the hand-copied weights
    const long long ts1 = clock64();

    // ---- passes 1+2 with block reductions ----
    PairNorm h;
    float sums[4], dist[2] = {0.f, 0.f};
    constexpr int kTile = 8;
    const bool tiled = N <= kTile * kSmallThreads;       // register-tile passes: one round of loads, 8-way ILP
    if (tiled) {
        float4 q[kTile];
        tile_load(sp, N, tid, kSmallThreads, q);
        tile_sums(q, sums);
    } else {
        pass_sums(sp, N, tid, kSmallThreads, sums);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) sums[k] = warp_sum(sums[k]);
    if (lane == 0) { red[warp][0] = sums[0]; red[warp][1] = sums[1]; red[warp][2] = sums[2]; red[warp][3] = sums[3]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kSmallWarps; ++w) t += red[w][k];
        sums[k] = t;
    }
    finish_norm(h, sums, dist, N, ax, bx, ay, by, false);
    __syncthreads();
    if (tiled) {
        float4 q[kTile];
        tile_load(sp, N, tid, kSmallThreads, q);
        tile_dist(q, N, tid, kSmallThreads, ax, ay, h, dist);
    } else {
        pass_dist(sp, N, tid, kSmallThreads, ax, ay, h, dist);
    }
    dist[0] = warp_sum(dist[0]);
    dist[1] = warp_sum(dist[1]);
    if (lane == 0) { red[warp][0] = dist[0]; red[warp][1] = dist[1]; }
    __syncthreads();
    dist[0] = 0.f; dist[1] = 0.f;
#pragma unroll
    for (int w = 0; w < kSmallWarps; ++w) { dist[0] += red[w][0]; dist[1] += red[w][1]; }
    finish_norm(h, sums, dist, N, ax, bx, ay, by, true);
    const PairMap m = make_map(h, ax, ay);
    const long long ts2 = clock64();

    // ---- pass 3 ----
    {
        double acc[36];
        pass_gram(sp, sw, N, tid, kSmallThreads, m, acc);
        int base = 0, cnt = 36;
        ReduceScatter<36, 16>::run(acc, lane, base, cnt);
        if (cnt > 0) gram_w[warp][base] = acc[0];
        if (cnt > 1) gram_w[warp][base + 1] = acc[1];
    }
    __syncthreads();
    if (tid < 36) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kSmallWarps; ++w) t += gram_w[w][tid];
        gram[tid] = t;
    }
    __syncthreads();

    const long long ts3 = clock64();
    // ---- eigenvector on all warps (4 x 32 shifts per round), the small uniform tail redundantly on every thread ----
    float ff[9], Fo[9];
    {
        double f[9], lambda;
        const long long te0 = clock64();
        const int rounds = eig9_smallest_cta<kSmallWarps>(gram, f, lambda, warp, lane, eig_ok, eig_xch, eig_tri);
        const long long te1 = clock64();
        double F2[9], v3[3], sigma3;
        rank2_project(f, F2, v3, sigma3);
        denormalise_F(F2, h, Fo);
#pragma unroll
        for (int i = 0; i < 9; ++i) ff[i] = static_cast<float>(f[i]);
        if (warp == 0) {
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 9; ++i) gram[kSolF + i] = f[i];
                gram[kSolLambda] = lambda;
                gram[kSolV3] = v3[0]; gram[kSolV3 + 1] = v3[1]; gram[kSolV3 + 2] = v3[2];
                gram[kSolSigma3] = sigma3;
                gram[kSolRounds] = static_cast<double>(rounds);
                gram[kSolCycles] = static_cast<double>(te1 - te0);
                float* fo = reinterpret_cast<float*>(gram + kSolFo);
#pragma unroll
                for (int i = 0; i < 9; ++i) fo[i] = Fo[i];
            }
            __syncwarp();
            store_pair_state(p, pair, h, gram, lane);
        }
    }
    const long long ts4 = clock64();

    // ---- pass 4 (and, fused, the pose head on the last warp) ----
    if constexpr (POSE) {
        if (warp == kSmallWarps - 1) {
            float Kf[9], qgf[4], tgf[3], rtf[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) { Kf[i] = pose_in[i]; rtf[i] = pose_in[16 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) qgf[i] = pose_in[9 + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) tgf[i] = pose_in[13 + i];
            float loss = 0.f;
            if (pp.virt1 != nullptr && pp.V > 0) {
#pragma unroll
                for (int u = 0; u < kVirtPerLane; ++u) {
                    if (lane + 32 * u < pp.V)
                        loss += virt_term(Fo, ax, bx, ay, by, pp.clamp_at,
                                          pose_virt[((0 * kVirtPerLane + u) * 3 + 0) * 32 + lane],
                                          pose_virt[((0 * kVirtPerLane + u) * 3 + 1) * 32 + lane],
                                          pose_virt[((0 * kVirtPerLane + u) * 3 + 2) * 32 + lane],
                                          pose_virt[((1 * kVirtPerLane + u) * 3 + 0) * 32 + lane],
                                          pose_virt[((1 * kVirtPerLane + u) * 3 + 1) * 32 + lane],
                                          pose_virt[((1 * kVirtPerLane + u) * 3 + 2) * 32 + lane]);
                }
                const float* v1 = pp.virt1 + pair * pp.V * 3;
                const float* v2 = pp.virt2 + pair * pp.V * 3;
                for (int i = lane + 32 * kVirtPerLane; i < pp.V; i += 32)
                    loss += virt_term(Fo, ax, bx, ay, by, pp.clamp_at, v1[3 * i], v1[3 * i + 1], v1[3 * i + 2],
                                      v2[3 * i], v2[3 * i + 1], v2[3 * i + 2]);
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o2);
                loss /= static_cast<float>(pp.V);
            }
            pose_head(Fo, Kf, qgf, tgf, pp.Rt != nullptr, rtf, loss, ax, bx, ay, by, lane,
                      pp.out + pair * FEPE_POSE_OUT_FLOATS);
        } else {
            pass_resid(sp, sw, N, tid, kSmallThreads - 32, m, ff, Fo, ax, bx, ay, by, p.clamp_at,
                       p.resid + pair * static_cast<size_t>(N),
                       (p.epi != nullptr) ? p.epi + pair * static_cast<size_t>(N) : nullptr);
        }
    } else {
        pass_resid(sp, sw, N, tid, kSmallThreads, m, ff, Fo, ax, bx, ay, by, p.clamp_at,
                   p.resid + pair * static_cast<size_t>(N),
                   (p.epi != nullptr) ? p.epi + pair * static_cast<size_t>(N) : nullptr);
    }
    if (tid == 0 && p.saved != nullptr) {      // per-phase SM cycles of this pair (diagnostics)
        double* sv = p.saved + pair * FEPE_SAVED_DOUBLES;
        const long long ts5 = clock64();
        sv[56] = static_cast<double>(ts1 - ts0);   // bulk copy
        sv[57] = static_cast<double>(ts2 - ts1);   // Hartley passes
        sv[58] = static_cast<double>(ts3 - ts2);   // Gram pass + reduction
        sv[59] = static_cast<double>(ts4 - ts3);   // eigen + rank 2
        sv[60] = static_cast<double>(ts5 - ts4);   // residual pass
        sv[61] = 0.0;
        sv[62] = gram[kSolCycles];
    }
}

}  // namespace fepe

namespace fepe {
static std::atomic<int> g_dispatch[FEPE_DISPATCH_COUNT];
int dispatch_get(int which) { return g_dispatch[which].load(std::memory_order_relaxed); }
}  // namespace fepe

extern "C" {

const char* fepe_version(void) { return "fepe_b200 0.2 sm_100a"; }

int fepe_set_dispatch(int which, int value) {
    if (which < 0 || which >= FEPE_DISPATCH_COUNT || value < 0 || value > 3) return FEPE_E_BADARG;
    return fepe::g_dispatch[which].exchange(value);
}

int fepe_max_correspondences(void) {
    fepe::DeviceInfo& d = fepe::device_info();
    if (d.ok != 1) return FEPE_E_NODEVICE;
    const int fixed = 2 * fepe::kMaxStages * 8 + (fepe::kMaxWarps - 1) * fepe::kScratchDoubles * 8 + 256;
    return ((d.smem_optin - fixed) / 2 / 128) * 128 / 20;
}

// `pose` != nullptr: also run the pose / loss head of every pair (L = 1); fused into the latency kernel when the
// batch takes that path, otherwise as a second launch (fepe_pose_fwd) on the same stream.
static int fit_fwd_impl(const float* matches, const float* weights, int B, int N, float ax, float bx, float ay,
                        float by, float clamp_at, float* F_out, float* resid, float* epi, double* saved,
                        const fepe::PoseParams* pose, void* stream) {
    if (B == 0) return 0;
    if (!matches || !weights || !F_out || !resid || B < 0 || N <= 0) return FEPE_E_BADARG;
    if (reinterpret_cast<uintptr_t>(matches) & 15u) return FEPE_E_BADARG;
    fepe::DeviceInfo& d = fepe::device_info();
    if (d.ok != 1) return FEPE_E_NODEVICE;
    fepe::FitParams p{};
    if (!fepe::make_ring(N, 20, d.smem_optin, p.ring)) return FEPE_E_TOOLARGE;
    p.matches = matches; p.weights = weights; p.B = B; p.N = N;
    p.ax = ax; p.bx = bx; p.ay = ay; p.by = by; p.clamp_at = clamp_at;
    p.F_out = F_out; p.resid = resid; p.epi = epi; p.saved = saved;
    cudaError_t e = cudaSuccess;
    if (!d.fwd_configured) {   // once per device, outside any stream capture that may follow
        e = cudaFuncSetAttribute(fepe::fepe_fit_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 d.smem_optin);
        if (e != cudaSuccess) return static_cast<int>(e);
        d.fwd_configured = 1;
    }
    // Small batches are bound by the latency of one pair: spread each pair over a CTA of 4 warps.
    const int small_bytes = ((N * 20 + 127) / 128) * 128;
    const int force = fepe::dispatch_get(FEPE_DISPATCH_FIT);   // 0 = by size (default); tests force each path
    // one wave of the latency kernel: 6 CTAs per SM while a pair's stage is <= 28 KB (N <= 1400), else 2 (measured
    // crossovers against the ring kernel, profiles/r1_kernel_crossover.txt)
    const int small_ctas_per_sm = (small_bytes <= 28 * 1024) ? 6 : 2;
    bool use_small = (B <= small_ctas_per_sm * d.sms) && (small_bytes <= 56 * 1024);
    if (force != 0) use_small = (force == 1) && (small_bytes <= 56 * 1024);
    // Batches that fill the machine several times over go through the split pipeline (fepe_fit_split.cu).
    bool use_split = !use_small && (B >= fepe::kSplitMinPairsPerSM * d.sms) && fepe::split_path_supported(p, d);
    if (force != 0) use_split = (force == 3) && fepe::split_path_supported(p, d);
    if (use_small && pose != nullptr) {
        if (!d.small_pose_configured) {
            e = cudaFuncSetAttribute(fepe::fepe_fit_fwd_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     56 * 1024);
            if (e != cudaSuccess) return static_cast<int>(e);
            d.small_pose_configured = 1;
        }
        fepe::fepe_fit_fwd_small_kernel<true><<<B, fepe::kSmallThreads, small_bytes, static_cast<cudaStream_t>(stream)>>>(p, *pose);
        return static_cast<int>(cudaGetLastError());
    }
    if (pose != nullptr) {      // two launches: any forward path, then the stand-alone head on its F
        const int st = fit_fwd_impl(matches, weights, B, N, ax, bx, ay, by, clamp_at, F_out, resid, epi, saved, nullptr,
                                    stream);
        if (st != 0) return st;
        // the predecessor on the stream is our own fit kernel (writes F / residual / epi only): the head may start early
        return fepe::launch_pose_fwd(*pose, static_cast<cudaStream_t>(stream), /*pdl=*/true);
    }
    if (use_split) return fepe::launch_split(p, d, static_cast<cudaStream_t>(stream));
    if (use_small) {
        if (!d.small_configured) {
            e = cudaFuncSetAttribute(fepe::fepe_fit_fwd_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     56 * 1024);
            if (e != cudaSuccess) return static_cast<int>(e);
            d.small_configured = 1;
        }
        fepe::fepe_fit_fwd_small_kernel<false><<<B, fepe::kSmallThreads, small_bytes, static_cast<cudaStream_t>(stream)>>>(p, fepe::PoseParams{});
        return static_cast<int>(cudaGetLastError());
    }
    const int grid = B < d.sms ? B : d.sms;
    fepe::fepe_fit_fwd_kernel<<<grid, fepe::kThreads, p.ring.total_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    e = cudaGetLastError();
    return static_cast<int>(e);
}

int fepe_fit_fwd(const float* matches, const float* weights, int B, int N, float ax, float bx, float ay,
                 float by, float clamp_at, float* F_out, float* resid, float* epi, double* saved,
                 void* stream) {
    return fit_fwd_impl(matches, weights, B, N, ax, bx, ay, by, clamp_at, F_out, resid, epi, saved, nullptr, stream);
}

int fepe_fit_pose_fwd(const float* matches, const float* weights, int B, int N, float ax, float bx, float ay,
                      float by, float clamp_at, float* F_out, float* resid, float* epi, double* saved,
                      const float* K, const float* q_gt, const float* t_gt, const float* Rt_scene,
                      const float* virt1, const float* virt2, int V, float virt_clamp_at, float* pose_out,
                      void* stream) {
    if (B == 0) return 0;
    if (!K || !q_gt || !t_gt || !pose_out || V < 0) return FEPE_E_BADARG;
    if ((virt1 == nullptr) != (virt2 == nullptr)) return FEPE_E_BADARG;
    const fepe::PoseParams pose{F_out, K, q_gt, t_gt, Rt_scene, virt1, virt2, 1, B, V, ax, bx, ay, by, virt_clamp_at,
                                pose_out};
    return fit_fwd_impl(matches, weights, B, N, ax, bx, ay, by, clamp_at, F_out, resid, epi, saved, &pose, stream);
}

}  // extern "C"
