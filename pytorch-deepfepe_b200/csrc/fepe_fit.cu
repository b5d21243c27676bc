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
#include <stdint.h>

#include "fepe_fit.cuh"

namespace fepe {

__device__ __forceinline__ PairNorm hartley_passes(const float4* __restrict__ sp, int N, int lane,
                                                   float ax, float bx, float ay, float by) {
    PairNorm h;
    // four independent accumulation chains per lane: the loop is latency bound (LDS -> FADD) otherwise
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f},
          d[4] = {0.f, 0.f, 0.f, 0.f};
    int i = lane;
    for (; i + 96 < N; i += 128) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 q = sp[i + 32 * u];
            a[u] += q.x; b[u] += q.y; c[u] += q.z; d[u] += q.w;
        }
    }
    for (; i < N; i += 32) {
        const float4 q = sp[i];
        a[0] += q.x; b[0] += q.y; c[0] += q.z; d[0] += q.w;
    }
    const float invN = 1.0f / static_cast<float>(N);
    h.m1x = warp_sum((a[0] + a[1]) + (a[2] + a[3])) * invN;
    h.m1y = warp_sum((b[0] + b[1]) + (b[2] + b[3])) * invN;
    h.m2x = warp_sum((c[0] + c[1]) + (c[2] + c[3])) * invN;
    h.m2y = warp_sum((d[0] + d[1]) + (d[2] + d[3])) * invN;
    float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
    const float o1x = -ax * h.m1x, o1y = -ay * h.m1y, o2x = -ax * h.m2x, o2y = -ay * h.m2y;
    i = lane;
    for (; i + 96 < N; i += 128) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 q = sp[i + 32 * u];
            const float u1 = fmaf(ax, q.x, o1x), v1 = fmaf(ay, q.y, o1y);
            const float u2 = fmaf(ax, q.z, o2x), v2 = fmaf(ay, q.w, o2y);
            d1[u] += approx_sqrt(fmaf(u1, u1, v1 * v1));
            d2[u] += approx_sqrt(fmaf(u2, u2, v2 * v2));
        }
    }
    for (; i < N; i += 32) {
        const float4 q = sp[i];
        const float u1 = fmaf(ax, q.x, o1x), v1 = fmaf(ay, q.y, o1y);
        const float u2 = fmaf(ax, q.z, o2x), v2 = fmaf(ay, q.w, o2y);
        d1[0] += approx_sqrt(fmaf(u1, u1, v1 * v1));
        d2[0] += approx_sqrt(fmaf(u2, u2, v2 * v2));
    }
    const float md1 = warp_sum((d1[0] + d1[1]) + (d1[2] + d1[3])) * invN;
    const float md2 = warp_sum((d2[0] + d2[1]) + (d2[2] + d2[3])) * invN;
    h.s1 = 1.4142f / md1;
    h.s2 = 1.4142f / md2;
    h.c1x = fmaf(ax, h.m1x, bx); h.c1y = fmaf(ay, h.m1y, by);
    h.c2x = fmaf(ax, h.m2x, bx); h.c2y = fmaf(ay, h.m2y, by);
    return h;
}

struct PairSolution {
    double f[9];      // unit eigenvector of the smallest eigenvalue (= vec of the normalised F before rank 2)
    double lambda;
    double S3[3];     // v3: right singular vector of reshape(f) for its smallest singular value
    double sigma3;    // that singular value (what the rank-2 projection removed)
    float Fo[9];      // T2^T F_ T1, fp32
    int iters;
    long long cyc_eig;
};

__device__ __noinline__ void solve_pair(const double* __restrict__ gram, const PairNorm& h, PairSolution& sol) {
    double f[9], lambda;
    const long long t0 = clock64();
    sol.iters = eig9_smallest(gram, f, lambda);
    sol.cyc_eig = clock64() - t0;
    double F2[9], v3[3], sigma3;
    rank2_project(f, F2, v3, sigma3);
    denormalise_F(F2, h, sol.Fo);
#pragma unroll
    for (int i = 0; i < 9; ++i) sol.f[i] = f[i];
    sol.lambda = lambda;
    sol.S3[0] = v3[0]; sol.S3[1] = v3[1]; sol.S3[2] = v3[2];
    sol.sigma3 = sigma3;
}

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
    __syncthreads();

    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;

    if (warp == C) {
        // ---------------- producer warp ----------------
        RingSources src{{p.matches, p.weights, nullptr, nullptr}, 2};
        ring_producer(smem, p.ring, full, empty, src, N, n_local, lane);
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
        const unsigned char* sb = smem + static_cast<size_t>(stage) * p.ring.stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(sb);
        const float* sw = reinterpret_cast<const float*>(sb + pts_bytes);
        const long long tc0 = clock64();
        mbar_wait(&full[stage], phase);
        const long long tc1 = clock64();

        // ---- passes 1+2: Hartley transforms of both images (Fit.normalize with unit weights) ----
        const PairNorm h = hartley_passes(sp, N, lane, ax, bx, ay, by);
        // raw -> Hartley-normalised:  x~ = k * (x - m)
        const float k1x = h.s1 * ax, k1y = h.s1 * ay, k2x = h.s2 * ax, k2y = h.s2 * ay;
        const float j1x = -k1x * h.m1x, j1y = -k1y * h.m1y, j2x = -k2x * h.m2x, j2y = -k2y * h.m2y;

        const long long tc2 = clock64();
        // ---- pass 3: the 36 distinct entries of G = sum_i s_i (a a^T) (x) (b b^T), fp64 ----
        double acc[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) acc[i] = 0.0;
#pragma unroll 1
        for (int i = lane; i < N; i += 32) {
            const float4 q = sp[i];
            const float wi = sw[i];
            const float x1 = fmaf(k1x, q.x, j1x), y1 = fmaf(k1y, q.y, j1y);
            const float x2 = fmaf(k2x, q.z, j2x), y2 = fmaf(k2y, q.w, j2y);
            const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
            const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
            const float s = __fdividef(wi * wi, na * nb);   // (w / |p|)^2, |p|^2 = |a|^2 |b|^2 (2 ulp is ample)
            const double dx1 = x1, dy1 = y1, dx2 = x2, dy2 = y2, ds = s;
            const double b0 = dx1 * dx1, b1 = dx1 * dy1, b3 = dy1 * dy1;
            const double t0 = ds * dx2, t1 = ds * dy2;
            const double a[6] = {t0 * dx2, t0 * dy2, t0, t1 * dy2, t1, ds};
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                acc[u * 6 + 0] = fma(a[u], b0, acc[u * 6 + 0]);
                acc[u * 6 + 1] = fma(a[u], b1, acc[u * 6 + 1]);
                acc[u * 6 + 2] = fma(a[u], dx1, acc[u * 6 + 2]);
                acc[u * 6 + 3] = fma(a[u], b3, acc[u * 6 + 3]);
                acc[u * 6 + 4] = fma(a[u], dy1, acc[u * 6 + 4]);
                acc[u * 6 + 5] += a[u];
            }
        }
        const long long tc3 = clock64();
        {
            int base = 0, cnt = 36;
            ReduceScatter<36, 16>::run(acc, lane, base, cnt);
            if (cnt > 0) gram[base] = acc[0];
            if (cnt > 1) gram[base + 1] = acc[1];
        }
        __syncwarp();
        const long long tc3b = clock64();

        // ---- smallest eigenvector, rank-2 projection, de-normalisation (every lane redundantly;
        //      the inputs are warp-uniform).  Kept out of line so that its fp64 working set does not
        //      inflate the register allocation of the streaming loops.
        PairSolution sol;
        solve_pair(gram, h, sol);
        const float* Fo = sol.Fo;
        const long long tc4 = clock64();
        if (lane < 9) p.F_out[pair * 9 + lane] = Fo[lane];
        if (p.saved != nullptr) {
            double* sv = p.saved + pair * FEPE_SAVED_DOUBLES;
            if (lane == 0) {
                sv[0] = h.m1x; sv[1] = h.m1y; sv[2] = h.s1; sv[3] = h.m2x; sv[4] = h.m2y; sv[5] = h.s2;
#pragma unroll
                for (int i = 0; i < 9; ++i) sv[6 + i] = sol.f[i];
                sv[15] = sol.lambda;
                sv[52] = static_cast<double>(sol.iters);
                sv[53] = sol.S3[0]; sv[54] = sol.S3[1]; sv[55] = sol.S3[2];
                sv[63] = sol.sigma3;
            }
            for (int i = lane; i < 36; i += 32) sv[16 + i] = gram[i];
        }

        // ---- pass 4: residual r_i = w_i p_hat_i . f and the clamped epipolar distance ----
        float ff[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) ff[i] = static_cast<float>(sol.f[i]);
        float* __restrict__ r_out = p.resid + pair * static_cast<size_t>(N);
        float* __restrict__ e_out = (p.epi != nullptr) ? p.epi + pair * static_cast<size_t>(N) : nullptr;
        const float clamp_at = p.clamp_at;
#pragma unroll 4
        for (int i = lane; i < N; i += 32) {
            const float4 q = sp[i];
            const float wi = sw[i];
            const float x1 = fmaf(k1x, q.x, j1x), y1 = fmaf(k1y, q.y, j1y);
            const float x2 = fmaf(k2x, q.z, j2x), y2 = fmaf(k2y, q.w, j2y);
            const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
            const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
            const float r0 = fmaf(ff[0], x1, fmaf(ff[1], y1, ff[2]));
            const float r1 = fmaf(ff[3], x1, fmaf(ff[4], y1, ff[5]));
            const float r2 = fmaf(ff[6], x1, fmaf(ff[7], y1, ff[8]));
            const float dot = fmaf(x2, r0, fmaf(y2, r1, r2));
            r_out[i] = wi * dot * rsqrtf(na * nb);
            if (e_out != nullptr) {
                const float u1 = fmaf(ax, q.x, bx), v1 = fmaf(ay, q.y, by);
                const float u2 = fmaf(ax, q.z, bx), v2 = fmaf(ay, q.w, by);
                // l1 = x2^T F (line in image 1), l2 = F x1 (line in image 2), dd = x2^T F x1
                const float l10 = fmaf(u2, Fo[0], fmaf(v2, Fo[3], Fo[6]));
                const float l11 = fmaf(u2, Fo[1], fmaf(v2, Fo[4], Fo[7]));
                const float l12 = fmaf(u2, Fo[2], fmaf(v2, Fo[5], Fo[8]));
                const float l20 = fmaf(Fo[0], u1, fmaf(Fo[1], v1, Fo[2]));
                const float l21 = fmaf(Fo[3], u1, fmaf(Fo[4], v1, Fo[5]));
                const float dd = fmaf(l10, u1, fmaf(l11, v1, l12));
                const float n1 = approx_sqrt(fmaf(l10, l10, l11 * l11)) + 1e-6f;
                const float n2 = approx_sqrt(fmaf(l20, l20, l21 * l21)) + 1e-6f;
                const float dist = fabsf(dd) * (approx_rcp(n1) + approx_rcp(n2));
                e_out[i] = fminf(dist, clamp_at);
            }
        }
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
                sv[62] = static_cast<double>(sol.cyc_eig); // of which: eigen iteration
            }
        }
    }
}

}  // namespace fepe

extern "C" {

const char* fepe_version(void) { return "fepe_b200 0.1 sm_100a"; }

int fepe_max_correspondences(void) {
    fepe::DeviceInfo& d = fepe::device_info();
    if (d.ok != 1) return FEPE_E_NODEVICE;
    const int fixed = 2 * fepe::kMaxStages * 8 + (fepe::kMaxWarps - 1) * fepe::kScratchDoubles * 8 + 256;
    return ((d.smem_optin - fixed) / 2 / 128) * 128 / 20;
}

int fepe_fit_fwd(const float* matches, const float* weights, int B, int N, float ax, float bx, float ay,
                 float by, float clamp_at, float* F_out, float* resid, float* epi, double* saved,
                 void* stream) {
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
    const int grid = B < d.sms ? B : d.sms;
    fepe::fepe_fit_fwd_kernel<<<grid, fepe::kThreads, p.ring.total_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    e = cudaGetLastError();
    return static_cast<int>(e);
}

}  // extern "C"
