// The four streaming passes of the weighted 8-point fit over one pair's correspondences (shared by the
// fused kernels in fepe_fit.cu and the split pipeline in fepe_fit_split.cu).
#pragma once

#include "fepe_fit.cuh"

namespace fepe {

// ------------------------------------------------------------------------------------------------
// The four streaming passes over one pair's correspondences, resident in shared memory.  `start` /
// `stride` select the correspondences this thread owns: (lane, 32) when one warp owns the pair (ring
// kernel), (threadIdx.x, blockDim.x) when a whole CTA does (latency kernel).
// ------------------------------------------------------------------------------------------------

// pass 1: coordinate sums (4-way ILP: the loop is an LDS -> FADD latency chain otherwise)
__device__ __forceinline__ void pass_sums(const float4* __restrict__ sp, int N, int start, int stride,
                                          float (&out)[4]) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f},
          d[4] = {0.f, 0.f, 0.f, 0.f};
    int i = start;
    for (; i + 3 * stride < N; i += 4 * stride) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 q = sp[i + stride * u];
            a[u] += q.x; b[u] += q.y; c[u] += q.z; d[u] += q.w;
        }
    }
    for (; i < N; i += stride) {
        const float4 q = sp[i];
        a[0] += q.x; b[0] += q.y; c[0] += q.z; d[0] += q.w;
    }
    out[0] = (a[0] + a[1]) + (a[2] + a[3]);
    out[1] = (b[0] + b[1]) + (b[2] + b[3]);
    out[2] = (c[0] + c[1]) + (c[2] + c[3]);
    out[3] = (d[0] + d[1]) + (d[2] + d[3]);
}

// pass 2: summed distance to the centroid in primed coordinates, both images
__device__ __forceinline__ void pass_dist(const float4* __restrict__ sp, int N, int start, int stride, float ax,
                                          float ay, const PairNorm& h, float (&out)[2]) {
    float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
    const float o1x = -ax * h.m1x, o1y = -ay * h.m1y, o2x = -ax * h.m2x, o2y = -ay * h.m2y;
    int i = start;
    for (; i + 3 * stride < N; i += 4 * stride) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 q = sp[i + stride * u];
            const float u1 = fmaf(ax, q.x, o1x), v1 = fmaf(ay, q.y, o1y);
            const float u2 = fmaf(ax, q.z, o2x), v2 = fmaf(ay, q.w, o2y);
            d1[u] += approx_sqrt(fmaf(u1, u1, v1 * v1));
            d2[u] += approx_sqrt(fmaf(u2, u2, v2 * v2));
        }
    }
    for (; i < N; i += stride) {
        const float4 q = sp[i];
        const float u1 = fmaf(ax, q.x, o1x), v1 = fmaf(ay, q.y, o1y);
        const float u2 = fmaf(ax, q.z, o2x), v2 = fmaf(ay, q.w, o2y);
        d1[0] += approx_sqrt(fmaf(u1, u1, v1 * v1));
        d2[0] += approx_sqrt(fmaf(u2, u2, v2 * v2));
    }
    out[0] = (d1[0] + d1[1]) + (d1[2] + d1[3]);
    out[1] = (d2[0] + d2[1]) + (d2[2] + d2[3]);
}

// Register-tile variants of passes 1+2 for threads that own at most TILE correspondences: one round of
// shared-memory loads serves both passes, and every chain (sum, sqrt) has TILE-way instruction-level
// parallelism -- these passes are latency-bound, not throughput-bound.
// LAST_ONLY: the caller guarantees (TILE - 1) * stride <= N, i.e. only the last element of a thread can lie beyond N
// (N = 1000 with 64 threads x 16: elements 0..14 reach 959) -- no bounds test, select or predicate for the others.
template <int TILE, bool LAST_ONLY = false>
__device__ __forceinline__ void tile_load(const float4* __restrict__ sp, int N, int start, int stride,
                                          float4 (&q)[TILE]) {
#pragma unroll
    for (int u = 0; u < TILE; ++u) {
        const int i = start + u * stride;
        if (LAST_ONLY && u < TILE - 1) q[u] = sp[i];
        else q[u] = (i < N) ? sp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int TILE>
__device__ __forceinline__ void tile_sums(const float4 (&q)[TILE], float (&out)[4]) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, c[4] = {0.f, 0.f, 0.f, 0.f},
          d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < TILE; ++u) { a[u & 3] += q[u].x; b[u & 3] += q[u].y; c[u & 3] += q[u].z; d[u & 3] += q[u].w; }
    out[0] = (a[0] + a[1]) + (a[2] + a[3]);
    out[1] = (b[0] + b[1]) + (b[2] + b[3]);
    out[2] = (c[0] + c[1]) + (c[2] + c[3]);
    out[3] = (d[0] + d[1]) + (d[2] + d[3]);
}
template <int TILE, bool LAST_ONLY = false>
__device__ __forceinline__ void tile_dist(const float4 (&q)[TILE], int N, int start, int stride, float ax, float ay,
                                          const PairNorm& h, float (&out)[2]) {
    float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
    const float o1x = -ax * h.m1x, o1y = -ay * h.m1y, o2x = -ax * h.m2x, o2y = -ay * h.m2y;
#pragma unroll
    for (int u = 0; u < TILE; ++u) {
        const float u1 = fmaf(ax, q[u].x, o1x), v1 = fmaf(ay, q[u].y, o1y);
        const float u2 = fmaf(ax, q[u].z, o2x), v2 = fmaf(ay, q[u].w, o2y);
        const bool live = (LAST_ONLY && u < TILE - 1) ? true : (start + u * stride < N);
        d1[u & 3] += live ? approx_sqrt(fmaf(u1, u1, v1 * v1)) : 0.f;
        d2[u & 3] += live ? approx_sqrt(fmaf(u2, u2, v2 * v2)) : 0.f;
    }
    out[0] = (d1[0] + d1[1]) + (d1[2] + d1[3]);
    out[1] = (d2[0] + d2[1]) + (d2[2] + d2[3]);
}

__device__ __forceinline__ void finish_norm(PairNorm& h, const float (&sums)[4], const float (&dist)[2], int N,
                                            float ax, float bx, float ay, float by, bool have_dist) {
    const float invN = 1.0f / static_cast<float>(N);
    if (!have_dist) {
        h.m1x = sums[0] * invN; h.m1y = sums[1] * invN; h.m2x = sums[2] * invN; h.m2y = sums[3] * invN;
        h.c1x = fmaf(ax, h.m1x, bx); h.c1y = fmaf(ay, h.m1y, by);
        h.c2x = fmaf(ax, h.m2x, bx); h.c2y = fmaf(ay, h.m2y, by);
    } else {
        h.s1 = 1.4142f / (dist[0] * invN);       // the reference's literal (DeepFNet.py:168)
        h.s2 = 1.4142f / (dist[1] * invN);
    }
}

// raw -> Hartley-normalised coordinates: x~ = k x + j
struct PairMap {
    float k1x, k1y, k2x, k2y, j1x, j1y, j2x, j2y;
};
__device__ __forceinline__ PairMap make_map(const PairNorm& h, float ax, float ay) {
    PairMap m;
    m.k1x = h.s1 * ax; m.k1y = h.s1 * ay; m.k2x = h.s2 * ax; m.k2y = h.s2 * ay;
    m.j1x = -m.k1x * h.m1x; m.j1y = -m.k1y * h.m1y; m.j2x = -m.k2x * h.m2x; m.j2y = -m.k2y * h.m2y;
    return m;
}

// pass 3: the 36 distinct entries of G = sum_i s_i (a a^T) (x) (b b^T), fp64 accumulation.
// Software pipelined by hand: the fp32 front end + the five F2F conversions of correspondence i+1 are
// independent of the 44-instruction fp64 burst of correspondence i, so they are issued around it
// (one warp per scheduler has nothing else to hide the LDS -> FFMA -> MUFU -> F2F chain behind).
struct GramTerm {
    double x1, y1, x2, y2, s;
};
__device__ __forceinline__ GramTerm gram_prepare(const float4 q, const float wi, const PairMap& m) {
    const float x1 = fmaf(m.k1x, q.x, m.j1x), y1 = fmaf(m.k1y, q.y, m.j1y);
    const float x2 = fmaf(m.k2x, q.z, m.j2x), y2 = fmaf(m.k2y, q.w, m.j2y);
    const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
    const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
    const float s = (wi * wi) * approx_rcp(na * nb);   // (w / |p|)^2, |p|^2 = |a|^2 |b|^2 >= 1 (2 ulp is ample)
    GramTerm t;
    t.x1 = x1; t.y1 = y1; t.x2 = x2; t.y2 = y2; t.s = s;
    return t;
}
__device__ __forceinline__ void gram_accumulate(const GramTerm& c, double (&acc)[36]) {
    const double b0 = c.x1 * c.x1, b1 = c.x1 * c.y1, b3 = c.y1 * c.y1;
    const double t0 = c.s * c.x2, t1 = c.s * c.y2;
    const double a[6] = {t0 * c.x2, t0 * c.y2, t0, t1 * c.y2, t1, c.s};
#pragma unroll
    for (int u = 0; u < 6; ++u) {
        acc[u * 6 + 0] = fma(a[u], b0, acc[u * 6 + 0]);
        acc[u * 6 + 1] = fma(a[u], b1, acc[u * 6 + 1]);
        acc[u * 6 + 2] = fma(a[u], c.x1, acc[u * 6 + 2]);
        acc[u * 6 + 3] = fma(a[u], b3, acc[u * 6 + 3]);
        acc[u * 6 + 4] = fma(a[u], c.y1, acc[u * 6 + 4]);
        acc[u * 6 + 5] += a[u];
    }
}
// Two correspondences per trip (ping-pong A/B, no register moves); the loads + fp32 front end of the next
// correspondence are issued before the fp64 burst of the current one.
__device__ __forceinline__ void pass_gram(const float4* __restrict__ sp, const float* __restrict__ sw, int N,
                                          int start, int stride, const PairMap& m, double (&acc)[36]) {
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0.0;
    if (start >= N) return;
    int i = start;
    GramTerm A = gram_prepare(sp[i], sw[i], m);
#pragma unroll 1
    for (;;) {
        const int i1 = i + stride;
        if (i1 >= N) { gram_accumulate(A, acc); break; }
        const GramTerm B = gram_prepare(sp[i1], sw[i1], m);
        gram_accumulate(A, acc);
        const int i2 = i1 + stride;
        if (i2 >= N) { gram_accumulate(B, acc); break; }
        A = gram_prepare(sp[i2], sw[i2], m);
        gram_accumulate(B, acc);
        i = i2;
    }
}

// pass 4: residual r_i = w_i p^_i . f and the clamped symmetric epipolar distance with out = T2^T F_ T1
__device__ __forceinline__ float resid_one(const float4 q, const float wi, const PairMap& m, const float (&ff)[9]) {
    const float x1 = fmaf(m.k1x, q.x, m.j1x), y1 = fmaf(m.k1y, q.y, m.j1y);
    const float x2 = fmaf(m.k2x, q.z, m.j2x), y2 = fmaf(m.k2y, q.w, m.j2y);
    const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
    const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
    const float r0 = fmaf(ff[0], x1, fmaf(ff[1], y1, ff[2]));
    const float r1 = fmaf(ff[3], x1, fmaf(ff[4], y1, ff[5]));
    const float r2 = fmaf(ff[6], x1, fmaf(ff[7], y1, ff[8]));
    const float dot = fmaf(x2, r0, fmaf(y2, r1, r2));
    return wi * dot * rsqrtf(na * nb);
}
__device__ __forceinline__ float epi_one(const float4 q, const float (&Fo)[9], float ax, float bx, float ay, float by,
                                         float clamp_at) {
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
    return fminf(dist, clamp_at);
}
__device__ __forceinline__ void pass_resid(const float4* __restrict__ sp, const float* __restrict__ sw, int N,
                                           int start, int stride, const PairMap& m, const float (&ff)[9],
                                           const float (&Fo)[9], float ax, float bx, float ay, float by,
                                           float clamp_at, float* __restrict__ r_out, float* __restrict__ e_out) {
#pragma unroll 4
    for (int i = start; i < N; i += stride) {
        const float4 q = sp[i];
        const float wi = sw[i];
        __stcs(r_out + i, resid_one(q, wi, m, ff));    // streaming store: do not displace L1 lines
        if (e_out != nullptr) __stcs(e_out + i, epi_one(q, Fo, ax, bx, ay, by, clamp_at));
    }
}

}  // namespace fepe
