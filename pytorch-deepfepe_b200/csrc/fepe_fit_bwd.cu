// Analytic backward of the fused weighted 8-point forward w.r.t. the correspondence weights.
//
// Replaces what autograd does in the reference when loss.backward() (Train_model_pipeline.py:595)
// walks DeepFNet.py:198-256 and utils_F.py:400-413: SvdBackward of the N x 9 SVD (materialising U
// [N,9] per pair per layer), SvdBackward of the 3x3 SVD, and ~40 elementwise / bmm nodes.
//
// With x_i = w_i p^_i, G = sum x_i x_i^T, (lambda, f) its smallest eigenpair, F2 = rank2(reshape f),
// out = T2^T F2 T1, r_i = x_i . f, e_i = clamp(epi_i(out)):
//   outbar = gF + sum_i ebar_i de_i/dout                                   (pass 1, 9 sums)
//   F2bar  = T2 outbar T1^T ;  F0bar = rank2 adjoint (fepe_math.cuh)
//   fbar   = vec(F0bar) + sum_i rbar_i x_i                                 (pass 1, 9 more sums)
//   z      = (G - lambda I)^+ fbar                                         (one 9x9 solve, G saved by forward)
//   wbar_i = -2 w_i (p^_i . z)(p^_i . f) + rbar_i (p^_i . f)               (pass 2)
// The Hartley transforms depend on the coordinates only (Fit.normalize is called with unit weights,
// DeepFNet.py:194-199), so they carry no weight gradient.  Same pair ring as the forward: one warp per
// pair, the pair's coordinates, weights and the two upstream rows staged once (28 B / correspondence).
//
// Coordinate gradient (`gmatches` != null; the reference needs it with if_learn_offsets, DeepFNet.py:373,489-505,
// or a trainable keypoint front-end): pass 2 also forms the adjoint of every constraint row w.r.t. the
// Hartley-normalised coordinates and of the epipolar distance w.r.t. the primed ones (fepe_fit_adjoint.cuh), parks
// the four partial gradients of correspondence i in the four dead row slots of the stage (weight, rbar, ebar and
// one spare row: 32 B / correspondence staged) and accumulates the ten sums the adjoint of Fit.normalize needs;
// pass 3 adds the mean / mean-distance terms and writes [N,4] once.  No read-modify-write of global memory.
#include "fepe_dispatch.cuh"
#include "fepe_fit.cuh"
#include "fepe_fit_adjoint.cuh"

namespace fepe {

template <bool COORDS>
__global__ void __launch_bounds__(kThreads, 1) fepe_fit_bwd_kernel(const FitParams p) {
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

    if (warp == C) {
        ring_producer(smem, p.ring, full, empty, RingSources{{p.matches, p.weights, p.gresid, p.gepi}, 4}, N, n_local,
                      lane);
        return;
    }
    if (warp > C) return;

    double* gram = reinterpret_cast<double*>(smem + p.ring.scratch_off) + warp * kScratchDoubles;
    const float ax = p.ax, bx = p.bx, ay = p.ay, by = p.by;
    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;
    const uint32_t row_bytes = static_cast<uint32_t>(N) * 4u;
    const bool has_gr = p.gresid != nullptr, has_ge = p.gepi != nullptr;

    for (int j = warp; j < n_local; j += C) {
        const int stage = j % S;
        const uint32_t phase = static_cast<uint32_t>(j / S) & 1u;
        const size_t pair = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        unsigned char* sb = smem + static_cast<size_t>(stage) * p.ring.stage_bytes;
        const float4* sp = reinterpret_cast<const float4*>(sb);
        float* sw = reinterpret_cast<float*>(sb + pts_bytes);
        float* sgr = reinterpret_cast<float*>(sb + pts_bytes + row_bytes);
        float* sge = reinterpret_cast<float*>(sb + pts_bytes + 2 * row_bytes);
        float* sx4 = reinterpret_cast<float*>(sb + pts_bytes + 3 * row_bytes);   // spare row (coordinate gradient only)

        // ---- per-pair state saved by the forward (warp-uniform) ----
        const double* sv = p.saved + pair * FEPE_SAVED_DOUBLES;
        PairNorm h;
        h.m1x = static_cast<float>(sv[0]); h.m1y = static_cast<float>(sv[1]); h.s1 = static_cast<float>(sv[2]);
        h.m2x = static_cast<float>(sv[3]); h.m2y = static_cast<float>(sv[4]); h.s2 = static_cast<float>(sv[5]);
        h.c1x = fmaf(ax, h.m1x, bx); h.c1y = fmaf(ay, h.m1y, by);
        h.c2x = fmaf(ax, h.m2x, bx); h.c2y = fmaf(ay, h.m2y, by);
        double f[9], v3[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = sv[6 + i];
        const double lambda = sv[15];
        v3[0] = sv[53]; v3[1] = sv[54]; v3[2] = sv[55];
        for (int i = lane; i < 36; i += 32) gram[i] = sv[16 + i];
        double F2[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double wv = f[3 * r] * v3[0] + f[3 * r + 1] * v3[1] + f[3 * r + 2] * v3[2];
            F2[3 * r] = f[3 * r] - wv * v3[0]; F2[3 * r + 1] = f[3 * r + 1] - wv * v3[1]; F2[3 * r + 2] = f[3 * r + 2] - wv * v3[2];
        }
        float Fo[9], ff[9];
        denormalise_F(F2, h, Fo);
#pragma unroll
        for (int i = 0; i < 9; ++i) ff[i] = static_cast<float>(f[i]);
        const float k1x = h.s1 * ax, k1y = h.s1 * ay, k2x = h.s2 * ax, k2y = h.s2 * ay;
        const float j1x = -k1x * h.m1x, j1y = -k1y * h.m1y, j2x = -k2x * h.m2x, j2y = -k2y * h.m2y;
        const float clamp_at = p.clamp_at;

        mbar_wait(&full[stage], phase);
        ring_fill_ragged(sb, RingSources{{p.matches, p.weights, p.gresid, p.gepi}, 4}, pair, N, lane);

        // ---- pass 1: hsum = sum rbar_i x_i  and  ge = sum ebar_i d epi_i / d out ----
        float hs[9], ge[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { hs[i] = 0.f; ge[i] = 0.f; }
#pragma unroll 2
        for (int i = lane; i < N; i += 32) {
            const float4 q = sp[i];
            if (has_gr) {
                const float x1 = fmaf(k1x, q.x, j1x), y1 = fmaf(k1y, q.y, j1y);
                const float x2 = fmaf(k2x, q.z, j2x), y2 = fmaf(k2y, q.w, j2y);
                const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
                const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
                const float c = sgr[i] * sw[i] * rsqrtf(na * nb);
                const float c0 = c * x2, c1 = c * y2;
                hs[0] = fmaf(c0, x1, hs[0]); hs[1] = fmaf(c0, y1, hs[1]); hs[2] += c0;
                hs[3] = fmaf(c1, x1, hs[3]); hs[4] = fmaf(c1, y1, hs[4]); hs[5] += c1;
                hs[6] = fmaf(c, x1, hs[6]);  hs[7] = fmaf(c, y1, hs[7]);  hs[8] += c;
            }
            if (has_ge) {
                const float u1 = fmaf(ax, q.x, bx), v1 = fmaf(ay, q.y, by);
                const float u2 = fmaf(ax, q.z, bx), v2 = fmaf(ay, q.w, by);
                const float l10 = fmaf(u2, Fo[0], fmaf(v2, Fo[3], Fo[6]));
                const float l11 = fmaf(u2, Fo[1], fmaf(v2, Fo[4], Fo[7]));
                const float l12 = fmaf(u2, Fo[2], fmaf(v2, Fo[5], Fo[8]));
                const float l20 = fmaf(Fo[0], u1, fmaf(Fo[1], v1, Fo[2]));
                const float l21 = fmaf(Fo[3], u1, fmaf(Fo[4], v1, Fo[5]));
                const float dd = fmaf(l10, u1, fmaf(l11, v1, l12));
                const float m1 = sqrtf(fmaf(l10, l10, l11 * l11)), m2 = sqrtf(fmaf(l20, l20, l21 * l21));
                const float i1 = 1.0f / (m1 + 1e-6f), i2 = 1.0f / (m2 + 1e-6f);
                const float dist = fabsf(dd) * (i1 + i2);
                // torch.clamp(max=c) passes the gradient where d <= c
                const float g = (dist <= clamp_at) ? sge[i] : 0.f;
                const float sg = (dd > 0.f) ? g : ((dd < 0.f) ? -g : 0.f);
                const float S12 = sg * (i1 + i2);
                const float a1 = -g * fabsf(dd) * i1 * i1 / fmaxf(m1, 1e-30f);   // d(1/(m1+eps)) = -i1^2 dm1, dm1 = l1.dl1/m1
                const float a2 = -g * fabsf(dd) * i2 * i2 / fmaxf(m2, 1e-30f);
                // d dd / dF_jk = x2_j x1_k ; d m1 / dF_jk = l1_k x2_j / m1 (k<2) ; d m2 / dF_jk = l2_j x1_k / m2 (j<2)
                const float uk0 = fmaf(S12, u1, a1 * l10), uk1 = fmaf(S12, v1, a1 * l11), uk2 = S12;   // times x2_j
                const float vj0 = a2 * l20, vj1 = a2 * l21;                                            // times x1_k
                ge[0] += fmaf(u2, uk0, vj0 * u1); ge[1] += fmaf(u2, uk1, vj0 * v1); ge[2] += fmaf(u2, uk2, vj0);
                ge[3] += fmaf(v2, uk0, vj1 * u1); ge[4] += fmaf(v2, uk1, vj1 * v1); ge[5] += fmaf(v2, uk2, vj1);
                ge[6] += uk0;                     ge[7] += uk1;                     ge[8] += uk2;
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) { hs[i] = warp_sum(hs[i]); ge[i] = warp_sum(ge[i]); }

        // ---- small algebra (warp-uniform, every lane) ----
        double ob[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) ob[i] = static_cast<double>(ge[i]) + static_cast<double>(p.gF[pair * 9 + i]);
        double Ab[9];
        {
            const double s1 = h.s1, s2 = h.s2;
            const double t1x = -s1 * h.c1x, t1y = -s1 * h.c1y, t2x = -s2 * h.c2x, t2y = -s2 * h.c2y;
            double X[9];   // T2 * outbar
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                X[k] = s2 * ob[k] + t2x * ob[6 + k];
                X[3 + k] = s2 * ob[3 + k] + t2y * ob[6 + k];
                X[6 + k] = ob[6 + k];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {   // ... * T1^T
                Ab[3 * r] = X[3 * r] * s1 + X[3 * r + 2] * t1x;
                Ab[3 * r + 1] = X[3 * r + 1] * s1 + X[3 * r + 2] * t1y;
                Ab[3 * r + 2] = X[3 * r + 2];
            }
        }
        double fb[9], z[9];
        rank2_project_adjoint(f, v3, Ab, fb);
#pragma unroll
        for (int i = 0; i < 9; ++i) fb[i] += static_cast<double>(hs[i]);
        __syncwarp();
        eig9_pinv_apply(gram, f, lambda, fb, z);
        float zf[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) zf[i] = static_cast<float>(z[i]);

        // ---- pass 2: wbar_i (and the partial coordinate gradient) ----
        float* __restrict__ gw_out = p.gweights + pair * static_cast<size_t>(N);
        if constexpr (!COORDS) {
#pragma unroll 2
            for (int i = lane; i < N; i += 32) {
                const float4 q = sp[i];
                const float x1 = fmaf(k1x, q.x, j1x), y1 = fmaf(k1y, q.y, j1y);
                const float x2 = fmaf(k2x, q.z, j2x), y2 = fmaf(k2y, q.w, j2y);
                const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
                const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
                const float inv = rsqrtf(na * nb);
                const float f0 = fmaf(ff[0], x1, fmaf(ff[1], y1, ff[2]));
                const float f1 = fmaf(ff[3], x1, fmaf(ff[4], y1, ff[5]));
                const float f2 = fmaf(ff[6], x1, fmaf(ff[7], y1, ff[8]));
                const float pf = fmaf(x2, f0, fmaf(y2, f1, f2)) * inv;
                const float z0 = fmaf(zf[0], x1, fmaf(zf[1], y1, zf[2]));
                const float z1 = fmaf(zf[3], x1, fmaf(zf[4], y1, zf[5]));
                const float z2 = fmaf(zf[6], x1, fmaf(zf[7], y1, zf[8]));
                const float pz = fmaf(x2, z0, fmaf(y2, z1, z2)) * inv;
                const float gr = has_gr ? sgr[i] : 0.f;
                gw_out[i] = fmaf(-2.0f * sw[i] * pz, pf, gr * pf);
            }
        } else {
            // sums for the adjoint of Fit.normalize: [0..1] sum x~bar, [2..3] sum y~bar, [4..5] sum x~bar.(u-c),
            // [6..7] sum (u-cx)/d, [8..9] sum (v-cy)/d   (even: image 1, odd: image 2)
            float ns[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) ns[k] = 0.f;
            const float s1f = h.s1, s2f = h.s2;
            for (int i = lane; i < N; i += 32) {
                const float4 q = sp[i];
                const float u1 = fmaf(ax, q.x, bx), v1 = fmaf(ay, q.y, by);
                const float u2 = fmaf(ax, q.z, bx), v2 = fmaf(ay, q.w, by);
                const float du1 = u1 - h.c1x, dv1 = v1 - h.c1y, du2 = u2 - h.c2x, dv2 = v2 - h.c2y;
                const float gr = has_gr ? sgr[i] : 0.f;
                float wb, xb[4];
                row_adjoint<float>(s1f * du1, s1f * dv1, s2f * du2, s2f * dv2, sw[i], gr, ff, zf, wb, xb);
                gw_out[i] = wb;
                float cb[4] = {0.f, 0.f, 0.f, 0.f};
                if (has_ge) {
                    float ge_unused[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    epi_adjoint<float>(u1, v1, u2, v2, Fo, clamp_at, sge[i], ge_unused, cb);
                }
                ns[0] += xb[0]; ns[2] += xb[1]; ns[4] = fmaf(xb[0], du1, fmaf(xb[1], dv1, ns[4]));
                ns[1] += xb[2]; ns[3] += xb[3]; ns[5] = fmaf(xb[2], du2, fmaf(xb[3], dv2, ns[5]));
                const float q1 = fmaf(du1, du1, dv1 * dv1), q2 = fmaf(du2, du2, dv2 * dv2);
                const float id1 = (q1 > 0.f) ? rsqrtf(q1) : 0.f, id2 = (q2 > 0.f) ? rsqrtf(q2) : 0.f;
                ns[6] = fmaf(du1, id1, ns[6]); ns[8] = fmaf(dv1, id1, ns[8]);
                ns[7] = fmaf(du2, id2, ns[7]); ns[9] = fmaf(dv2, id2, ns[9]);
                // partial gradient w.r.t. the primed coordinates; same lane reads and writes index i
                sw[i] = fmaf(s1f, xb[0], cb[0]);
                sgr[i] = fmaf(s1f, xb[1], cb[1]);
                sge[i] = fmaf(s2f, xb[2], cb[2]);
                sx4[i] = fmaf(s2f, xb[3], cb[3]);
            }
#pragma unroll
            for (int k = 0; k < 10; ++k) ns[k] = warp_sum(ns[k]);
            NormAdjointSums S;
#pragma unroll
            for (int im = 0; im < 2; ++im) {
                S.sx[im] = ns[im]; S.sy[im] = ns[2 + im]; S.sd[im] = ns[4 + im];
                S.dx[im] = ns[6 + im]; S.dy[im] = ns[8 + im];
            }
            const double sc[2] = {static_cast<double>(h.s1), static_cast<double>(h.s2)};
            const double ccx[2] = {static_cast<double>(h.c1x), static_cast<double>(h.c2x)};
            const double ccy[2] = {static_cast<double>(h.c1y), static_cast<double>(h.c2y)};
            NormAdjointCoef C;
            norm_adjoint(ob, F2, sc, ccx, ccy, S, N, C);
            const float A1 = static_cast<float>(C.A[0]), A2 = static_cast<float>(C.A[1]);
            const float B1x = static_cast<float>(C.Bx[0]), B1y = static_cast<float>(C.By[0]);
            const float B2x = static_cast<float>(C.Bx[1]), B2y = static_cast<float>(C.By[1]);
            __syncwarp();
            // ---- pass 3: add the mean / mean-distance terms, chain through the affine, write [N,4] ----
            float4* __restrict__ gm_out = reinterpret_cast<float4*>(p.gmatches) + pair * static_cast<size_t>(N);
#pragma unroll 2
            for (int i = lane; i < N; i += 32) {
                const float4 q = sp[i];
                const float du1 = fmaf(ax, q.x, bx) - h.c1x, dv1 = fmaf(ay, q.y, by) - h.c1y;
                const float du2 = fmaf(ax, q.z, bx) - h.c2x, dv2 = fmaf(ay, q.w, by) - h.c2y;
                const float q1 = fmaf(du1, du1, dv1 * dv1), q2 = fmaf(du2, du2, dv2 * dv2);
                const float id1 = (q1 > 0.f) ? A1 * rsqrtf(q1) : 0.f, id2 = (q2 > 0.f) ? A2 * rsqrtf(q2) : 0.f;
                float4 g;
                g.x = ax * (sw[i] + fmaf(du1, id1, B1x));
                g.y = ay * (sgr[i] + fmaf(dv1, id1, B1y));
                g.z = ax * (sge[i] + fmaf(du2, id2, B2x));
                g.w = ay * (sx4[i] + fmaf(dv2, id2, B2y));
                gm_out[i] = g;
            }
        }
        // the coordinate path wrote into the stage through the generic proxy; the refill is an async-proxy write
        if constexpr (COORDS) fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
    }
}

// ------------------------------------------------------------------------------------------------
// Latency variant for small batches (weights gradient only): ONE CTA of 4 warps per pair instead of one warp of a
// persistent ring -- the same two passes spread over 128 threads with block reductions, the warp-uniform 9x9 algebra on
// every thread (it is a serial fp64 chain: redundancy costs nothing and saves a broadcast).  The ring kernel walks a pair
// with 32 lanes (2 x 31 trips): 33.5 us for a 256-pair launch (profiles/r2_bwd.md); this kernel needs 2 x 8 trips.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdSmallThreads = 128;

__global__ void __launch_bounds__(kBwdSmallThreads, 3) fepe_fit_bwd_small_kernel(const FitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar;
    __shared__ float red[kBwdSmallThreads / 32][18];
    __shared__ double gram[36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int N = p.N;
    const size_t pair = blockIdx.x;
    const uint32_t pts_bytes = static_cast<uint32_t>(N) * 16u;
    const uint32_t row_bytes = static_cast<uint32_t>(N) * 4u;
    const float4* sp = reinterpret_cast<const float4*>(smem);
    float* sw = reinterpret_cast<float*>(smem + pts_bytes);
    float* sgr = reinterpret_cast<float*>(smem + pts_bytes + row_bytes);
    float* sge = reinterpret_cast<float*>(smem + pts_bytes + 2 * row_bytes);
    const bool has_gr = p.gresid != nullptr, has_ge = p.gepi != nullptr;
    const float* rows[3] = {p.weights, p.gresid, p.gepi};
    float* dsts[3] = {sw, sgr, sge};

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t tx = pts_bytes;
        for (int a = 0; a < 3; ++a)
            if (ring_row_is_bulk(rows[a], pair, N)) tx += row_bytes;
        mbar_arrive_expect_tx(&full_bar, tx);
        bulk_g2s(smem, p.matches + pair * static_cast<size_t>(N) * 4, pts_bytes, &full_bar);
        for (int a = 0; a < 3; ++a)
            if (ring_row_is_bulk(rows[a], pair, N))
                bulk_g2s(dsts[a], rows[a] + pair * static_cast<size_t>(N), row_bytes, &full_bar);
    }
    for (int a = 0; a < 3; ++a) {                           // ragged rows: plain loads by the CTA itself
        if (rows[a] == nullptr || ring_row_is_bulk(rows[a], pair, N)) continue;
        const float* g = rows[a] + pair * static_cast<size_t>(N);
        for (int i = tid; i < N; i += kBwdSmallThreads) dsts[a][i] = __ldg(g + i);
    }
    // ---- per-pair state saved by the forward (uniform) -- fetched behind the bulk copy ----
    const float ax = p.ax, bx = p.bx, ay = p.ay, by = p.by;
    const double* sv = p.saved + pair * FEPE_SAVED_DOUBLES;
    PairNorm h;
    h.m1x = static_cast<float>(sv[0]); h.m1y = static_cast<float>(sv[1]); h.s1 = static_cast<float>(sv[2]);
    h.m2x = static_cast<float>(sv[3]); h.m2y = static_cast<float>(sv[4]); h.s2 = static_cast<float>(sv[5]);
    h.c1x = fmaf(ax, h.m1x, bx); h.c1y = fmaf(ay, h.m1y, by);
    h.c2x = fmaf(ax, h.m2x, bx); h.c2y = fmaf(ay, h.m2y, by);
    double f[9], v3[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = sv[6 + i];
    const double lambda = sv[15];
    v3[0] = sv[53]; v3[1] = sv[54]; v3[2] = sv[55];
    if (tid < 36) gram[tid] = sv[16 + tid];
    double F2[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double wv = f[3 * r] * v3[0] + f[3 * r + 1] * v3[1] + f[3 * r + 2] * v3[2];
        F2[3 * r] = f[3 * r] - wv * v3[0]; F2[3 * r + 1] = f[3 * r + 1] - wv * v3[1]; F2[3 * r + 2] = f[3 * r + 2] - wv * v3[2];
    }
    float Fo[9], ff[9];
    denormalise_F(F2, h, Fo);
#pragma unroll
    for (int i = 0; i < 9; ++i) ff[i] = static_cast<float>(f[i]);
    const float k1x = h.s1 * ax, k1y = h.s1 * ay, k2x = h.s2 * ax, k2y = h.s2 * ay;
    const float j1x = -k1x * h.m1x, j1y = -k1y * h.m1y, j2x = -k2x * h.m2x, j2y = -k2y * h.m2y;
    const float clamp_at = p.clamp_at;
    double gFv[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) gFv[i] = static_cast<double>(p.gF[pair * 9 + i]);

    mbar_wait(&full_bar, 0);
    __syncthreads();                                         // also orders the hand-copied rows and gram[]

    // ---- pass 1: hs = sum rbar_i x_i  and  ge = sum ebar_i d epi_i / d out ----
    float hs[9], ge[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { hs[i] = 0.f; ge[i] = 0.f; }
#pragma unroll 2
    for (int i = tid; i < N; i += kBwdSmallThreads) {
        const float4 q = sp[i];
        if (has_gr) {
            const float x1 = fmaf(k1x, q.x, j1x), y1 = fmaf(k1y, q.y, j1y);
            const float x2 = fmaf(k2x, q.z, j2x), y2 = fmaf(k2y, q.w, j2y);
            const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
            const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
            const float c = sgr[i] * sw[i] * rsqrtf(na * nb);
            const float c0 = c * x2, c1 = c * y2;
            hs[0] = fmaf(c0, x1, hs[0]); hs[1] = fmaf(c0, y1, hs[1]); hs[2] += c0;
            hs[3] = fmaf(c1, x1, hs[3]); hs[4] = fmaf(c1, y1, hs[4]); hs[5] += c1;
            hs[6] = fmaf(c, x1, hs[6]);  hs[7] = fmaf(c, y1, hs[7]);  hs[8] += c;
        }
        if (has_ge) {
            const float u1 = fmaf(ax, q.x, bx), v1 = fmaf(ay, q.y, by);
            const float u2 = fmaf(ax, q.z, bx), v2 = fmaf(ay, q.w, by);
            float cb_unused[4];
            epi_adjoint<float>(u1, v1, u2, v2, Fo, clamp_at, sge[i], ge, cb_unused);
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) { hs[i] = warp_sum(hs[i]); ge[i] = warp_sum(ge[i]); }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) { red[warp][i] = hs[i]; red[warp][9 + i] = ge[i]; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < kBwdSmallThreads / 32; ++w) { a += red[w][i]; b += red[w][9 + i]; }
        hs[i] = a; ge[i] = b;
    }

    // ---- small algebra (uniform, every thread) ----
    double ob[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) ob[i] = static_cast<double>(ge[i]) + gFv[i];
    double Ab[9];
    {
        const double s1 = h.s1, s2 = h.s2;
        const double t1x = -s1 * h.c1x, t1y = -s1 * h.c1y, t2x = -s2 * h.c2x, t2y = -s2 * h.c2y;
        double X[9];   // T2 * outbar
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            X[k] = s2 * ob[k] + t2x * ob[6 + k];
            X[3 + k] = s2 * ob[3 + k] + t2y * ob[6 + k];
            X[6 + k] = ob[6 + k];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {   // ... * T1^T
            Ab[3 * r] = X[3 * r] * s1 + X[3 * r + 2] * t1x;
            Ab[3 * r + 1] = X[3 * r + 1] * s1 + X[3 * r + 2] * t1y;
            Ab[3 * r + 2] = X[3 * r + 2];
        }
    }
    double fb[9], z[9];
    rank2_project_adjoint(f, v3, Ab, fb);
#pragma unroll
    for (int i = 0; i < 9; ++i) fb[i] += static_cast<double>(hs[i]);
    eig9_pinv_apply(gram, f, lambda, fb, z);
    float zf[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) zf[i] = static_cast<float>(z[i]);

    // ---- pass 2: wbar_i ----
    float* __restrict__ gw_out = p.gweights + pair * static_cast<size_t>(N);
#pragma unroll 2
    for (int i = tid; i < N; i += kBwdSmallThreads) {
        const float4 q = sp[i];
        const float x1 = fmaf(k1x, q.x, j1x), y1 = fmaf(k1y, q.y, j1y);
        const float x2 = fmaf(k2x, q.z, j2x), y2 = fmaf(k2y, q.w, j2y);
        const float nb = fmaf(x1, x1, fmaf(y1, y1, 1.0f));
        const float na = fmaf(x2, x2, fmaf(y2, y2, 1.0f));
        const float inv = rsqrtf(na * nb);
        const float f0 = fmaf(ff[0], x1, fmaf(ff[1], y1, ff[2]));
        const float f1 = fmaf(ff[3], x1, fmaf(ff[4], y1, ff[5]));
        const float f2 = fmaf(ff[6], x1, fmaf(ff[7], y1, ff[8]));
        const float pf = fmaf(x2, f0, fmaf(y2, f1, f2)) * inv;
        const float z0 = fmaf(zf[0], x1, fmaf(zf[1], y1, zf[2]));
        const float z1 = fmaf(zf[3], x1, fmaf(zf[4], y1, zf[5]));
        const float z2 = fmaf(zf[6], x1, fmaf(zf[7], y1, zf[8]));
        const float pz = fmaf(x2, z0, fmaf(y2, z1, z2)) * inv;
        const float gr = has_gr ? sgr[i] : 0.f;
        gw_out[i] = fmaf(-2.0f * sw[i] * pz, pf, gr * pf);
    }
}

}  // namespace fepe

extern "C" int fepe_fit_bwd(const float* matches, const float* weights, int B, int N, float ax, float bx, float ay,
                            float by, float clamp_at, const double* saved, const float* gF, const float* gresid,
                            const float* gepi, float* gweights, void* stream) {
    return fepe_fit_bwd_coords(matches, weights, B, N, ax, bx, ay, by, clamp_at, saved, gF, gresid, gepi, gweights,
                               nullptr, stream);
}

extern "C" int fepe_fit_bwd_coords(const float* matches, const float* weights, int B, int N, float ax, float bx,
                                   float ay, float by, float clamp_at, const double* saved, const float* gF,
                                   const float* gresid, const float* gepi, float* gweights, float* gmatches,
                                   void* stream) {
    if (B == 0) return 0;
    if (!matches || !weights || !saved || !gF || !gweights || B < 0 || N <= 0) return FEPE_E_BADARG;
    if (reinterpret_cast<uintptr_t>(matches) & 15u) return FEPE_E_BADARG;
    fepe::DeviceInfo& d = fepe::device_info();
    if (d.ok != 1) return FEPE_E_NODEVICE;
    fepe::FitParams p{};
    if (gmatches != nullptr && (reinterpret_cast<uintptr_t>(gmatches) & 15u)) return FEPE_E_BADARG;
    if (!fepe::make_ring(N, gmatches != nullptr ? 32 : 28, d.smem_optin, p.ring)) return FEPE_E_TOOLARGE;
    p.matches = matches; p.weights = weights; p.B = B; p.N = N;
    p.ax = ax; p.bx = bx; p.ay = ay; p.by = by; p.clamp_at = clamp_at;
    p.saved = const_cast<double*>(saved);
    p.gF = gF; p.gresid = gresid; p.gepi = gepi; p.gweights = gweights; p.gmatches = gmatches;
    cudaError_t e = cudaSuccess;
    if (!d.bwd_configured) {
        e = cudaFuncSetAttribute(fepe::fepe_fit_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 d.smem_optin);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(fepe::fepe_fit_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 d.smem_optin);
        if (e != cudaSuccess) return static_cast<int>(e);
        d.bwd_configured = 1;
    }
    // small batches, weights gradient only: the one-CTA-per-pair latency kernel (3 CTAs of 28 KB per SM and wave)
    const int small_bytes = ((N * 28 + 127) / 128) * 128;
    const int force = fepe::dispatch_get(FEPE_DISPATCH_FIT);          // tests force each path (1 = small, 2 = ring)
    bool use_small = gmatches == nullptr && small_bytes <= 64 * 1024 && B <= 3 * d.sms;
    if (force != 0) use_small = (force == 1) && gmatches == nullptr && small_bytes <= 64 * 1024;
    if (use_small) {
        static int configured[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev & 63]) {
            e = cudaFuncSetAttribute(fepe::fepe_fit_bwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return static_cast<int>(e);
            configured[dev & 63] = 1;
        }
        fepe::fepe_fit_bwd_small_kernel<<<B, fepe::kBwdSmallThreads, small_bytes, static_cast<cudaStream_t>(stream)>>>(p);
        return static_cast<int>(cudaGetLastError());
    }
    const int grid = B < d.sms ? B : d.sms;
    if (gmatches != nullptr)
        fepe::fepe_fit_bwd_kernel<true><<<grid, fepe::kThreads, p.ring.total_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    else
        fepe::fepe_fit_bwd_kernel<false><<<grid, fepe::kThreads, p.ring.total_bytes, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}
