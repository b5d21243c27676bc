// Pose / loss head: per image pair and per layer
//   E = K^T T2^T F T1 K                      (deepFEPE/train_good_utils.py:356-358)
//   E^T -> {R1,R2},{t,-t}                    (train_good_utils.py:106, dsac_tools/utils_F.py:478-498)
//   quaternions, L2 to the GT pose, min-select, angular errors   (train_good_utils.py:149-188,
//                                            dsac_tools/utils_geo.py:58-86,150-152,175-179)
//   F-loss: clamped epipolar residual of the virtual correspondences, mean over them
//                                            (train_good_utils.py:325-354)
// The reference does all of this on the HOST in a Python loop over layers x batch with .cpu()
// round trips.  Here one WARP owns one (layer, pair): every lane runs the (warp-uniform) 3x3 SVD and
// quaternion code redundantly -- it is a serial dependency chain, so this costs nothing extra and
// keeps the latency of a 256-pair launch at one SVD -- while the V virtual correspondences of the
// F-loss are spread over the lanes and summed with shuffles.  ~O(100) bytes and O(1e3) flops per
// pair: negligible next to the streaming kernel, it just has to stay on the device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "fepe_common.cuh"
#include "fepe_pose_head.cuh"

namespace fepe {

__global__ void __launch_bounds__(128) fepe_pose_fwd_kernel(const PoseParams p) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per (layer, pair)
    const int lane = threadIdx.x & 31;
    if (idx >= p.L * p.B) return;
    const int b = idx % p.B;
    float* __restrict__ o = p.out + static_cast<size_t>(idx) * FEPE_POSE_OUT_FLOATS;

    // Every global load of the item is issued here, before any arithmetic: the kernel is one dependent chain per
    // warp, so a load that sits behind the SVD costs a full memory round trip of latency.
    const float* __restrict__ gK = p.K + static_cast<size_t>(b) * 9;
    const float* __restrict__ gq = p.q_gt + static_cast<size_t>(b) * 4;
    const float* __restrict__ gt = p.t_gt + static_cast<size_t>(b) * 3;
    float Kf[9], qgf[4], tgf[3], rtf[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 9; ++i) Kf[i] = __ldg(gK + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) qgf[i] = __ldg(gq + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) tgf[i] = __ldg(gt + i);
    if (p.Rt != nullptr) {
        const float* __restrict__ rt = p.Rt + static_cast<size_t>(b) * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rtf[3 * r + c] = __ldg(rt + 4 * r + c);
        }
    }
    const bool has_virt = (p.virt1 != nullptr) && (p.V > 0);
    const float* __restrict__ v1 = has_virt ? p.virt1 + static_cast<size_t>(b) * p.V * 3 : nullptr;
    const float* __restrict__ v2 = has_virt ? p.virt2 + static_cast<size_t>(b) * p.V * 3 : nullptr;
    float va[kVirtPerLane][3], vb[kVirtPerLane][3];
#pragma unroll
    for (int u = 0; u < kVirtPerLane; ++u) {
        const int i = lane + 32 * u;
        const bool live = has_virt && i < p.V;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            va[u][c] = live ? __ldg(v1 + 3 * i + c) : 0.f;
            vb[u][c] = live ? __ldg(v2 + 3 * i + c) : 0.f;
        }
    }
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");   // F comes from the preceding kernel of the stream (PDL)
#endif
    float Ff[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Ff[i] = p.F[static_cast<size_t>(idx) * 9 + i];

    // F-loss over the virtual correspondences (fp32 like the reference); independent of the SVD chain below
    float loss = 0.f;
    if (has_virt) {
#pragma unroll
        for (int u = 0; u < kVirtPerLane; ++u) {
            if (lane + 32 * u < p.V)
                loss += virt_term(Ff, p.ax, p.bx, p.ay, p.by, p.clamp_at, va[u][0], va[u][1], va[u][2], vb[u][0],
                                  vb[u][1], vb[u][2]);
        }
        for (int i = lane + 32 * kVirtPerLane; i < p.V; i += 32)
            loss += virt_term(Ff, p.ax, p.bx, p.ay, p.by, p.clamp_at, v1[3 * i], v1[3 * i + 1], v1[3 * i + 2], v2[3 * i],
                              v2[3 * i + 1], v2[3 * i + 2]);
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o2);
        loss /= static_cast<float>(p.V);
    }

    pose_head(Ff, Kf, qgf, tgf, p.Rt != nullptr, rtf, loss, p.ax, p.bx, p.ay, p.by, lane, o);
}

// Backward of the head: dL/dF from the upstream gradients of (q L2 error, t L2 error, F-loss) of every (layer, pair).
// One WARP per (layer, pair), like the forward: the 3x3 SVD / quaternion adjoint is a serial fp64 chain that every lane
// runs redundantly (warp-uniform), the V virtual correspondences of the F-loss gradient are spread over the lanes and
// folded with shuffles.  (Round 1 ran one thread per item: the 100-point loop alone was a ~6 k-instruction serial chain.)
struct PoseBwdParams {
    const float* F;        // [L,B,9]
    const float* K;        // [B,9]
    const float* q_gt;     // [B,4]
    const float* t_gt;     // [B,3]
    const float* virt1;    // [B,V,3] or null
    const float* virt2;
    const float* pose_out; // [L,B,32] forward output (which candidates won)
    const float* g_q;      // [L,B] or null
    const float* g_t;      // [L,B] or null
    const float* g_loss;   // [L,B] or null
    int L, B, V;
    float ax, bx, ay, by, clamp_at;
    float* dF;             // [L,B,9]
};

__global__ void __launch_bounds__(128) fepe_pose_bwd_kernel(const PoseBwdParams p) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= p.L * p.B) return;
    const int b = idx % p.B;
    double F[9], K[9], M[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { F[i] = p.F[static_cast<size_t>(idx) * 9 + i]; K[i] = p.K[static_cast<size_t>(b) * 9 + i]; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        M[j] = p.ax * K[j] + p.bx * K[6 + j];
        M[3 + j] = p.ay * K[3 + j] + p.by * K[6 + j];
        M[6 + j] = K[6 + j];
    }
    double dF[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) dF[i] = 0.0;
    const double gq = p.g_q ? static_cast<double>(p.g_q[idx]) : 0.0;
    const double gt = p.g_t ? static_cast<double>(p.g_t[idx]) : 0.0;
    if (gq != 0.0 || gt != 0.0) {
        double FM[9], E[9];
        mat3_mul(F, M, FM);
        mat3_mul_tn(M, FM, E);
        const double Ec[9] = {E[0], E[3], E[6], E[1], E[4], E[7], E[2], E[5], E[8]};
        double qg[4], tg[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) qg[i] = p.q_gt[static_cast<size_t>(b) * 4 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) tg[i] = p.t_gt[static_cast<size_t>(b) * 3 + i];
        const double n = sqrt(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
        const double inv = 1.0 / fmax(n, 1e-12);
        tg[0] *= inv; tg[1] *= inv; tg[2] *= inv;
        const float* po = p.pose_out + static_cast<size_t>(idx) * FEPE_POSE_OUT_FLOATS;
        double Ecb[9];
        pose_head_adjoint(Ec, qg, tg, po[26] == 0.f, po[27] == 0.f, gq, gt, Ecb);
        // dE = Ecb^T ; dF = M dE M^T
        const double Eb[9] = {Ecb[0], Ecb[3], Ecb[6], Ecb[1], Ecb[4], Ecb[7], Ecb[2], Ecb[5], Ecb[8]};
        double ME[9];
        mat3_mul(M, Eb, ME);
        mat3_mul_nt(ME, M, dF);
    }
    const float gl = p.g_loss ? p.g_loss[idx] : 0.f;
    if (gl != 0.f && p.virt1 != nullptr && p.V > 0) {
        float Ff[9], acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { Ff[i] = static_cast<float>(F[i]); acc[i] = 0.f; }
        const float* v1 = p.virt1 + static_cast<size_t>(b) * p.V * 3;
        const float* v2 = p.virt2 + static_cast<size_t>(b) * p.V * 3;
        for (int i = lane; i < p.V; i += 32) {
            const float z1 = v1[3 * i + 2], z2 = v2[3 * i + 2];
            const float u1 = fmaf(p.ax, v1[3 * i], p.bx * z1), w1 = fmaf(p.ay, v1[3 * i + 1], p.by * z1);
            const float u2 = fmaf(p.ax, v2[3 * i], p.bx * z2), w2 = fmaf(p.ay, v2[3 * i + 1], p.by * z2);
            const float l10 = u2 * Ff[0] + w2 * Ff[3] + z2 * Ff[6];
            const float l11 = u2 * Ff[1] + w2 * Ff[4] + z2 * Ff[7];
            const float l12 = u2 * Ff[2] + w2 * Ff[5] + z2 * Ff[8];
            const float l20 = Ff[0] * u1 + Ff[1] * w1 + Ff[2] * z1;
            const float l21 = Ff[3] * u1 + Ff[4] * w1 + Ff[5] * z1;
            const float dd = l10 * u1 + l11 * w1 + l12 * z1;
            const float m1 = sqrtf(l10 * l10 + l11 * l11), m2 = sqrtf(l20 * l20 + l21 * l21);
            const float i1 = 1.0f / (m1 + 1e-6f), i2 = 1.0f / (m2 + 1e-6f);
            const float dist = fabsf(dd) * (i1 + i2);
            if (dist > p.clamp_at) continue;                      // clamp(max=c): zero gradient above c
            const float sg = (dd > 0.f) ? 1.f : ((dd < 0.f) ? -1.f : 0.f);
            const float S12 = sg * (i1 + i2);
            const float a1 = -fabsf(dd) * i1 * i1 / fmaxf(m1, 1e-30f);
            const float a2 = -fabsf(dd) * i2 * i2 / fmaxf(m2, 1e-30f);
            const float uk0 = S12 * u1 + a1 * l10, uk1 = S12 * w1 + a1 * l11, uk2 = S12 * z1;
            const float vj0 = a2 * l20, vj1 = a2 * l21;
            acc[0] += u2 * uk0 + vj0 * u1; acc[1] += u2 * uk1 + vj0 * w1; acc[2] += u2 * uk2 + vj0 * z1;
            acc[3] += w2 * uk0 + vj1 * u1; acc[4] += w2 * uk1 + vj1 * w1; acc[5] += w2 * uk2 + vj1 * z1;
            acc[6] += z2 * uk0;            acc[7] += z2 * uk1;            acc[8] += z2 * uk2;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = warp_sum(acc[i]);
        const double sc = static_cast<double>(gl) / static_cast<double>(p.V);
#pragma unroll
        for (int i = 0; i < 9; ++i) dF[i] += sc * static_cast<double>(acc[i]);
    }
    if (lane < 9) {
        double out = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) out = (lane == i) ? dF[i] : out;
        p.dF[static_cast<size_t>(idx) * 9 + lane] = static_cast<float>(out);
    }
}

}  // namespace fepe

extern "C" int fepe_pose_bwd(const float* F, const float* K, int L, int B, float ax, float bx, float ay, float by,
                             const float* q_gt, const float* t_gt, const float* virt1, const float* virt2, int V,
                             float clamp_at, const float* pose_out, const float* g_q, const float* g_t,
                             const float* g_loss, float* dF, void* stream) {
    if (L == 0 || B == 0) return 0;
    if (!F || !K || !q_gt || !t_gt || !pose_out || !dF || L < 0 || B < 0 || V < 0) return FEPE_E_BADARG;
    fepe::PoseBwdParams p{F, K, q_gt, t_gt, virt1, virt2, pose_out, g_q, g_t, g_loss, L, B, V, ax, bx, ay, by, clamp_at, dF};
    const int n = L * B;
    fepe::fepe_pose_bwd_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}

namespace fepe {
// `pdl`: programmatic dependent launch -- the grid may be scheduled while the preceding kernel of the stream is still
// running; the kernel fetches everything that does not depend on F and then blocks in griddepcontrol.wait until that
// kernel has completed.  ONLY safe when the predecessor is known to write nothing but F / residual / epi, i.e. on the
// internal two-launch path of fit_fwd_impl (fepe_fit.cu).  The public entry point launches with ordinary stream
// serialization: there the preceding kernel may be the producer of K, q_gt, t_gt, Rt or the virtual points, which this
// kernel reads BEFORE its griddepcontrol.wait.
int launch_pose_fwd(const PoseParams& p, cudaStream_t stream, bool pdl) {
    const int n = p.L * p.B;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((n + 3) / 4);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, fepe_pose_fwd_kernel, p));
}
}  // namespace fepe

extern "C" int fepe_pose_fwd(const float* F, const float* K, int L, int B, float ax, float bx, float ay, float by,
                             const float* q_gt, const float* t_gt, const float* Rt_scene, const float* virt1,
                             const float* virt2, int V, float clamp_at, float* out, void* stream) {
    if (L == 0 || B == 0) return 0;
    if (!F || !K || !q_gt || !t_gt || !out || L < 0 || B < 0 || V < 0) return FEPE_E_BADARG;
    if ((virt1 == nullptr) != (virt2 == nullptr)) return FEPE_E_BADARG;
    fepe::PoseParams p{F, K, q_gt, t_gt, Rt_scene, virt1, virt2, L, B, V, ax, bx, ay, by, clamp_at, out};
    return fepe::launch_pose_fwd(p, static_cast<cudaStream_t>(stream), /*pdl=*/false);
}
