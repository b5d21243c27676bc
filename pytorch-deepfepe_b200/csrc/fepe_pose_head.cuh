// Pose / loss head shared by the stand-alone kernel (fepe_pose.cu) and the fused fit+pose latency kernel
// (fepe_fit.cu): launch parameters, the virtual-correspondence F-loss term and the per-item dependent chain
// E = K^T T^T F T K -> SVD -> {R1,R2},{+-t} -> quaternions -> errors (see fepe_pose.cu for the reference lines).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_math.cuh"

namespace fepe {

struct PoseParams {
    const float* F;       // [L,B,9]
    const float* K;       // [B,9]
    const float* q_gt;    // [B,4]   (w,x,y,z)
    const float* t_gt;    // [B,3]   un-normalised
    const float* Rt;      // [B,16]  scene motion 4x4 or null
    const float* virt1;   // [B,V,3] or null
    const float* virt2;
    int L, B, V;
    float ax, bx, ay, by, clamp_at;
    float* out;           // [L,B,FEPE_POSE_OUT_FLOATS]
};

// fepe_pose.cu; `pdl` = programmatic dependent launch, only for a predecessor that writes nothing this kernel reads
// before its griddepcontrol.wait (the internal fit -> head pair of fepe_fit_pose_fwd)
int launch_pose_fwd(const PoseParams& p, cudaStream_t stream, bool pdl);

constexpr int kVirtPerLane = 4;     // virtual correspondences held in registers per lane (V <= 128 in one trip)

__device__ __forceinline__ float virt_term(const float (&Ff)[9], float ax, float bx, float ay, float by, float clamp_at,
                                           float x1, float y1, float z1, float x2, float y2, float z2) {
    const float u1 = fmaf(ax, x1, bx * z1), w1 = fmaf(ay, y1, by * z1);
    const float u2 = fmaf(ax, x2, bx * z2), w2 = fmaf(ay, y2, by * z2);
    const float l10 = u2 * Ff[0] + w2 * Ff[3] + z2 * Ff[6];
    const float l11 = u2 * Ff[1] + w2 * Ff[4] + z2 * Ff[7];
    const float l12 = u2 * Ff[2] + w2 * Ff[5] + z2 * Ff[8];
    const float l20 = Ff[0] * u1 + Ff[1] * w1 + Ff[2] * z1;
    const float l21 = Ff[3] * u1 + Ff[4] * w1 + Ff[5] * z1;
    const float dd = l10 * u1 + l11 * w1 + l12 * z1;
    const float d = fabsf(dd) * (1.0f / (sqrtf(l10 * l10 + l11 * l11) + 1e-6f) +
                                 1.0f / (sqrtf(l20 * l20 + l21 * l21) + 1e-6f));
    return fminf(d, clamp_at);
}

// The warp-uniform chain of one (layer, pair): every lane computes all of it, lane i writes out[i].
// `loss` is the already reduced F-loss of the item (0 without virtual correspondences).
__device__ __forceinline__ void pose_head(const float (&Ff)[9], const float (&Kf)[9], const float (&qgf)[4],
                                          const float (&tgf)[3], bool has_rt, const float (&rtf)[9], float loss,
                                          float ax, float bx, float ay, float by, int lane, float* __restrict__ o) {
    double F[9], K[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { F[i] = Ff[i]; K[i] = Kf[i]; }
    // M = T K with T = [[ax,0,bx],[0,ay,by],[0,0,1]];  E = M^T F M
    double M[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        M[j] = ax * K[j] + bx * K[6 + j];
        M[3 + j] = ay * K[3 + j] + by * K[6 + j];
        M[6 + j] = K[6 + j];
    }
    double FM[9], E[9];
    mat3_mul(F, M, FM);
    mat3_mul_tn(M, FM, E);

    // decompose E^T
    double Et[9] = {E[0], E[3], E[6], E[1], E[4], E[7], E[2], E[5], E[8]};
    double R1[9], R2[9], t[3], U[9], S[3], V[9];
    essential_decompose(Et, R1, R2, t, U, S, V);
    double q1[4], q2[4];
    rot_to_quat(R1, q1);
    rot_to_quat(R2, q2);
    double qg[4], tg[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) qg[i] = qgf[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tg[i] = tgf[i];
    {   // F.normalize(t_gt, p=2, dim=0): x / max(|x|, 1e-12)
        const double n2 = tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2];
        const double inv = (n2 > 1e-24) ? fast_rsqrt(n2) : 1e12;
        tg[0] *= inv; tg[1] *= inv; tg[2] *= inv;
    }
    double eq1 = 0, eq2 = 0, et1 = 0, et2 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { eq1 += (q1[i] - qg[i]) * (q1[i] - qg[i]); eq2 += (q2[i] - qg[i]) * (q2[i] - qg[i]); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { et1 += (t[i] - tg[i]) * (t[i] - tg[i]); et2 += (-t[i] - tg[i]) * (-t[i] - tg[i]); }
    const bool q_first = eq1 < eq2;     // strict, like the reference's q12_error[0] < q12_error[1] (sqrt is monotone)
    const bool t_first = et1 < et2;
    const double tsg = t_first ? 1.0 : -1.0;
    float res[FEPE_POSE_OUT_FLOATS];      // lane-uniform results, written once at the end
#pragma unroll
    for (int i = 0; i < 9; ++i) res[i] = static_cast<float>(E[i]);
    double Re[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { Re[i] = q_first ? R1[i] : R2[i]; res[9 + i] = static_cast<float>(Re[i]); }
#pragma unroll
    for (int i = 0; i < 3; ++i) res[18 + i] = static_cast<float>(tsg * t[i]);
    res[21] = static_cast<float>(fast_sqrt(q_first ? eq1 : eq2));
    res[22] = static_cast<float>(fast_sqrt(t_first ? et1 : et2));

    // angular metrics
    float r_ang = 0.f;
    if (has_rt) {
        // R_gt = inverse(Rt)[:3,:3] = R_scene^T ; angle of R_est R_gt^T = R_est R_scene
        double Rs[9], D[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rs[i] = rtf[i];
        mat3_mul(Re, Rs, D);
        const double c = 0.5 * (D[0] + D[4] + D[8] - 1.0);
        const double s = 0.5 * fast_sqrt((D[7] - D[5]) * (D[7] - D[5]) + (D[2] - D[6]) * (D[2] - D[6]) +
                                         (D[3] - D[1]) * (D[3] - D[1]));
        // the arguments are fp64-accurate; fp32 atan2 of them is good to ~1e-5 deg and far cheaper than the fp64 routine
        r_ang = atan2f(static_cast<float>(s), static_cast<float>(c)) * 57.29577951f;
    }
    res[23] = r_ang;
    {
        // angle between unit vectors as atan2(|a x b|, a.b): no acos cancellation near 0, fp32 evaluation
        const double dot = tsg * (t[0] * tg[0] + t[1] * tg[1] + t[2] * tg[2]);
        const double cx = t[1] * tg[2] - t[2] * tg[1], cy = t[2] * tg[0] - t[0] * tg[2], cz = t[0] * tg[1] - t[1] * tg[0];
        const double sn = fast_sqrt(cx * cx + cy * cy + cz * cz);
        res[24] = atan2f(static_cast<float>(sn), static_cast<float>(dot)) * 57.29577951f;
    }
    res[25] = loss;
    res[26] = q_first ? 0.f : 1.f;
    res[27] = t_first ? 0.f : 1.f;
    res[28] = static_cast<float>(S[0]); res[29] = static_cast<float>(S[1]); res[30] = static_cast<float>(S[2]);
    res[31] = 0.f;
    float mine = 0.f;                     // one coalesced 128-byte store per item
#pragma unroll
    for (int i = 0; i < FEPE_POSE_OUT_FLOATS; ++i) mine = (lane == i) ? res[i] : mine;
    o[lane] = mine;
}

}  // namespace fepe
