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

__global__ void __launch_bounds__(128) fepe_pose_fwd_kernel(const PoseParams p) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per (layer, pair)
    const int lane = threadIdx.x & 31;
    if (idx >= p.L * p.B) return;
    const int b = idx % p.B;
    float* o = p.out + static_cast<size_t>(idx) * FEPE_POSE_OUT_FLOATS;

    double F[9], K[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { F[i] = p.F[static_cast<size_t>(idx) * 9 + i]; K[i] = p.K[static_cast<size_t>(b) * 9 + i]; }
    // M = T K with T = [[ax,0,bx],[0,ay,by],[0,0,1]];  E = M^T F M
    double M[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        M[j] = p.ax * K[j] + p.bx * K[6 + j];
        M[3 + j] = p.ay * K[3 + j] + p.by * K[6 + j];
        M[6 + j] = K[6 + j];
    }
    double FM[9], E[9];
    mat3_mul(F, M, FM);
    mat3_mul_tn(M, FM, E);
    if (lane < 9) o[lane] = static_cast<float>(E[lane]);

    // decompose E^T
    double Et[9] = {E[0], E[3], E[6], E[1], E[4], E[7], E[2], E[5], E[8]};
    double R1[9], R2[9], t[3], U[9], S[3], V[9];
    essential_decompose(Et, R1, R2, t, U, S, V);
    double q1[4], q2[4];
    rot_to_quat(R1, q1);
    rot_to_quat(R2, q2);
    double qg[4], tg[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) qg[i] = p.q_gt[static_cast<size_t>(b) * 4 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tg[i] = p.t_gt[static_cast<size_t>(b) * 3 + i];
    {   // F.normalize(t_gt, p=2, dim=0): x / max(|x|, 1e-12)
        const double n = sqrt(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
        const double inv = 1.0 / fmax(n, 1e-12);
        tg[0] *= inv; tg[1] *= inv; tg[2] *= inv;
    }
    double eq1 = 0, eq2 = 0, et1 = 0, et2 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { eq1 += (q1[i] - qg[i]) * (q1[i] - qg[i]); eq2 += (q2[i] - qg[i]) * (q2[i] - qg[i]); }
#pragma unroll
    for (int i = 0; i < 3; ++i) { et1 += (t[i] - tg[i]) * (t[i] - tg[i]); et2 += (-t[i] - tg[i]) * (-t[i] - tg[i]); }
    eq1 = sqrt(eq1); eq2 = sqrt(eq2); et1 = sqrt(et1); et2 = sqrt(et2);
    const bool q_first = eq1 < eq2;     // strict, like the reference's q12_error[0] < q12_error[1]
    const bool t_first = et1 < et2;
    const double tsg = t_first ? 1.0 : -1.0;
    float res[FEPE_POSE_OUT_FLOATS];      // lane-uniform results, written once at the end
#pragma unroll
    for (int i = 0; i < 9; ++i) res[9 + i] = static_cast<float>(q_first ? R1[i] : R2[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) res[18 + i] = static_cast<float>(tsg * t[i]);
    res[21] = static_cast<float>(q_first ? eq1 : eq2);
    res[22] = static_cast<float>(t_first ? et1 : et2);

    // angular metrics
    float r_ang = 0.f;
    if (p.Rt != nullptr) {
        // R_gt = inverse(Rt)[:3,:3] = R_scene^T ; angle of R_est R_gt^T = R_est R_scene
        const float* rt = p.Rt + static_cast<size_t>(b) * 16;
        double Rs[9] = {rt[0], rt[1], rt[2], rt[4], rt[5], rt[6], rt[8], rt[9], rt[10]};
        double Re[9], D[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Re[i] = q_first ? R1[i] : R2[i];
        mat3_mul(Re, Rs, D);
        const double c = 0.5 * (D[0] + D[4] + D[8] - 1.0);
        const double s = 0.5 * sqrt((D[7] - D[5]) * (D[7] - D[5]) + (D[2] - D[6]) * (D[2] - D[6]) +
                                    (D[3] - D[1]) * (D[3] - D[1]));
        r_ang = static_cast<float>(atan2(s, c) * 57.29577951308232);
    }
    res[23] = r_ang;
    {
        const double dot = tsg * (t[0] * tg[0] + t[1] * tg[1] + t[2] * tg[2]);
        const double c = fmin(1.0, fmax(-1.0, dot));
        res[24] = static_cast<float>(acos(c) * 57.29577951308232);
    }

    // F-loss over the virtual correspondences (fp32 like the reference)
    float loss = 0.f;
    if (p.virt1 != nullptr && p.V > 0) {
        float Ff[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Ff[i] = static_cast<float>(F[i]);
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
            const float d = fabsf(dd) * (1.0f / (sqrtf(l10 * l10 + l11 * l11) + 1e-6f) +
                                         1.0f / (sqrtf(l20 * l20 + l21 * l21) + 1e-6f));
            loss += fminf(d, p.clamp_at);
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o2);
        loss /= static_cast<float>(p.V);
    }
    res[25] = loss;
    res[26] = q_first ? 0.f : 1.f;
    res[27] = t_first ? 0.f : 1.f;
    res[28] = static_cast<float>(S[0]); res[29] = static_cast<float>(S[1]); res[30] = static_cast<float>(S[2]);
    res[31] = 0.f;
#pragma unroll
    for (int i = 9; i < FEPE_POSE_OUT_FLOATS; ++i) {
        if (lane == i) o[i] = res[i];
    }
}

}  // namespace fepe

extern "C" int fepe_pose_fwd(const float* F, const float* K, int L, int B, float ax, float bx, float ay, float by,
                             const float* q_gt, const float* t_gt, const float* Rt_scene, const float* virt1,
                             const float* virt2, int V, float clamp_at, float* out, void* stream) {
    if (L == 0 || B == 0) return 0;
    if (!F || !K || !q_gt || !t_gt || !out || L < 0 || B < 0 || V < 0) return FEPE_E_BADARG;
    if ((virt1 == nullptr) != (virt2 == nullptr)) return FEPE_E_BADARG;
    fepe::PoseParams p{F, K, q_gt, t_gt, Rt_scene, virt1, virt2, L, B, V, ax, bx, ay, by, clamp_at, out};
    const int n = L * B;
    fepe::fepe_pose_fwd_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}
