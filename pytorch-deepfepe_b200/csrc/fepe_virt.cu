// fepe_gt_virt: the ground-truth side of a batch on the device.
//
// Replaces, per sample, the host work of the reference's dataset before the loss of the hot path can be evaluated
// (deepFEPE/datasets/kitti_odo_corr.py:290-302 get_E_F, :526-566 __getitem__):
//     E, F = utils_F.E_F_from_Rt_np(R, t, K)                                               (utils_F.py:835-846)
//     pts1_virt_normalized, pts2_virt_normalized, pts1_virt, pts2_virt =
//         utils_misc.get_virt_x1x2_np(image_size, F, K, pts1_virt_b, pts2_virt_b)          (utils_misc.py:173-199)
//         = cv2.correctMatches(F, pts2_virt_b, pts1_virt_b), NaN -> 0, homogeneous, K^-1 applied
//     q_cam, t_cam from inv(Rt_scene); q_scene, t_scene from Rt_scene                      (utils_geo.py:88-117)
// One thread per (pair, grid point); every thread of a pair rebuilds the pair's 3x3 algebra (a few hundred fp64
// operations) instead of exchanging it, thread 0 of the pair stores it.  Output is 24 B per grid point in, 36 B out:
// the kernel is bound by the latency of one thread's fp64 chain (3x3 Jacobi SVD + the Durand-Kerner sweeps OpenCV
// runs) and exists to take the last per-sample host loop -- and its H2D copies -- out of the training input path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fepe_b200.h"
#include "fepe_virt.cuh"

namespace fepe {

constexpr int kVirtThreads = 128;

struct VirtParams {
    const float* K;        // [B,9]
    const float* Rt;       // [B,16] scene motion, or null when F_in is given
    const float* F_in;     // [B,9] or null
    const float* grid1;    // [P,2] pts1_virt_b
    const float* grid2;    // [P,2] pts2_virt_b
    int B, P;
    float* gt;             // [B,FEPE_GT_FLOATS] or null
    float* pts1;           // [B,P,3]
    float* pts2;           // [B,P,3]
    float* ptsn;           // [B,P,3] or null
};

__global__ void __launch_bounds__(kVirtThreads) fepe_gt_virt_kernel(const VirtParams p) {
    const int pair = blockIdx.y;
    const int pt = blockIdx.x * kVirtThreads + threadIdx.x;
    double K[9], Kinv[9], F[9], gt[32];
#pragma unroll
    for (int i = 0; i < 9; ++i) K[i] = static_cast<double>(__ldg(p.K + pair * 9 + i));
    if (p.Rt != nullptr) {
        double Rt[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) Rt[i] = static_cast<double>(__ldg(p.Rt + pair * 16 + i));
        gt_from_motion(K, Rt, gt, Kinv);
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] = gt[9 + i];
        if (pt == 0 && p.gt != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p.gt[pair * FEPE_GT_FLOATS + i] = static_cast<float>(gt[i]);
        }
    } else {
        inv3(K, Kinv);
    }
    if (p.F_in != nullptr) {
#pragma unroll
        for (int i = 0; i < 9; ++i) F[i] = static_cast<double>(__ldg(p.F_in + pair * 9 + i));
    }
    if (pt >= p.P) return;
    // the reference's call is cv2.correctMatches(F, pts2_virt_b, pts1_virt_b): the SECOND grid plays OpenCV's points1
    const double x1 = static_cast<double>(__ldg(p.grid2 + 2 * pt)), y1 = static_cast<double>(__ldg(p.grid2 + 2 * pt + 1));
    const double x2 = static_cast<double>(__ldg(p.grid1 + 2 * pt)), y2 = static_cast<double>(__ldg(p.grid1 + 2 * pt + 1));
    double o[4];
    const bool ok = correct_match_pair(F, x1, y1, x2, y2, o);
    float q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q[i] = static_cast<float>(o[i]);                       // cv2 hands back the points' own dtype (float32)
        if (!ok || q[i] != q[i]) q[i] = 0.f;                   // utils_misc.py:177-178
    }
    const size_t row = (static_cast<size_t>(pair) * p.P + pt) * 3;
    p.pts1[row] = q[0]; p.pts1[row + 1] = q[1]; p.pts1[row + 2] = 1.f;
    p.pts2[row] = q[2]; p.pts2[row + 1] = q[3]; p.pts2[row + 2] = 1.f;
    if (p.ptsn != nullptr) {                                   // K^-1 [x1', y1', 1]: both normalised outputs (:197-198)
        const double hx = static_cast<double>(q[0]), hy = static_cast<double>(q[1]);
#pragma unroll
        for (int r = 0; r < 3; ++r)
            p.ptsn[row + r] = static_cast<float>(Kinv[3 * r] * hx + Kinv[3 * r + 1] * hy + Kinv[3 * r + 2]);
    }
}

}  // namespace fepe

extern "C" int fepe_gt_virt(const float* K, const float* Rt_scene, const float* F_in, const float* grid1,
                            const float* grid2, int B, int P, float* gt, float* pts1_virt, float* pts2_virt,
                            float* pts_virt_normalized, void* stream) {
    if (B == 0) return 0;
    if (!K || (!Rt_scene && !F_in) || !grid1 || !grid2 || !pts1_virt || !pts2_virt || B < 0 || P <= 0) return FEPE_E_BADARG;
    if (gt != nullptr && Rt_scene == nullptr) return FEPE_E_BADARG;
    if (B > 65535) return FEPE_E_TOOLARGE;
    fepe::VirtParams p{K, Rt_scene, F_in, grid1, grid2, B, P, gt, pts1_virt, pts2_virt, pts_virt_normalized};
    const dim3 grid((P + fepe::kVirtThreads - 1) / fepe::kVirtThreads, B);
    fepe::fepe_gt_virt_kernel<<<grid, fepe::kVirtThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return static_cast<int>(cudaGetLastError());
}
