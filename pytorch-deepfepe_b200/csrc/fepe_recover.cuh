// Small fp64 pieces of the validation pose recovery (fepe_recover_pose, fepe_recover.cu): what
// cv2.recoverPose does for deepFEPE/dsac_tools/utils_F.py:936 (goodCorr_eval_nondecompose, called per sample from
// train_good_utils.py:553-646 val_rt through a pebble process pool).  OpenCV is a third-party dependency of the
// reference (opencv-python 3.4.2.16 pinned in requirements.txt); its published algorithm is
//   decomposeEssentialMat -> four candidates [R1|t],[R2|t],[R1|-t],[R2|-t] -> linear (DLT) triangulation of every
//   correspondence against [I|0] -> cheirality + distance test -> the candidate with the most points in front.
// __host__ __device__ like fepe_math.cuh: tests/host_shim.cpp compiles these with g++ and tests/test_recover_pose_host.py
// compares them with cv2.recoverPose itself.
#pragma once

#include "fepe_math.cuh"

namespace fepe {

// One Jacobi rotation annihilating a[p][q] of the symmetric 4x4 `a` (full storage), accumulated into `v`.
template <int P, int Q>
FEPE_HD void jacobi4_rotate(double (&a)[16], double (&v)[16]) {
    const double apq = a[P * 4 + Q];
    const double app = a[P * 4 + P], aqq = a[Q * 4 + Q];
    if (fabs(apq) <= 1e-300) return;
    const double tau = (aqq - app) / (2.0 * apq);
    const double t = ((tau >= 0.0) ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
    const double c = 1.0 / sqrt(1.0 + t * t);
    const double s = t * c;
    a[P * 4 + P] = app - t * apq;
    a[Q * 4 + Q] = aqq + t * apq;
    a[P * 4 + Q] = 0.0;
    a[Q * 4 + P] = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k != P && k != Q) {
            const double akp = a[k * 4 + P], akq = a[k * 4 + Q];
            const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
            a[k * 4 + P] = np_; a[P * 4 + k] = np_;
            a[k * 4 + Q] = nq_; a[Q * 4 + k] = nq_;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double vkp = v[k * 4 + P], vkq = v[k * 4 + Q];
        v[k * 4 + P] = c * vkp - s * vkq;
        v[k * 4 + Q] = s * vkp + c * vkq;
    }
}

// Eigenvector of the smallest eigenvalue of a symmetric 4x4 (cyclic Jacobi: exact to rounding whatever the gaps, which
// is what the comparison with an SVD-based triangulation on badly conditioned points needs).  `a` is destroyed.
FEPE_HD void sym4_smallest_eigvec(double (&a)[16], double (&x)[4]) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (i % 5 == 0) ? 1.0 : 0.0;
    const double scale = a[0] + a[5] + a[10] + a[15];
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a[1]) + fabs(a[2]) + fabs(a[3]) + fabs(a[6]) + fabs(a[7]) + fabs(a[11]);
        if (off <= 1e-22 * scale) break;
        jacobi4_rotate<0, 1>(a, v); jacobi4_rotate<0, 2>(a, v); jacobi4_rotate<0, 3>(a, v);
        jacobi4_rotate<1, 2>(a, v); jacobi4_rotate<1, 3>(a, v); jacobi4_rotate<2, 3>(a, v);
    }
    int k = 0;
    double best = a[0];
    if (a[5] < best) { best = a[5]; k = 1; }
    if (a[10] < best) { best = a[10]; k = 2; }
    if (a[15] < best) { best = a[15]; k = 3; }
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (k == 0) ? v[i * 4] : (k == 1) ? v[i * 4 + 1] : (k == 2) ? v[i * 4 + 2] : v[i * 4 + 3];
}

// Linear triangulation of one correspondence between P0 = [I|0] and P (row-major 3x4), normalised image points:
// null vector (smallest right singular vector) of the 4x4 DLT matrix (OpenCV triangulate.cpp), via its Gram matrix.
FEPE_HD void triangulate_dlt(double x1, double y1, double x2, double y2, const double (&P)[12], double (&X)[4]) {
    double r2[4], r3[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r2[k] = x2 * P[8 + k] - P[k];
        r3[k] = y2 * P[8 + k] - P[4 + k];
    }
    // rows 0,1 are (-1,0,x1,0) and (0,-1,y1,0)
    double a[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i * 4 + j] = r2[i] * r2[j] + r3[i] * r3[j];
    a[0] += 1.0;  a[5] += 1.0;
    a[2] -= x1;   a[8] -= x1;
    a[6] -= y1;   a[9] -= y1;
    a[10] += x1 * x1 + y1 * y1;
    sym4_smallest_eigvec(a, X);
}

// cv2.recoverPose's test of one triangulated point for the candidate P.
FEPE_HD bool cheirality_ok(const double (&X)[4], const double (&P)[12], double thresh) {
    if (!(X[2] * X[3] > 0.0)) return false;
    const double iw = 1.0 / X[3];
    const double z1 = X[2] * iw;
    if (!(z1 < thresh)) return false;
    const double z2 = (P[8] * X[0] + P[9] * X[1] + P[10] * X[2]) * iw + P[11];
    return (z2 > 0.0) && (z2 < thresh);
}

// Candidate k of cv2.recoverPose: rotation R1 (k even) or R2 (k odd), translation +t (k < 2) or -t.
FEPE_HD void recover_candidate(int k, const double (&R1)[9], const double (&R2)[9], const double (&t)[3],
                               double (&P)[12]) {
    const double sg = (k < 2) ? 1.0 : -1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) P[r * 4 + c] = (k & 1) ? R2[r * 3 + c] : R1[r * 3 + c];
        P[r * 4 + 3] = sg * t[r];
    }
}

// cv2's selection rule (ties go to the earlier candidate in its if / else-if chain).
FEPE_HD int recover_select(int g1, int g2, int g3, int g4) {
    if (g1 >= g2 && g1 >= g3 && g1 >= g4) return 0;
    if (g2 >= g1 && g2 >= g3 && g2 >= g4) return 1;
    if (g3 >= g1 && g3 >= g2 && g3 >= g4) return 2;
    return 3;
}

// goodCorr_eval_nondecompose (utils_F.py:937-940): errors of the recovered (R, t) against the scene motion
// Rs|ts = delta_Rtijs_4_4: err_q = angle(R^T Rs) (|cv2.Rodrigues|), err_t = angle(R^T t, Rs^T ts), degrees.
FEPE_HD void recover_errors(const double (&R)[9], const double (&t)[3], const double (&Rs)[9], const double (&ts)[3],
                            double& err_q, double& err_t) {
    double D[9];
    mat3_mul_tn(R, Rs, D);
    const double c = 0.5 * (D[0] + D[4] + D[8] - 1.0);
    const double s = 0.5 * sqrt((D[7] - D[5]) * (D[7] - D[5]) + (D[2] - D[6]) * (D[2] - D[6]) + (D[3] - D[1]) * (D[3] - D[1]));
    err_q = atan2(s, c) * 57.29577951308232;
    double a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        a[i] = R[i] * t[0] + R[3 + i] * t[1] + R[6 + i] * t[2];          // R^T t
        b[i] = Rs[i] * ts[0] + Rs[3 + i] * ts[1] + Rs[6 + i] * ts[2];    // Rs^T ts
    }
    const double la = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) + 1e-10;
    const double lb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) + 1e-10;
    double d = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (la * lb + 1e-10);   // utils_geo.vector_angle (:175-179)
    d = d > 1.0 ? 1.0 : (d < -1.0 ? -1.0 : d);
    err_t = acos(d) * 57.29577951308232;
}

}  // namespace fepe
