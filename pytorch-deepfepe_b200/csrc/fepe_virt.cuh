// Scalar pieces of fepe_gt_virt: the two-view point correction OpenCV's cv2.correctMatches
// performs for deepFEPE/dsac_tools/utils_misc.py:173-199 get_virt_x1x2_np, and the ground-truth E / F / q / t of a
// sample (deepFEPE/dsac_tools/utils_F.py:835-846 E_F_from_Rt_np, utils_geo.py:88-117 R_to_q_np,
// deepFEPE/datasets/kitti_odo_corr.py:290-302, :526-566).  __host__ __device__ so tests/test_virt_points_host.py can
// check them against cv2 and the reference's outputs on a machine without a GPU.
//
// The polynomial solver is a restatement of cv::solvePoly as icvCorrectMatches calls it (100 Durand-Kerner sweeps,
// leading coefficients <= DBL_EPSILON dropped, uncomputed root slots ~ 0): with a pixel-unit F OpenCV ends up solving a
// linear or quadratic truncation, and parity with the reference means reproducing that.
// Its complex arithmetic is written with explicitly rounded products so that nvcc's fused multiply-add contraction
// cannot change the rounding sequence of a sweep relative to the x86 library.
#pragma once
#include "fepe_math.cuh"

namespace fepe {

#ifdef __CUDA_ARCH__
#define FEPE_MUL(a, b) __dmul_rn((a), (b))
#define FEPE_ADD(a, b) __dadd_rn((a), (b))
#else
#define FEPE_MUL(a, b) ((a) * (b))
#define FEPE_ADD(a, b) ((a) + (b))
#endif

// cv::solvePoly for real coefficients k[0..6] (k[i] multiplies t^i): real parts of the six output slots.
FEPE_HD int solve_poly6_cv(const double (&k)[7], double (&t_re)[6]) {
    int n = 6;
    while (n > 1 && fabs(k[n]) <= 2.220446049250313e-16) --n;
    double re[6], im[6];
    {
        double pr = 1.0, pi = 0.0;
        for (int i = 0; i < n; ++i) {
            re[i] = pr; im[i] = pi;
            const double nr = pr - pi, ni = pr + pi;        // p * (1 + i)
            pr = nr; pi = ni;
        }
    }
    for (int it = 0; it < 100; ++it) {
        double max_diff = 0.0;
        for (int i = 0; i < n; ++i) {
            const double p_r = re[i], p_i = im[i];
            double nr = k[n], ni = 0.0, dr = k[n], di = 0.0;
            for (int j = 0; j < n; ++j) {
                const double tr = FEPE_ADD(FEPE_ADD(FEPE_MUL(nr, p_r), -FEPE_MUL(ni, p_i)), k[n - j - 1]);
                const double ti = FEPE_ADD(FEPE_MUL(nr, p_i), FEPE_MUL(ni, p_r));
                nr = tr; ni = ti;
                if (j != i) {
                    const double qr = p_r - re[j], qi = p_i - im[j];
                    if (qr != 0.0 || qi != 0.0) {
                        const double er = FEPE_ADD(FEPE_MUL(dr, qr), -FEPE_MUL(di, qi));
                        const double ei = FEPE_ADD(FEPE_MUL(dr, qi), FEPE_MUL(di, qr));
                        dr = er; di = ei;
                    }
                }
            }
            const double s = 1.0 / FEPE_ADD(FEPE_MUL(dr, dr), FEPE_MUL(di, di));
            const double cr = FEPE_MUL(FEPE_ADD(FEPE_MUL(nr, dr), FEPE_MUL(ni, di)), s);
            const double ci = FEPE_MUL(FEPE_ADD(-FEPE_MUL(nr, di), FEPE_MUL(ni, dr)), s);
            re[i] = p_r - cr; im[i] = p_i - ci;
            const double mag = sqrt(FEPE_ADD(FEPE_MUL(cr, cr), FEPE_MUL(ci, ci)));
            max_diff = mag > max_diff ? mag : max_diff;
        }
        if (max_diff <= 0.0) break;
    }
    for (int i = 0; i < 6; ++i) t_re[i] = (i < n) ? re[i] : 0.0;
    return n;
}

// One correspondence of cv2.correctMatches(F, (x1,y1), (x2,y2)): new points with p2'^T F p1' = 0.
// Returns false when OpenCV's result is NaN (the value at t = infinity wins); the reference then stores 0.
FEPE_HD bool correct_match_pair(const double (&F)[9], double x1, double y1, double x2, double y2, double (&o)[4]) {
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
    // F0 = T2^T F T1,  T = [[1,0,x],[0,1,y],[0,0,1]]
    double G[9], F0[9];
    for (int r = 0; r < 3; ++r) {                       // G = F T1
        G[3 * r] = F[3 * r]; G[3 * r + 1] = F[3 * r + 1];
        G[3 * r + 2] = F[3 * r] * x1 + F[3 * r + 1] * y1 + F[3 * r + 2];
    }
    for (int c = 0; c < 3; ++c) {                       // F0 = T2^T G
        F0[c] = G[c]; F0[3 + c] = G[3 + c];
        F0[6 + c] = x2 * G[c] + y2 * G[3 + c] + G[6 + c];
    }
    double U[9], S[3], V[9];
    svd3(F0, U, S, V);
    double e1x = V[2], e1y = V[5], e1z = V[8], e2x = U[2], e2y = U[5], e2z = U[8];
    const double n1 = sqrt(e1x * e1x + e1y * e1y), n2 = sqrt(e2x * e2x + e2y * e2y);
    if (n1 == 0.0 || n2 == 0.0) return true;            // epipole at infinity: OpenCV leaves the pair unchanged
    e1x /= n1; e1y /= n1; e1z /= n1;
    e2x /= n2; e2y /= n2; e2z /= n2;
    // F1 = R2 F0 R1^T,  R = [[ex,ey,0],[-ey,ex,0],[0,0,1]]; only its lower-right 2x2 is used
    double H[9];
    for (int c = 0; c < 3; ++c) {                       // H = R2 F0
        H[c] = e2x * F0[c] + e2y * F0[3 + c];
        H[3 + c] = -e2y * F0[c] + e2x * F0[3 + c];
        H[6 + c] = F0[6 + c];
    }
    const double a = -e1y * H[3] + e1x * H[4], b = H[5];
    const double c = -e1y * H[6] + e1x * H[7], d = H[8];
    const double f1 = e1z, f2 = e2z;
    const double f1s = f1 * f1, f2s = f2 * f2, f14 = f1s * f1s, f24 = f2s * f2s;
    const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d;
    double k[7];
    k[6] = b * c2 * f14 * a - a2 * d * f14 * c;
    k[5] = f24 * c2 * c2 + 2 * a2 * f2s * c2 - a2 * d2 * f14 + b2 * c2 * f14 + a2 * a2;
    k[4] = 4 * a2 * a * b + 2 * b * c2 * f1s * a + 4 * f24 * c2 * c * d + 4 * a * b * f2s * c2 + 4 * a2 * f2s * c * d
           - 2 * a2 * d * f1s * c - a * d2 * f14 * b + b2 * c * f14 * d;
    k[3] = 6 * a2 * b2 + 6 * f24 * c2 * d2 + 2 * b2 * f2s * c2 + 2 * a2 * f2s * d2 - 2 * a2 * d2 * f1s
           + 2 * b2 * c2 * f1s + 8 * a * b * f2s * c * d;
    k[2] = 4 * a * b2 * b + 4 * b2 * f2s * c * d + 4 * f24 * c * d2 * d - a2 * d * c + b * c2 * a
           + 4 * a * b * f2s * d2 - 2 * a * d2 * f1s * b + 2 * b2 * c * f1s * d;
    k[1] = f24 * d2 * d2 + b2 * b2 + 2 * b2 * f2s * d2 - a2 * d2 + b2 * c2;
    k[0] = -a * d2 * b + b2 * c * d;
    double ts[6];
    solve_poly6_cv(k, ts);
    double s_val = 1.0 / f1s + c2 / (a2 + f2s * c2);   // the cost at t = infinity
    double t = 0.0;
    bool have = false;
    for (int i = 0; i < 6; ++i) {
        const double ti = ts[i];
        const double u = c * ti + d, v = a * ti + b;
        const double s = ti * ti / (1.0 + f1s * ti * ti) + u * u / (v * v + f2s * u * u);
        if (s < s_val) { s_val = s; t = ti; have = true; }
    }
    if (!have) return false;
    // closest points to the origin on l1 = (t f1, 1, -t) and l2 = (-f2 (ct+d), at+b, ct+d), mapped back by T R^T
    {
        const double w = t * t * f1s + 1.0;
        const double px = t * t * f1 / w, py = t / w;
        o[0] = e1x * px - e1y * py + x1;
        o[1] = e1y * px + e1x * py + y1;
    }
    {
        const double u = c * t + d, v = a * t + b;
        const double w = f2s * u * u + v * v;
        const double px = f2 * u * u / w, py = -v * u / w;
        o[2] = e2x * px - e2y * py + x2;
        o[3] = e2y * px + e2x * py + y2;
    }
    return true;
}

FEPE_HD void inv3(const double (&A)[9], double (&Ai)[9]) {
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double idet = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
    Ai[0] = c00 * idet; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * idet; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * idet;
    Ai[3] = c01 * idet; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * idet; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * idet;
    Ai[6] = c02 * idet; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * idet; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * idet;
}

// Ground truth of one pair from its scene motion x2 = R x1 + t (Rt row-major 4x4) and intrinsics:
//   gt[0..8] E = [t]x R   gt[9..17] F = K^-T E K^-1   gt[18..21] q_cam  gt[22..24] t_cam  (inverse motion)
//   gt[25..28] q_scene    gt[29..31] t_scene;  Kinv is returned for the normalised virtual points.
FEPE_HD void gt_from_motion(const double (&K)[9], const double (&Rt)[16], double (&gt)[32], double (&Kinv)[9]) {
    double R[9], t[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = Rt[4 * r + c];
        t[r] = Rt[4 * r + 3];
    }
    const double tx[9] = {0.0, -t[2], t[1], t[2], 0.0, -t[0], -t[1], t[0], 0.0};
    double E[9], M[9], Fm[9];
    mat3_mul(tx, R, E);
    inv3(K, Kinv);
    mat3_mul_tn(Kinv, E, M);
    mat3_mul(M, Kinv, Fm);
    for (int i = 0; i < 9; ++i) { gt[i] = E[i]; gt[9 + i] = Fm[i]; }
    double Ri[9], q[4];
    inv3(R, Ri);                                        // np.linalg.inv of the padded 4x4: [R^-1 | -R^-1 t]
    rot_to_quat(Ri, q);
    for (int i = 0; i < 4; ++i) gt[18 + i] = q[i];
    for (int r = 0; r < 3; ++r) gt[22 + r] = -(Ri[3 * r] * t[0] + Ri[3 * r + 1] * t[1] + Ri[3 * r + 2] * t[2]);
    rot_to_quat(R, q);
    for (int i = 0; i < 4; ++i) gt[25 + i] = q[i];
    for (int r = 0; r < 3; ++r) gt[29 + r] = t[r];
}

}  // namespace fepe
