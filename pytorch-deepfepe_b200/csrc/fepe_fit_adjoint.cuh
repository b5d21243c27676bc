// Per-correspondence and per-pair adjoints of the weighted 8-point fit, including the gradient w.r.t. the
// COORDINATES (needed by the reference when `if_learn_offsets` adds a learned offset to the matches,
// deepFEPE/models/DeepFNet.py:373,489-505, or when the keypoint front-end is trained,
// Train_model_pipeline.py:384).  What autograd does there, walking Fit.normalize (:148-179), the constraint
// rows and their L2 normalisation (:203-212), the SVD (:233) and compute_epi_residual (utils_F.py:400-413),
// is restated in closed form:
//
//   x_i = w_i p^_i,  p_i = a (x) b,  a = (x~2, y~2, 1),  b = (x~1, y~1, 1),  n_i = |a||b|
//   xbar_i = -(x_i.f) z - (x_i.z) f + rbar_i f             (eigenvector adjoint, z = (G - lambda)^+ fbar)
//   p^bar_i = w_i xbar_i = alpha z + beta f,   alpha = -w^2 (p^.f),  beta = -w^2 (p^.z) + w rbar
//   pbar_i  = (p^bar - (p^bar.p^) p^) / n
//   abar_j  = sum_k pbar[3j+k] b_k,   bbar_k = sum_j pbar[3j+k] a_j
// and, for the Hartley transform x~ = s (u - c) with c = mean u, s = 1.4142 / mean |u - c|,
//   sbar = (from T in out = T2^T F2 T1) + sum_i x~bar_i.(u_i - c);  cbar = (from T) - s sum_i x~bar_i
//   ubar_i = s x~bar_i + A (u_i - c)/|u_i - c| + B,   A = -sbar s^2 / (1.4142 N),  B = (cbar - A sum_i dir_i)/N.
//
// Templated on the scalar type and __host__ __device__: the kernels instantiate the streaming pieces in fp32
// and the pair-level algebra in fp64; tests/host_shim.cpp instantiates everything in fp64 with g++ and
// tests/test_math_host.py compares the result with fp64 autograd over the same graph (a test of the product code).
#pragma once

#include "fepe_math.cuh"

namespace fepe {

FEPE_HD float inv_sqrt_t(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
FEPE_HD double inv_sqrt_t(double x) { return 1.0 / sqrt(x); }
FEPE_HD float sqrt_t(float x) { return sqrtf(x); }
FEPE_HD double sqrt_t(double x) { return sqrt(x); }
FEPE_HD float abs_t(float x) { return fabsf(x); }
FEPE_HD double abs_t(double x) { return fabs(x); }

// Adjoint of one constraint row.  (x1,y1,x2,y2): Hartley-normalised coordinates; w: weight; rbar: upstream
// gradient of the signed residual r_i = x_i.f; f, z: row-major 3x3.  wbar: gradient of the weight;
// xb = d/d(x~1, y~1, x~2, y~2).
template <typename T>
FEPE_HD void row_adjoint(T x1, T y1, T x2, T y2, T w, T rbar, const T (&f)[9], const T (&z)[9], T& wbar,
                         T (&xb)[4]) {
    const T nb = x1 * x1 + y1 * y1 + T(1);
    const T na = x2 * x2 + y2 * y2 + T(1);
    const T inv = inv_sqrt_t(na * nb);                       // 1 / |p|
    const T fb0 = f[0] * x1 + f[1] * y1 + f[2], fb1 = f[3] * x1 + f[4] * y1 + f[5], fb2 = f[6] * x1 + f[7] * y1 + f[8];
    const T zb0 = z[0] * x1 + z[1] * y1 + z[2], zb1 = z[3] * x1 + z[4] * y1 + z[5], zb2 = z[6] * x1 + z[7] * y1 + z[8];
    const T pf = (x2 * fb0 + y2 * fb1 + fb2) * inv;
    const T pz = (x2 * zb0 + y2 * zb1 + zb2) * inv;
    wbar = T(-2) * w * pz * pf + rbar * pf;
    const T w2 = w * w;
    const T alpha = -w2 * pf;
    const T beta = -w2 * pz + w * rbar;
    const T gamma = (alpha * pz + beta * pf) * inv;          // (p^bar.p^) / n
    const T fa0 = f[0] * x2 + f[3] * y2 + f[6], fa1 = f[1] * x2 + f[4] * y2 + f[7];   // (F0^T a)_k
    const T za0 = z[0] * x2 + z[3] * y2 + z[6], za1 = z[1] * x2 + z[4] * y2 + z[7];
    xb[0] = (alpha * za0 + beta * fa0 - gamma * x1 * na) * inv;
    xb[1] = (alpha * za1 + beta * fa1 - gamma * y1 * na) * inv;
    xb[2] = (alpha * zb0 + beta * fb0 - gamma * x2 * nb) * inv;
    xb[3] = (alpha * zb1 + beta * fb1 - gamma * y2 * nb) * inv;
}

// Adjoint of one clamped symmetric epipolar distance e = min(|dd| (1/(m1+eps) + 1/(m2+eps)), clamp_at)
// (utils_F.py:400-413) in the frame of (u1,v1,u2,v2) and Fo.  g: upstream gradient.  ge (accumulated):
// d/dFo; cb (set): d/d(u1,v1,u2,v2).  torch.clamp(max=c) passes the gradient where e <= c.
template <typename T>
FEPE_HD void epi_adjoint(T u1, T v1, T u2, T v2, const T (&Fo)[9], T clamp_at, T g_up, T (&ge)[9], T (&cb)[4]) {
    const T l10 = u2 * Fo[0] + v2 * Fo[3] + Fo[6];
    const T l11 = u2 * Fo[1] + v2 * Fo[4] + Fo[7];
    const T l12 = u2 * Fo[2] + v2 * Fo[5] + Fo[8];
    const T l20 = Fo[0] * u1 + Fo[1] * v1 + Fo[2];
    const T l21 = Fo[3] * u1 + Fo[4] * v1 + Fo[5];
    const T dd = l10 * u1 + l11 * v1 + l12;
    const T m1 = sqrt_t(l10 * l10 + l11 * l11), m2 = sqrt_t(l20 * l20 + l21 * l21);
    const T i1 = T(1) / (m1 + T(1e-6)), i2 = T(1) / (m2 + T(1e-6));
    const T ad = abs_t(dd);
    const T dist = ad * (i1 + i2);
    const T g = (dist <= clamp_at) ? g_up : T(0);
    const T sg = (dd > T(0)) ? g : ((dd < T(0)) ? -g : T(0));
    const T S12 = sg * (i1 + i2);
    const T tiny = T(1e-30);
    const T a1 = -g * ad * i1 * i1 / (m1 > tiny ? m1 : tiny);   // d(1/(m1+eps)) = -i1^2 dm1, dm1 = l1.dl1/m1
    const T a2 = -g * ad * i2 * i2 / (m2 > tiny ? m2 : tiny);
    // d dd / dF_jk = x2_j x1_k ; d m1 / dF_jk = l1_k x2_j / m1 (k<2) ; d m2 / dF_jk = l2_j x1_k / m2 (j<2)
    const T uk0 = S12 * u1 + a1 * l10, uk1 = S12 * v1 + a1 * l11, uk2 = S12;   // times x2_j
    const T vj0 = a2 * l20, vj1 = a2 * l21;                                    // times x1_k
    ge[0] += u2 * uk0 + vj0 * u1; ge[1] += u2 * uk1 + vj0 * v1; ge[2] += u2 * uk2 + vj0;
    ge[3] += v2 * uk0 + vj1 * u1; ge[4] += v2 * uk1 + vj1 * v1; ge[5] += v2 * uk2 + vj1;
    ge[6] += uk0;                 ge[7] += uk1;                 ge[8] += uk2;
    // d dd / dx1_k = l1_k, d dd / dx2_j = l2_j ; m1 depends on x2 only, m2 on x1 only
    cb[0] = S12 * l10 + a2 * (l20 * Fo[0] + l21 * Fo[3]);
    cb[1] = S12 * l11 + a2 * (l20 * Fo[1] + l21 * Fo[4]);
    cb[2] = S12 * l20 + a1 * (l10 * Fo[0] + l11 * Fo[1]);
    cb[3] = S12 * l21 + a1 * (l10 * Fo[3] + l11 * Fo[4]);
}

// Pair-level sums the coordinate gradient needs from the streaming pass, per image (index 0: image 1).
struct NormAdjointSums {
    double sx[2], sy[2];     // sum_i x~bar_i, sum_i y~bar_i
    double sd[2];            // sum_i x~bar_i (u_i - cx) + y~bar_i (v_i - cy)
    double dx[2], dy[2];     // sum_i (u_i - cx)/d_i, sum_i (v_i - cy)/d_i
};
// ubar_i = s x~bar_i + A (u_i - c)/d_i + (Bx, By)
struct NormAdjointCoef {
    double A[2], Bx[2], By[2];
};

// ob: gradient w.r.t. out = T2^T F2 T1 (row-major 3x3); F2: the rank-2 matrix in the normalised frame;
// (s, cx, cy)[2]: Hartley transforms of image 1 and 2; N: correspondences.
FEPE_HD void norm_adjoint(const double (&ob)[9], const double (&F2)[9], const double (&s)[2], const double (&cx)[2],
                          const double (&cy)[2], const NormAdjointSums& S, int N, NormAdjointCoef& out) {
    // X = T2 ob, T1bar = F2^T X
    const double t2x = -s[1] * cx[1], t2y = -s[1] * cy[1], t1x = -s[0] * cx[0], t1y = -s[0] * cy[0];
    double X[9], T1b[9], Y[9], T2b[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        X[k] = s[1] * ob[k] + t2x * ob[6 + k];
        X[3 + k] = s[1] * ob[3 + k] + t2y * ob[6 + k];
        X[6 + k] = ob[6 + k];
    }
    mat3_mul_tn(F2, X, T1b);
    // Y = T1 ob^T, T2bar = F2 Y
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Y[c] = s[0] * ob[3 * c] + t1x * ob[3 * c + 2];
        Y[3 + c] = s[0] * ob[3 * c + 1] + t1y * ob[3 * c + 2];
        Y[6 + c] = ob[3 * c + 2];
    }
    mat3_mul(F2, Y, T2b);
    const double* Tb[2] = {T1b, T2b};
    const double invN = 1.0 / static_cast<double>(N);
#pragma unroll
    for (int im = 0; im < 2; ++im) {
        const double* tb = Tb[im];
        const double sbar = tb[0] + tb[4] - cx[im] * tb[2] - cy[im] * tb[5] + S.sd[im];
        double cxb = -s[im] * (tb[2] + S.sx[im]);
        double cyb = -s[im] * (tb[5] + S.sy[im]);
        const double A = -sbar * s[im] * s[im] * (1.0 / 1.4142) * invN;
        cxb -= A * S.dx[im];
        cyb -= A * S.dy[im];
        out.A[im] = A;
        out.Bx[im] = cxb * invN;
        out.By[im] = cyb * invN;
    }
}

}  // namespace fepe
