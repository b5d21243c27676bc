// Small dense linear algebra for the fused weighted 8-point kernels.
//
// Everything here is straight-line, fully unrolled, register-resident code that every lane of a
// warp executes redundantly (the inputs are warp-uniform), so there is no divergence and no
// broadcast step afterwards.  The functions are __host__ __device__ so that tests/ can compile
// the very same code with g++ and check it against LAPACK before it ever runs on a GPU
// (tests/test_math_host.py); that is a test of the product code, not a CPU fallback.
//
// What the reference does at these points (deepFEPE/models/DeepFNet.py):
//   :233  torch.svd(X[b]) -> V[:,-1]          here: smallest eigenpair of G = X^T X   (eig9_smallest)
//   :236  torch.svd(F) and S*[1,1,0]          here: one-sided Jacobi SVD of a 3x3     (svd3)
// and deepFEPE/dsac_tools/utils_F.py:480 torch.svd(E) for the pose head.
#pragma once

#if defined(__CUDACC__)
#define FEPE_HD __host__ __device__ __forceinline__
#else
#define FEPE_HD inline
#endif

#include <math.h>

namespace fepe {

// ---------------------------------------------------------------------------------------------
// Gram matrix storage.  A constraint row is p = a (x) b with a = (x2,y2,1), b = (x1,y1,1), so
// p p^T = (a a^T) (x) (b b^T) has only 6 x 6 = 36 distinct entries.  g36[u*6+v] holds
// sum_i s_i * mA_u * mB_v with the monomial order m = [c0c0, c0c1, c0c2, c1c1, c1c2, c2c2].
// G(r,c) with r = 3j+k, c = 3l+m is g36[sym6(j,l)*6 + sym6(k,m)].
// ---------------------------------------------------------------------------------------------
FEPE_HD constexpr int sym6(int i, int j) {
    return (i <= j) ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : (j == 0 ? i : (j == 1 ? 2 + i : 5));
}
FEPE_HD constexpr int g36_index(int r, int c) { return sym6(r / 3, c / 3) * 6 + sym6(r % 3, c % 3); }

// LDL^T of M = G - mu*I (unit lower L, reciprocal pivots rd).  Returns the number of negative
// pivots = number of eigenvalues of G below mu (Sylvester inertia).
FEPE_HD int ldl9(const double* __restrict__ g36, double mu, double tiny, double (&L)[36], double (&rd)[9]) {
    // L is stored strictly-lower packed: L(i,j), i>j, at i*(i-1)/2 + j
    int nneg = 0;
    double d[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double v[9];
        double dj = g36[g36_index(j, j)] - mu;
#pragma unroll
        for (int k = 0; k < j; ++k) {
            v[k] = L[j * (j - 1) / 2 + k] * d[k];
            dj -= L[j * (j - 1) / 2 + k] * v[k];
        }
        if (dj < 0.0) ++nneg;
        if (fabs(dj) < tiny) dj = (dj < 0.0) ? -tiny : tiny;
        d[j] = dj;
        const double r = 1.0 / dj;
        rd[j] = r;
#pragma unroll
        for (int i = j + 1; i < 9; ++i) {
            double lij = g36[g36_index(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) lij -= L[i * (i - 1) / 2 + k] * v[k];
            L[i * (i - 1) / 2 + j] = lij * r;
        }
    }
    return nneg;
}

// x <- M^{-1} x using the factorisation above.
FEPE_HD void ldl9_solve(const double (&L)[36], const double (&rd)[9], double (&x)[9]) {
#pragma unroll
    for (int i = 1; i < 9; ++i) {
#pragma unroll
        for (int k = 0; k < i; ++k) x[i] -= L[i * (i - 1) / 2 + k] * x[k];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] *= rd[i];
#pragma unroll
    for (int i = 7; i >= 0; --i) {
#pragma unroll
        for (int k = i + 1; k < 9; ++k) x[i] -= L[k * (k - 1) / 2 + i] * x[k];
    }
}

FEPE_HD void gram9_matvec(const double* __restrict__ g36, const double (&x)[9], double (&y)[9]) {
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) s += g36[g36_index(r, c)] * x[c];
        y[r] = s;
    }
}

// Smallest eigenpair of the symmetric positive semi-definite 9x9 Gram matrix.
//
// Inverse iteration with shifts that approach lambda_min from BELOW: mu = rho - |G x - rho x|
// (some eigenvalue lies within the residual norm of the Rayleigh quotient rho).  While the shift is
// below lambda_min the iteration can only converge to the smallest eigenvector; the inertia count
// of the LDL^T factorisation detects a shift that overshot and bisects back.  Once x is dominated by
// the two lowest eigenvectors the contraction per step is <= 0.17 whatever the gap, so nearly
// repeated lambda_8 ~ lambda_9 (degenerate scenes) cost no extra iterations.
// Output: unit f, sign fixed so that the entry of largest magnitude is positive.  Returns the
// number of factorisations used.
FEPE_HD int eig9_smallest(const double* __restrict__ g36, double (&f)[9], double& lambda) {
    double tr = 0.0;
#pragma unroll
    for (int r = 0; r < 9; ++r) tr += g36[g36_index(r, r)];
    if (!(tr > 0.0) || !(tr < 1e300)) {   // empty / all-zero-weight / non-finite input
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
        lambda = 0.0;
        return 0;
    }
    const double tiny = 1e-18 * tr;
    // fixed generic start vector (unit norm); any vector not orthogonal to the answer works
    double x[9] = {0.3713906763541037, -0.2228344058124622, 0.4456688116249244, 0.1485562705416415,
                   -0.5199469468957452, 0.2971125410832830, -0.0742781352708207, 0.3342516087186933,
                   0.3565350492999395};
    double mu = -1e-14 * tr;   // G + 1e-14 tr I is numerically positive definite
    double lo = mu;
    double rho = 0.0;
    double r_prev = -1.0;      // residual after the previous solve (none yet)
    int it = 0;
    for (; it < 24; ++it) {
        double L[36], rd[9];
        const int nneg = ldl9(g36, mu, tiny, L, rd);
        if (nneg > 0) {         // overshot lambda_min: go back half way to the last safe shift
            mu = 0.5 * (mu + lo);
            continue;
        }
        lo = mu;
        // The first factorisation (shift ~0) is used for two solves: cheap extra contraction.
        double r = 0.0;
        for (int rep = (it == 0) ? 2 : 1; rep > 0; --rep) {
            double y[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) y[i] = x[i];
            ldl9_solve(L, rd, y);
            double nrm2 = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) nrm2 += y[i] * y[i];
            const double inv = 1.0 / sqrt(nrm2);
            // With x' = y/|y| and (G - mu I) y = x:  G x' = mu x' + x/|y|, hence
            //   rho = mu + (x'.x)/|y|   and   G x' - rho x' = (x - (x'.x) x')/|y|   (no mat-vec needed)
            double c = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { y[i] *= inv; c += y[i] * x[i]; }
            double e2 = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { const double e = x[i] - c * y[i]; e2 += e * e; x[i] = y[i]; }
            rho = mu + c * inv;
            r = sqrt(e2) * inv;
        }
        // Stop when the eigenVECTOR has converged: its error is ~ r / (lambda_8 - lambda_9) and the
        // gap is estimated from the observed contraction q = r/r_prev = (lambda_9-mu)/(lambda_8-mu).
        if (r <= 1e-17 * tr) { ++it; break; }
        if (r_prev >= 0.0) {
            const double q = r / r_prev;
            if (q < 1.0) {
                const double gap = (rho - mu) * (1.0 / q - 1.0);
                if (r <= 1e-9 * gap) { ++it; break; }
            }
            if (q >= 0.5 && r_prev <= 1e-9 * tr) { ++it; break; }   // stagnated at the rounding floor
        }
        r_prev = r;
        const double cand = rho - r * 1.0000001 - 4e-16 * tr;
        if (cand > mu) mu = cand;
    }
    // canonical sign
    int imax = 0;
    double amax = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (fabs(x[i]) > amax) { amax = fabs(x[i]); imax = i; }
    }
    double sgn = 1.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (i == imax && x[i] < 0.0) sgn = -1.0;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = sgn * x[i];
    lambda = rho;
    return it;
}

// z = (G - lambda I)^+ rhs restricted to the complement of f (used by the backward pass):
// solve (G - lambda I + tau f f^T) y = rhs - f (f.rhs), then z = y - f (f.y).
FEPE_HD void eig9_pinv_apply(const double* __restrict__ g36, const double (&f)[9], double lambda,
                             const double (&rhs)[9], double (&z)[9]) {
    double tr = 0.0;
#pragma unroll
    for (int r = 0; r < 9; ++r) tr += g36[g36_index(r, r)];
    const double tau = tr / 9.0 + 1e-300;
    // dense symmetric M, Cholesky-free LDL^T on the full 9x9 (M is positive definite up to rounding)
    double M[45];   // lower packed incl. diagonal: (i,j), i>=j at i*(i+1)/2 + j
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j)
            M[i * (i + 1) / 2 + j] = g36[g36_index(i, j)] + tau * f[i] * f[j] - (i == j ? lambda : 0.0);
    }
    double fr = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) fr += f[i] * rhs[i];
    double y[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) y[i] = rhs[i] - f[i] * fr;
    // in-place LDL^T: M(i,j), i>j becomes L(i,j); diagonal becomes d_j
    const double tiny = 1e-18 * tr + 1e-300;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double v[9];
        double dj = M[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) {
            v[k] = M[j * (j + 1) / 2 + k] * M[k * (k + 1) / 2 + k];
            dj -= M[j * (j + 1) / 2 + k] * v[k];
        }
        if (fabs(dj) < tiny) dj = (dj < 0.0) ? -tiny : tiny;
        M[j * (j + 1) / 2 + j] = dj;
        const double r = 1.0 / dj;
#pragma unroll
        for (int i = j + 1; i < 9; ++i) {
            double lij = M[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) lij -= M[i * (i + 1) / 2 + k] * v[k];
            M[i * (i + 1) / 2 + j] = lij * r;
        }
    }
#pragma unroll
    for (int i = 1; i < 9; ++i) {
#pragma unroll
        for (int k = 0; k < i; ++k) y[i] -= M[i * (i + 1) / 2 + k] * y[k];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) y[i] /= M[i * (i + 1) / 2 + i];
#pragma unroll
    for (int i = 7; i >= 0; --i) {
#pragma unroll
        for (int k = i + 1; k < 9; ++k) y[i] -= M[k * (k + 1) / 2 + i] * y[k];
    }
    double fy = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) fy += f[i] * y[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) z[i] = y[i] - f[i] * fy;
}

// ---------------------------------------------------------------------------------------------
// 3x3 SVD, A = U diag(S) V^T, one-sided (Hestenes) Jacobi in fp64, singular values sorted
// descending, det(U) = det(V) = +1 is NOT enforced (signs follow the rotations); u3 is rebuilt as
// +-u1 x u2 when sigma_3 is negligible so rank-2 inputs (essential matrices) are handled.
// Row-major 3x3 arrays.
// ---------------------------------------------------------------------------------------------
FEPE_HD void svd3(const double (&A)[9], double (&U)[9], double (&S)[3], double (&V)[9]) {
    double W[9], Q[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { W[i] = A[i]; Q[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            const double alpha = W[p] * W[p] + W[3 + p] * W[3 + p] + W[6 + p] * W[6 + p];
            const double beta = W[q] * W[q] + W[3 + q] * W[3 + q] + W[6 + q] * W[6 + q];
            const double gamma = W[p] * W[q] + W[3 + p] * W[3 + q] + W[6 + p] * W[6 + q];
            const double lim = 1e-32 * alpha * beta;
            if (gamma * gamma > lim && gamma != 0.0) {
                off += gamma * gamma / (alpha * beta + 1e-300);
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = ((zeta >= 0.0) ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t);
                const double s = c * t;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double wp = W[3 * r + p], wq = W[3 * r + q];
                    W[3 * r + p] = c * wp - s * wq;
                    W[3 * r + q] = s * wp + c * wq;
                    const double vp = Q[3 * r + p], vq = Q[3 * r + q];
                    Q[3 * r + p] = c * vp - s * vq;
                    Q[3 * r + q] = s * vp + c * vq;
                }
            }
        }
        if (off < 1e-30) break;
    }
    double n[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) n[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    // sort columns by descending norm (3-element network), applied to W, Q and n
#define FEPE_SWAPCOL(a, b)                                                           \
    if (n[a] < n[b]) {                                                               \
        double tmp = n[a]; n[a] = n[b]; n[b] = tmp;                                  \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {                              \
            tmp = W[3 * r + a]; W[3 * r + a] = W[3 * r + b]; W[3 * r + b] = tmp;     \
            tmp = Q[3 * r + a]; Q[3 * r + a] = Q[3 * r + b]; Q[3 * r + b] = tmp;     \
        }                                                                            \
    }
    FEPE_SWAPCOL(0, 1)
    FEPE_SWAPCOL(1, 2)
    FEPE_SWAPCOL(0, 1)
#undef FEPE_SWAPCOL
#pragma unroll
    for (int j = 0; j < 3; ++j) S[j] = n[j];
#pragma unroll
    for (int i = 0; i < 9; ++i) V[i] = Q[i];
    const double i0 = 1.0 / (n[0] + 1e-300), i1 = 1.0 / (n[1] + 1e-300);
#pragma unroll
    for (int r = 0; r < 3; ++r) { U[3 * r] = W[3 * r] * i0; U[3 * r + 1] = W[3 * r + 1] * i1; }
    // third left vector: u1 x u2 with the sign of w3 (exact when sigma3 > 0, well defined when it is 0)
    double c0 = U[3] * U[7] - U[6] * U[4];
    double c1 = U[6] * U[1] - U[0] * U[7];
    double c2 = U[0] * U[4] - U[3] * U[1];
    const double cn = 1.0 / (sqrt(c0 * c0 + c1 * c1 + c2 * c2) + 1e-300);
    c0 *= cn; c1 *= cn; c2 *= cn;
    const double dotw = c0 * W[2] + c1 * W[5] + c2 * W[8];
    const double sg = (dotw < 0.0) ? -1.0 : 1.0;
    U[2] = sg * c0; U[5] = sg * c1; U[8] = sg * c2;
}

FEPE_HD void mat3_mul(const double (&A)[9], const double (&B)[9], double (&C)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
FEPE_HD void mat3_mul_tn(const double (&A)[9], const double (&B)[9], double (&C)[9]) {   // A^T B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
FEPE_HD void mat3_mul_nt(const double (&A)[9], const double (&B)[9], double (&C)[9]) {   // A B^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
FEPE_HD double det3(const double (&A)[9]) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
           A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// Rank-2 projection of F0 = reshape(f): F0 - sigma3 u3 v3^T  (DeepFNet.py:236-237, S*[1,1,0]).
// Since sigma3 u3 = F0 v3 this is F0 (I - v3 v3^T); only V is needed.
FEPE_HD void rank2_project(const double (&F0)[9], double (&F2)[9], double (&U)[9], double (&S)[3], double (&V)[9]) {
    svd3(F0, U, S, V);
    const double v0 = V[2], v1 = V[5], v2 = V[8];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double w = F0[3 * r] * v0 + F0[3 * r + 1] * v1 + F0[3 * r + 2] * v2;
        F2[3 * r] = F0[3 * r] - w * v0;
        F2[3 * r + 1] = F0[3 * r + 1] - w * v1;
        F2[3 * r + 2] = F0[3 * r + 2] - w * v2;
    }
}

}  // namespace fepe
