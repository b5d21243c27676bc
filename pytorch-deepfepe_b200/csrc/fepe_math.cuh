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

// fp64 reciprocal / reciprocal square root for the latency-critical solvers: hardware seed
// (MUFU.RCP64H / RSQ64H, ~2^-23) plus two Newton steps, no special-case handling (callers guarantee
// finite non-zero arguments).  About half the latency of the IEEE division sequence.
FEPE_HD double fast_rcp(double d) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / d;
#endif
}
FEPE_HD double fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double h = 0.5 * x;
    double e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-h * r, r, 0.5);
    return fma(r, e, r);
#else
    return 1.0 / sqrt(x);
#endif
}

// sqrt for non-negative arguments through the fast reciprocal square root (0 -> 0).
FEPE_HD double fast_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return (x > 0.0) ? x * fast_rsqrt(x) : 0.0;
#else
    return sqrt(x);
#endif
}

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

// View of one Gram matrix inside an [entry][problem] array (entry e of this problem at p[e * STRIDE]): lets
// one LANE own one eigenproblem with bank-conflict-free shared-memory reads (fepe_fit_split.cu).
template <int STRIDE>
struct StridedG36 {
    const double* p;
    FEPE_HD double operator[](int e) const { return p[e * STRIDE]; }
};

// LDL^T of M = G - mu*I (unit lower L, reciprocal pivots rd).  Returns the number of negative
// pivots = number of eigenvalues of G below mu (Sylvester inertia).
// Right-looking (outer-product) form: once column j is scaled every trailing update is an
// independent FMA, so the dependent chain per column is just reciprocal -> scale -> one FMA.
// G36 is anything indexable by the g36 entry number: a plain pointer, or a strided view (StridedG36).
template <class G36>
FEPE_HD int ldl9(const G36& g36, double mu, double tiny, double (&A)[45]) {
    // A: lower triangle incl. diagonal, (i,j), i>=j at i*(i+1)/2 + j.  On exit the strictly lower part
    // holds L and the diagonal holds the RECIPROCAL pivots.
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) A[i * (i + 1) / 2 + j] = g36[g36_index(i, j)] - ((i == j) ? mu : 0.0);
    }
    int nneg = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double dj = A[j * (j + 1) / 2 + j];
        if (dj < 0.0) ++nneg;
        if (fabs(dj) < tiny) dj = (dj < 0.0) ? -tiny : tiny;
        const double r = fast_rcp(dj);
        A[j * (j + 1) / 2 + j] = r;
        double t[9];
#pragma unroll
        for (int i = j + 1; i < 9; ++i) t[i] = A[i * (i + 1) / 2 + j] * r;      // l_ij
#pragma unroll
        for (int i = j + 1; i < 9; ++i) {
#pragma unroll
            for (int k = j + 1; k <= i; ++k)                                     // a_kj still unscaled here
                A[i * (i + 1) / 2 + k] = fma(-t[i], A[k * (k + 1) / 2 + j], A[i * (i + 1) / 2 + k]);
        }
#pragma unroll
        for (int i = j + 1; i < 9; ++i) A[i * (i + 1) / 2 + j] = t[i];
    }
    return nneg;
}

// x <- M^{-1} x using the factorisation above (column-oriented substitutions: each finished
// unknown is pushed into all later rows at once, keeping the dependent chain at one FMA per step).
FEPE_HD void ldl9_solve(const double (&A)[45], double (&x)[9]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int i = k + 1; i < 9; ++i) x[i] = fma(-A[i * (i + 1) / 2 + k], x[k], x[i]);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] *= A[i * (i + 1) / 2 + i];
#pragma unroll
    for (int k = 8; k >= 1; --k) {
#pragma unroll
        for (int i = 0; i < k; ++i) x[i] = fma(-A[k * (k + 1) / 2 + i], x[k], x[i]);
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
template <class G36>
FEPE_HD int eig9_smallest(const G36& g36, double (&f)[9], double& lambda) {
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
        double A[45];
        const int nneg = ldl9(g36, mu, tiny, A);
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
            ldl9_solve(A, y);
            double nrm2 = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) nrm2 += y[i] * y[i];
            const double inv = fast_rsqrt(nrm2);
            // With x' = y/|y| and (G - mu I) y = x:  G x' = mu x' + x/|y|, hence
            //   rho = mu + (x'.x)/|y|   and   G x' - rho x' = (x - (x'.x) x')/|y|   (no mat-vec needed)
            double c = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { y[i] *= inv; c += y[i] * x[i]; }
            double e2 = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { const double e = x[i] - c * y[i]; e2 += e * e; x[i] = y[i]; }
            rho = mu + c * inv;
            r = fast_sqrt(e2) * inv;
        }
        // Stop when the eigenVECTOR has converged: its error is ~ r / (lambda_8 - lambda_9) and the
        // gap is estimated from the observed contraction q = r/r_prev = (lambda_9-mu)/(lambda_8-mu).
        if (r <= 1e-17 * tr) { ++it; break; }
        if (r_prev >= 0.0) {
            const double q = (r_prev > 0.0) ? r * fast_rcp(r_prev) : 2.0;
            if (q < 1.0) {
                // r <= 1e-8 gap  with  gap = (rho - mu) (1/q - 1), written without the division
                if (r * q <= 1e-8 * (rho - mu) * (1.0 - q)) { ++it; break; }
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

// ---------------------------------------------------------------------------------------------
// Multi-shift variant for a warp that owns ONE eigenproblem: the 32 lanes would otherwise run the
// serial iteration above redundantly, so each lane factors G - mu_l I with its OWN shift instead.
// The inertia counts bracket lambda_min 32-fold per round (guaranteed, no heuristics), every lane
// also does its two inverse-iteration solves, and the lane with the largest shift still below
// lambda_min supplies the next iterate.  Three rounds typically replace six sequential
// factorisations.  The scalar pieces below are shared by the device driver (fepe_fit.cuh, uses
// ballot/shuffle) and by the host emulation in tests/host_shim.cpp.
// ---------------------------------------------------------------------------------------------
struct Eig9Bracket {
    double tr;        // trace(G)
    double lo;        // largest shift known to be below lambda_min (inertia count 0)
    double hi;        // upper limit for lambda_min (first failing shift, Rayleigh quotient, min diagonal)
    double r_prev;    // residual of the previous round's best lane (<0: none)
    double lo_heur;   // heuristic (unverified) lower end for lanes 1..31: max(lo, rho - r)
    float geo_step;   // log2 step of the geometric Sturm probe (tri9_probe_begin)
    int round;
};

FEPE_HD bool eig9_bracket_init(const double* __restrict__ g36, Eig9Bracket& b) {
    double tr = 0.0, dmin = 1e300;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const double d = g36[g36_index(r, r)];
        tr += d;
        dmin = d < dmin ? d : dmin;
    }
    b.tr = tr;
    b.lo = -1e-14 * tr;
    b.hi = dmin;                 // e_i^T G e_i >= lambda_min
    b.r_prev = -1.0;
    b.lo_heur = b.lo;
    b.round = 0;
    return (tr > 0.0) && (tr < 1e300);
}

// Shift of lane `lane` (0..nlanes-1) for the current round; lane 0 always carries the known-safe shift.
FEPE_HD double eig9_lane_shift(const Eig9Bracket& b, int lane, int nlanes = 32) {
    if (lane == 0) return b.lo;
    if (b.round == 0) {
        // geometric ladder from 1e-13 tr up to the upper limit: lambda_min can sit anywhere on 13 decades
        const double a = 1e-13 * b.tr;
        const double top = (b.hi > 2.0 * a) ? b.hi : 2.0 * a;
#if defined(__CUDA_ARCH__)
        // fp32 transcendentals: the ladder only has to be ascending, not accurate
        return a * static_cast<double>(exp2f(log2f(static_cast<float>(top * fast_rcp(a))) *
                                             (static_cast<float>(lane - 1) / static_cast<float>(nlanes - 2))));
#else
        return a * exp2(log2(top / a) * (static_cast<double>(lane - 1) / static_cast<double>(nlanes - 2)));
#endif
    }
    const double lo = (b.lo_heur > 0.0) ? b.lo_heur : 0.0;
    return lo + (b.hi - lo) * (static_cast<double>(lane - 1) * (1.0 / static_cast<double>(nlanes - 1))) * 0.999;
}

// One lane's work for a round: factor at `mu`, `nsolve` inverse-iteration solves starting from x.
FEPE_HD void eig9_lane_round(const double* __restrict__ g36, double mu, double tiny, int nsolve, double (&x)[9],
                             int& nneg, double& rho, double& r, double& contraction) {
    double A[45];
    nneg = ldl9(g36, mu, tiny, A);
    rho = 0.0;
    r = 0.0;
    contraction = 1.0;     // r_last / r_previous under THIS shift = (lambda_9 - mu) / (lambda_8 - mu)
    for (int rep = 0; rep < nsolve; ++rep) {
        double y[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) y[i] = x[i];
        ldl9_solve(A, y);
        double nrm2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) nrm2 += y[i] * y[i];
        const double inv = fast_rsqrt(nrm2);
        double c = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { y[i] *= inv; c += y[i] * x[i]; }
        double e2 = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { const double e = x[i] - c * y[i]; e2 += e * e; x[i] = y[i]; }
        rho = mu + c * inv;
        const double r_new = fast_sqrt(e2) * inv;
        if (rep > 0) contraction = r_new * fast_rcp(r + 1e-300);
        r = r_new;
    }
}

// Fold the round's outcome into the bracket.  mu_best / rho / r come from the best lane (largest
// shift with inertia 0), mu_fail is the smallest failing shift (or <0 when every lane passed).
// Returns true when the eigenvector has converged (same rule as eig9_smallest).
FEPE_HD bool eig9_bracket_update(Eig9Bracket& b, double mu_best, double mu_fail, double rho, double r, double c) {
    bool done = false;
    if (r <= 1e-17 * b.tr) done = true;
    // eigenvector error ~ r / (lambda_8 - lambda_9); the gap follows from the contraction c the best lane
    // observed between its two solves at the same shift: gap = (rho - mu)(1/c - 1).  Stop at r <= 1e-8 gap.
    if (c < 1.0 && r * c <= 1e-8 * (rho - mu_best) * (1.0 - c)) done = true;
    if (c >= 0.5 && r <= 1e-9 * b.tr) done = true;                 // stagnated at the rounding floor
    b.r_prev = r;
    b.lo = mu_best;
    double hi = rho;                                  // Rayleigh quotient >= lambda_min
    if (mu_fail >= 0.0 && mu_fail < hi) hi = mu_fail;
    // rho - r is (heuristically) a lower bound: start the next ladder there when it beats the safe shift
    const double cand = rho - r * 1.0000001 - 4e-16 * b.tr;
    if (cand > b.lo && cand < hi) {
        // keep lane 0 on the verified shift; lanes 1.. spread over [cand, hi)
        b.lo_heur = cand;
    } else {
        b.lo_heur = b.lo;
    }
    b.hi = hi;
    b.round += 1;
    return done;
}

// ---------------------------------------------------------------------------------------------
// Tridiagonal form of the eigenproblem.  Every shift of the (multi-shift) inverse iteration above costs a
// 9x9 LDL^T factorisation -- 165 dependent-ish FMAs and 9 reciprocals -- and the drivers need 3..6 of them per
// pair.  One Householder reduction G = Q T Q^T (450 FMAs, once per pair) makes every later shift a
// TRIDIAGONAL factorisation: 8 multipliers, 9 pivots, ~30 operations; the inertia count, the two solves per
// shift and the stop rule are unchanged, and the eigenvector of T is carried back with the seven reflectors.
// Backward stable like the shifted LDL^T it replaces: eigenvector error ~ eps |G| / (lambda_8 - lambda_9).
//   ta[9], tb[8]    diagonal / sub-diagonal of T
//   hv[28], htau[7] reflector k acts on indices k+1..8: H_k = I - tau_k v_k v_k^T, v_k = (1, hv[off_k..]),
//                   off_k = k(15-k)/2, 7-k stored entries (LAPACK dsytd2 convention, lower triangle)
// ---------------------------------------------------------------------------------------------
FEPE_HD constexpr int tri9_off(int k) { return k * (15 - k) / 2; }

template <class G36>
FEPE_HD void tridiag9(const G36& g36, double (&ta)[9], double (&tb)[8], double (&hv)[28], double (&htau)[7]) {
    double A[45];      // lower triangle, (i,j), i >= j at i(i+1)/2 + j
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) A[i * (i + 1) / 2 + j] = g36[g36_index(i, j)];
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        // every sum below is split over two accumulators: the reduction is one dependent chain per reflector, and a
        // chain of m FMAs costs m x 8 cycles on a warp that has nothing else to issue
        const double alpha = A[(k + 1) * (k + 2) / 2 + k];
        double xa = 0.0, xb = 0.0;
#pragma unroll
        for (int i = k + 2; i < 9; ++i) {
            const double e = A[i * (i + 1) / 2 + k];
            if ((i - k) & 1) xa = fma(e, e, xa); else xb = fma(e, e, xb);
        }
        const double xn2 = xa + xb;
        const bool live = xn2 > 0.0;
        const double nrm = fast_sqrt(fma(alpha, alpha, xn2));
        const double beta = live ? ((alpha >= 0.0) ? -nrm : nrm) : alpha;
        const double tau = live ? (beta - alpha) * fast_rcp(beta) : 0.0;
        const double sc = live ? fast_rcp(alpha - beta) : 0.0;
        double v[9], w[9];
        v[k + 1] = 1.0;
#pragma unroll
        for (int i = k + 2; i < 9; ++i) v[i] = A[i * (i + 1) / 2 + k] * sc;
        ta[k] = A[k * (k + 1) / 2 + k];
        tb[k] = beta;
        htau[k] = tau;
#pragma unroll
        for (int i = k + 2; i < 9; ++i) hv[tri9_off(k) + i - (k + 2)] = v[i];
        // w = tau A22 v - (tau/2)(tau v^T A22 v) v;  A22 -= v w^T + w v^T
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int i = k + 1; i < 9; ++i) {
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
            for (int j = k + 1; j < 9; ++j) {
                const double e = (i >= j) ? A[i * (i + 1) / 2 + j] : A[j * (j + 1) / 2 + i];
                if ((j - k) & 1) acc0 = fma(e, v[j], acc0); else acc1 = fma(e, v[j], acc1);
            }
            w[i] = tau * (acc0 + acc1);
            if ((i - k) & 1) pa = fma(w[i], v[i], pa); else pb = fma(w[i], v[i], pb);
        }
        const double h = 0.5 * tau * (pa + pb);
#pragma unroll
        for (int i = k + 1; i < 9; ++i) w[i] = fma(-h, v[i], w[i]);
#pragma unroll
        for (int i = k + 1; i < 9; ++i) {
#pragma unroll
            for (int j = k + 1; j <= i; ++j)
                A[i * (i + 1) / 2 + j] -= fma(v[i], w[j], w[i] * v[j]);
        }
    }
    ta[7] = A[7 * 8 / 2 + 7];
    tb[7] = A[8 * 9 / 2 + 7];
    ta[8] = A[8 * 9 / 2 + 8];
}

// x <- Q x = H_0 H_1 ... H_6 x: an eigenvector of T becomes the eigenvector of G.
FEPE_HD void tridiag9_back(const double* __restrict__ hv, const double* __restrict__ htau, double (&x)[9]) {
#pragma unroll
    for (int k = 6; k >= 0; --k) {
        double s0 = x[k + 1], s1 = 0.0;
#pragma unroll
        for (int i = k + 2; i < 9; ++i) {
            if ((i - k) & 1) s1 = fma(hv[tri9_off(k) + i - (k + 2)], x[i], s1);
            else s0 = fma(hv[tri9_off(k) + i - (k + 2)], x[i], s0);
        }
        const double s = (s0 + s1) * htau[k];
        x[k + 1] -= s;
#pragma unroll
        for (int i = k + 2; i < 9; ++i) x[i] = fma(-s, hv[tri9_off(k) + i - (k + 2)], x[i]);
    }
}

// sum_i a_i b_i over 9 entries as three chains of three (24 + 16 cycles instead of 72)
FEPE_HD double dot9(const double (&a)[9], const double (&b)[9]) {
    const double s0 = fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]));
    const double s1 = fma(a[5], b[5], fma(a[4], b[4], a[3] * b[3]));
    const double s2 = fma(a[8], b[8], fma(a[7], b[7], a[6] * b[6]));
    return (s0 + s1) + s2;
}

// Scale T to unit trace (the minors' recurrence then stays near 1); returns trace(G), or 0 for an empty / non-finite G.
FEPE_HD double tri9_normalise(double (&ta)[9], double (&tb)[8]) {
    double tr = 0.0;
#pragma unroll
    for (int r = 0; r < 9; ++r) tr += ta[r];
    if (!(tr > 0.0) || !(tr < 1e300)) return 0.0;
    const double inv = fast_rcp(tr);
#pragma unroll
    for (int r = 0; r < 9; ++r) ta[r] *= inv;
#pragma unroll
    for (int r = 0; r < 8; ++r) tb[r] *= inv;
    return tr;
}

// eig9_bracket_init for the tridiagonal form (trace and smallest diagonal entry of T bound lambda_min as G's do).
FEPE_HD bool tri9_bracket_init(const double (&ta)[9], Eig9Bracket& b) {
    double tr = 0.0, dmin = 1e300;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        tr += ta[r];
        dmin = ta[r] < dmin ? ta[r] : dmin;
    }
    b.tr = tr;
    b.lo = -1e-14 * tr;
    b.hi = dmin;
    b.r_prev = -1.0;
    b.lo_heur = b.lo;
    b.round = 0;
    return (tr > 0.0) && (tr < 1e300);
}

// Number of eigenvalues of T below mu from the three-term recurrence of the leading principal minors
// p_k = (a_k - mu) p_{k-1} - b_{k-1}^2 p_{k-2}: nine dependent FMAs, no division.  tb2 = tb^2.  T is expected scaled to
// unit trace (|p_k| stays within a few binades of 1 unless eigenvalues cluster at mu; nowhere near the fp64 range).
FEPE_HD int tri9_sturm_count(const double (&ta)[9], const double (&tb2)[8], double mu) {
    // branch free: the count is the number of sign-bit changes along 1, p_1 .. p_9 (an exact zero reads as positive and
    // hands its sign change to the successor); the integer work is off the FMA chain
    double p_prev = 1.0, p_cur = ta[0] - mu;
#if defined(__CUDA_ARCH__)
    unsigned sgn = static_cast<unsigned>(__double2hiint(p_cur)) >> 31, last = sgn;
    int cnt = static_cast<int>(sgn);
#else
    unsigned last = (p_cur < 0.0) ? 1u : 0u;
    int cnt = static_cast<int>(last);
#endif
#pragma unroll
    for (int k = 1; k < 9; ++k) {
        const double p_next = fma(ta[k] - mu, p_cur, -tb2[k - 1] * p_prev);
        p_prev = p_cur;
        p_cur = p_next;
#if defined(__CUDA_ARCH__)
        const unsigned sg = static_cast<unsigned>(__double2hiint(p_cur)) >> 31;
#else
        const unsigned sg = (p_cur < 0.0) ? 1u : 0u;
#endif
        cnt += static_cast<int>(sg ^ last);
        last = sg;
    }
    return cnt;
}

// Shift of lane `lane` in Sturm probe `sub` (pure function of the bracket: any lane can recompute any other lane's
// shift, so a probe exchanges nothing but its ballot).  sub 0: geometric ladder over [1e-13 tr, hi] (b.geo_step =
// log2(hi / 1e-13 tr) / (nlanes - 1), set by tri9_probe_begin); later probes: the interior points of
// [max(lo_heur, 0), hi].  No fp64 division or transcendental on this path (~200 cycles each, three shifts per probe).
FEPE_HD double tri9_probe_shift(const Eig9Bracket& b, int lane, int nlanes, int sub) {
    if (sub == 0)
        return (1e-13 * b.tr) * static_cast<double>(exp2f(static_cast<float>(lane) * b.geo_step));
    const double lo = (b.lo_heur > 0.0) ? b.lo_heur : 0.0;
    return fma(b.hi - lo, static_cast<double>(static_cast<float>(lane + 1) * (1.0f / static_cast<float>(nlanes + 1))), lo);
}
FEPE_HD void tri9_probe_begin(Eig9Bracket& b, int nlanes) {
    const double a = 1e-13 * b.tr;
    const double top = (b.hi > 2.0 * a) ? b.hi : 2.0 * a;
    b.geo_step = log2f(static_cast<float>(top * fast_rcp(a))) * (1.0f / static_cast<float>(nlanes - 1));
}

// Fold a probe's outcome into the bracket: lanes below `first_fail` counted no eigenvalue under their shift.
// Only the heuristic end moves -- b.lo stays the last shift whose FACTORISATION had inertia 0.
FEPE_HD void tri9_probe_update(Eig9Bracket& b, int first_fail, int nlanes, int sub) {
    const double new_lo = (first_fail > 0) ? tri9_probe_shift(b, first_fail - 1, nlanes, sub) : b.lo_heur;
    const double new_hi = (first_fail < nlanes) ? tri9_probe_shift(b, first_fail, nlanes, sub) : b.hi;
    b.lo_heur = new_lo;
    b.hi = new_hi;
}
// After the last probe: leave a rounding margin under the bracket and switch eig9_lane_shift to its linear ladder.
FEPE_HD void tri9_probe_finish(Eig9Bracket& b) {
    if (b.lo_heur > 0.0) b.lo_heur = b.lo_heur - 4e-16 * b.tr;
    if (b.round == 0) b.round = 1;
}

// One lane's work for a round on T: LDL^T of T - mu I (inertia count), `nsolve` inverse-iteration solves from x.
// Same outputs as eig9_lane_round.  The pivots come from the minors' recurrence (d_k = p_k / p_{k-1}, nine independent
// reciprocals) instead of the pivot recurrence (nine dependent ones); a lane whose shift is not below lambda_min may
// produce inf / NaN iterates -- its nneg is non-zero and the drivers never read the rest.
FEPE_HD void tri9_lane_round(const double (&ta)[9], const double (&tb)[8], double mu, double tiny, int nsolve,
                             double (&x)[9], int& nneg, double& rho, double& r, double& contraction) {
    (void)tiny;
    double p[10];
    p[0] = 1.0;
    p[1] = ta[0] - mu;
#pragma unroll
    for (int k = 1; k < 9; ++k) p[k + 1] = fma(ta[k] - mu, p[k], -(tb[k - 1] * tb[k - 1]) * p[k - 1]);
    nneg = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) nneg += (p[k + 1] * p[k] > 0.0) ? 0 : 1;      // zero or NaN counts as a failure
    double rd[9], l[8];
#pragma unroll
    for (int k = 0; k < 9; ++k) rd[k] = p[k] * fast_rcp(p[k + 1]);
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = tb[k] * rd[k];
    rho = 0.0;
    r = 0.0;
    contraction = 1.0;
    for (int rep = 0; rep < nsolve; ++rep) {
        double y[9];
        y[0] = x[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) y[k] = fma(-l[k - 1], y[k - 1], x[k]);
#pragma unroll
        for (int k = 0; k < 9; ++k) y[k] *= rd[k];
#pragma unroll
        for (int k = 7; k >= 0; --k) y[k] = fma(-l[k], y[k + 1], y[k]);
        const double inv = fast_rsqrt(dot9(y, y));
        const double c = dot9(y, x) * inv;              // x'.x with x' = y / |y|
        double e[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { y[i] *= inv; e[i] = fma(-c, y[i], x[i]); x[i] = y[i]; }
        rho = mu + c * inv;
        const double r_new = fast_sqrt(dot9(e, e)) * inv;
        if (rep > 0) contraction = r_new * fast_rcp(r + 1e-300);
        r = r_new;
    }
}

// Canonical sign: the entry of largest magnitude is made positive (LAPACK's sign is arbitrary).
FEPE_HD void canonical_sign9(const double (&x)[9], double (&f)[9]) {
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
}

FEPE_HD void eig9_start_vector(double (&x)[9]) {
    const double x0[9] = {0.3713906763541037, -0.2228344058124622, 0.4456688116249244, 0.1485562705416415,
                          -0.5199469468957452, 0.2971125410832830, -0.0742781352708207, 0.3342516087186933,
                          0.3565350492999395};
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = x0[i];
}

// Serial driver on the tridiagonal form (one lane per pair: fepe_solve_kernel).  One Householder reduction, then
// lambda_min is bracketed by BISECTION on Sturm counts -- a count is nine dependent FMAs, and every lane of a warp runs
// the same fixed number of them, so 32 different pairs do not diverge -- and a factorisation just under the bracket
// makes inverse iteration converge in its first round (two solves); the loop below only continues for tiny eigen-gaps.
template <class G36>
FEPE_HD int eig9_smallest_tri(const G36& g36, double (&f)[9], double& lambda) {
    double ta[9], tb[8], hv[28], htau[7];
    tridiag9(g36, ta, tb, hv, htau);
    const double tr_g = tri9_normalise(ta, tb);
    if (!(tr_g > 0.0)) {                  // empty / all-zero-weight / non-finite input
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
        lambda = 0.0;
        return 0;
    }
    // from here on T has unit trace
    double tb2[8], dmin = 1e300;
#pragma unroll
    for (int i = 0; i < 8; ++i) tb2[i] = tb[i] * tb[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) dmin = ta[i] < dmin ? ta[i] : dmin;
    double mu = -1e-14;                   // G + 1e-14 tr I is numerically positive definite
    {
        double lo = 1e-13, hi = (dmin > 2e-13) ? dmin : 2e-13;     // e_i^T T e_i >= lambda_min
        const bool below = tri9_sturm_count(ta, tb2, lo) > 0;      // lambda_min < 1e-13 tr: keep the safe shift
        // 6 geometric steps (13 decades -> a factor < 2), then 18 arithmetic ones (-> ~4e-6 relative)
#pragma unroll 1
        for (int step = 0; step < 24; ++step) {
            const double mid = (step < 6) ? fast_sqrt(lo * hi) : 0.5 * (lo + hi);
            if (tri9_sturm_count(ta, tb2, mid) > 0) hi = mid; else lo = mid;
        }
        if (!below) mu = lo - 4e-16;
    }
    double x[9];
    eig9_start_vector(x);
    double lo = -1e-14;
    double rho = 0.0;
    double r_prev = -1.0;
    int it = 0;
    for (; it < 24; ++it) {
        double xl[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) xl[i] = x[i];
        int nneg;
        double rho_l, r, c_l;
        tri9_lane_round(ta, tb, mu, 1e-18, (it == 0) ? 2 : 1, xl, nneg, rho_l, r, c_l);
        if (nneg > 0) {         // overshot lambda_min: go back half way to the last safe shift
            mu = 0.5 * (mu + lo);
            continue;
        }
        lo = mu;
        rho = rho_l;
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = xl[i];
        if (r <= 1e-17) { ++it; break; }
        // eigenVECTOR error ~ r / (lambda_8 - lambda_9); the gap follows from the contraction q between two solves:
        // gap = (rho - mu)(1/q - 1).  The first round measures q itself (two solves at one shift).
        const double q = (it == 0) ? c_l : ((r_prev > 0.0) ? r * fast_rcp(r_prev) : 2.0);
        if (it == 0 || r_prev >= 0.0) {
            if (q < 1.0 && r * q <= 1e-8 * (rho - mu) * (1.0 - q)) { ++it; break; }
            if (q >= 0.5 && ((it == 0) ? r : r_prev) <= 1e-9) { ++it; break; }   // stagnated at the rounding floor
        }
        r_prev = r;
        const double cand = rho - r * 1.0000001 - 4e-16;
        if (cand > mu) mu = cand;
    }
    tridiag9_back(hv, htau, x);
    canonical_sign9(x, f);
    lambda = rho * tr_g;
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

// One refinement step of the smallest eigenvector against an INEXACT Gram matrix: g36 is the Gram as
// accumulated (e.g. in fp32), (f0, lambda0) its smallest eigenpair, g = X^T (X f0) formed from the constraint rows with
// exact products and fp64 sums.  f1 = normalise(f0 - (G - lambda0)^+ (g - (f0.g) f0)), sign convention of
// canonical_sign9.  X^T scales the rounding of the row products by sigma_8 only, so f1 has the accuracy class of an SVD
// of X while G only needs to be good enough for the correction to contract (its relative error against the eigen-gap).
// Not yet called by a kernel; tested on the host (tests/test_math_host.py).
FEPE_HD void eig9_refine_step(const double* __restrict__ g36, const double (&f0)[9], double lambda0,
                              const double (&g)[9], double (&f1)[9]) {
    double rho = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) rho += f0[i] * g[i];
    double res[9], z[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) res[i] = g[i] - rho * f0[i];
    eig9_pinv_apply(g36, f0, lambda0, res, z);
    double x[9], n2 = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) { x[i] = f0[i] - z[i]; n2 += x[i] * x[i]; }
    const double inv = 1.0 / sqrt(n2);
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] *= inv;
    canonical_sign9(x, f1);
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

// Smallest right singular vector of a 3x3 matrix: eigenvector of M = A^T A for its smallest
// eigenvalue.  lambda_3 by Newton on the characteristic polynomial starting at 0 (monotone from
// the left of the smallest root, quadratic when lambda_3 << lambda_2 as for any F worth fitting),
// the vector as the largest column of adj(M - lambda_3 I).  ~1/10 of the latency of a Jacobi SVD.
FEPE_HD void smallest_right_sv3(const double (&A)[9], double (&v)[3], double& sigma3) {
    const double m00 = A[0] * A[0] + A[3] * A[3] + A[6] * A[6];
    const double m01 = A[0] * A[1] + A[3] * A[4] + A[6] * A[7];
    const double m02 = A[0] * A[2] + A[3] * A[5] + A[6] * A[8];
    const double m11 = A[1] * A[1] + A[4] * A[4] + A[7] * A[7];
    const double m12 = A[1] * A[2] + A[4] * A[5] + A[7] * A[8];
    const double m22 = A[2] * A[2] + A[5] * A[5] + A[8] * A[8];
    const double c2 = m00 + m11 + m22;
    const double c1 = (m00 * m11 - m01 * m01) + (m00 * m22 - m02 * m02) + (m11 * m22 - m12 * m12);
    const double det = m00 * (m11 * m22 - m12 * m12) - m01 * (m01 * m22 - m12 * m02) + m02 * (m01 * m12 - m11 * m02);
    const double c0 = det > 0.0 ? det : 0.0;
    double lam = 0.0;
    if (c1 > 0.0) {
        for (int it = 0; it < 16; ++it) {
            const double q = ((lam - c2) * lam + c1) * lam - c0;
            const double dq = (3.0 * lam - 2.0 * c2) * lam + c1;
            if (!(dq > 0.0)) break;
            const double step = -q * fast_rcp(dq);
            if (!(step > 1e-17 * c2)) break;
            lam += step;
        }
    }
    const double a00 = m00 - lam, a11 = m11 - lam, a22 = m22 - lam;
    // rows r0=(a00,m01,m02) r1=(m01,a11,m12) r2=(m02,m12,a22); null vector = cross product of two rows
    const double x0 = m01 * m12 - m02 * a11, y0 = m02 * m01 - a00 * m12, z0 = a00 * a11 - m01 * m01;   // r0 x r1
    const double x1 = m01 * a22 - m02 * m12, y1 = m02 * m02 - a00 * a22, z1 = a00 * m12 - m01 * m02;   // r0 x r2
    const double x2 = a11 * a22 - m12 * m12, y2 = m12 * m02 - m01 * a22, z2 = m01 * m12 - a11 * m02;   // r1 x r2
    const double n0 = x0 * x0 + y0 * y0 + z0 * z0, n1 = x1 * x1 + y1 * y1 + z1 * z1, n2 = x2 * x2 + y2 * y2 + z2 * z2;
    double vx = x0, vy = y0, vz = z0, nn = n0;
    if (n1 > nn) { vx = x1; vy = y1; vz = z1; nn = n1; }
    if (n2 > nn) { vx = x2; vy = y2; vz = z2; nn = n2; }
    if (!(nn > 0.0)) { vx = 0.0; vy = 0.0; vz = 1.0; nn = 1.0; }
    const double inv = fast_rsqrt(nn);
    v[0] = vx * inv; v[1] = vy * inv; v[2] = vz * inv;
    const double w0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    const double w1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    const double w2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    sigma3 = fast_sqrt(w0 * w0 + w1 * w1 + w2 * w2);
}

// Rank-2 projection of F0 = reshape(f): F0 - sigma3 u3 v3^T  (DeepFNet.py:236-237, S*[1,1,0]).
// Since sigma3 u3 = F0 v3 this is F0 (I - v3 v3^T); only v3 is needed.
FEPE_HD void rank2_project(const double (&F0)[9], double (&F2)[9], double (&v3)[3], double& sigma3) {
    smallest_right_sv3(F0, v3, sigma3);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double w = F0[3 * r] * v3[0] + F0[3 * r + 1] * v3[1] + F0[3 * r + 2] * v3[2];
        F2[3 * r] = F0[3 * r] - w * v3[0];
        F2[3 * r + 1] = F0[3 * r + 1] - w * v3[1];
        F2[3 * r + 2] = F0[3 * r + 2] - w * v3[2];
    }
}

}  // namespace fepe

namespace fepe {

// Direct 3x3 SVD built from two smallest-singular-vector solves and ONE plane rotation:
//   v3 / u3 = smallest right / left singular vectors (characteristic-polynomial Newton + adjugate),
//   the remaining 2-D problem A [b1 b2] (b1,b2 an orthonormal basis of v3's complement) is
//   diagonalised exactly by a single Hestenes rotation.  Same output contract as svd3 (sorted S,
//   U S V^T = A) at a fraction of the dependent-latency chain of the Jacobi sweeps; the pose head's
//   inputs are essential matrices (sigma_3 ~ 0) but the routine is exact for any 3x3.
FEPE_HD void svd3_direct(const double (&A)[9], double (&U)[9], double (&S)[3], double (&V)[9]) {
    double v3[3], s3;
    smallest_right_sv3(A, v3, s3);
    // orthonormal complement of v3: cross with the axis least aligned with it
    const double ax = fabs(v3[0]), ay = fabs(v3[1]), az = fabs(v3[2]);
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    if (ax <= ay && ax <= az) e0 = 1.0; else if (ay <= az) e1 = 1.0; else e2 = 1.0;
    double b1[3] = {v3[1] * e2 - v3[2] * e1, v3[2] * e0 - v3[0] * e2, v3[0] * e1 - v3[1] * e0};
    const double n1 = fast_rsqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    b1[0] *= n1; b1[1] *= n1; b1[2] *= n1;
    double b2[3] = {v3[1] * b1[2] - v3[2] * b1[1], v3[2] * b1[0] - v3[0] * b1[2], v3[0] * b1[1] - v3[1] * b1[0]};
    double w1[3], w2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        w1[r] = A[3 * r] * b1[0] + A[3 * r + 1] * b1[1] + A[3 * r + 2] * b1[2];
        w2[r] = A[3 * r] * b2[0] + A[3 * r + 1] * b2[1] + A[3 * r + 2] * b2[2];
    }
    const double alpha = w1[0] * w1[0] + w1[1] * w1[1] + w1[2] * w1[2];
    const double beta = w2[0] * w2[0] + w2[1] * w2[1] + w2[2] * w2[2];
    const double gamma = w1[0] * w2[0] + w1[1] * w2[1] + w1[2] * w2[2];
    double c = 1.0, s = 0.0;
    if (gamma * gamma > 1e-32 * alpha * beta) {
        const double zeta = (beta - alpha) * fast_rcp(2.0 * gamma);
        const double t = ((zeta >= 0.0) ? 1.0 : -1.0) * fast_rcp(fabs(zeta) + fast_sqrt(1.0 + zeta * zeta));
        c = fast_rsqrt(1.0 + t * t);
        s = c * t;
    }
    double p1[3], p2[3], q1[3], q2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        p1[r] = c * w1[r] - s * w2[r]; p2[r] = s * w1[r] + c * w2[r];
        q1[r] = c * b1[r] - s * b2[r]; q2[r] = s * b1[r] + c * b2[r];
    }
    double na = fast_sqrt(p1[0] * p1[0] + p1[1] * p1[1] + p1[2] * p1[2]);
    double nb = fast_sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
    if (na < nb) {
#pragma unroll
        for (int r = 0; r < 3; ++r) { double tmp = p1[r]; p1[r] = p2[r]; p2[r] = tmp; tmp = q1[r]; q1[r] = q2[r]; q2[r] = tmp; }
        const double tmp = na; na = nb; nb = tmp;
    }
    const double ia = fast_rcp(na + 1e-300), ib = fast_rcp(nb + 1e-300);
    double u1[3], u2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) { u1[r] = p1[r] * ia; u2[r] = p2[r] * ib; }
    // re-orthogonalise u2 against u1 (matters only when sigma_2 is itself tiny) and complete both bases
    {
        const double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
        u2[0] -= d * u1[0]; u2[1] -= d * u1[1]; u2[2] -= d * u1[2];
        const double n = fast_rsqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2] + 1e-300);
        u2[0] *= n; u2[1] *= n; u2[2] *= n;
    }
    double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    // sign of v3 so that u3^T A v3 = +sigma3 >= 0
    const double t0 = A[0] * v3[0] + A[1] * v3[1] + A[2] * v3[2];
    const double t1 = A[3] * v3[0] + A[4] * v3[1] + A[5] * v3[2];
    const double t2 = A[6] * v3[0] + A[7] * v3[1] + A[8] * v3[2];
    const double sg = (u3[0] * t0 + u3[1] * t1 + u3[2] * t2 < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        U[3 * r] = u1[r]; U[3 * r + 1] = u2[r]; U[3 * r + 2] = u3[r];
        V[3 * r] = q1[r]; V[3 * r + 1] = q2[r]; V[3 * r + 2] = sg * v3[r];
    }
    S[0] = na; S[1] = nb; S[2] = s3;
}

// ---------------------------------------------------------------------------------------------
// Pose head pieces (deepFEPE/dsac_tools/utils_F.py:478-498 _get_M2s,
// deepFEPE/dsac_tools/utils_geo.py:58-86 _R_to_q).
// ---------------------------------------------------------------------------------------------

// E = U S V^T -> R1 = U W V^T, R2 = U W^T V^T, t = u3.  U and V are completed to proper rotations
// (u3 = u1 x u2, v3 = v1 x v2), which makes det(U W V^T) = +1 without the reference's "W = -W if
// det < 0" test and yields the same SET {R1,R2}, {t,-t} whatever signs LAPACK would have picked.
FEPE_HD void essential_decompose(const double (&E)[9], double (&R1)[9], double (&R2)[9], double (&t)[3],
                                 double (&U)[9], double (&S)[3], double (&V)[9]) {
    svd3_direct(E, U, S, V);
    U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    V[2] = V[3] * V[7] - V[6] * V[4]; V[5] = V[6] * V[1] - V[0] * V[7]; V[8] = V[0] * V[4] - V[3] * V[1];
    // U W = [u2, -u1, u3],  U W^T = [-u2, u1, u3]   (columns)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double a = U[3 * i + 1] * V[3 * j] - U[3 * i] * V[3 * j + 1];
            const double b = U[3 * i + 2] * V[3 * j + 2];
            R1[3 * i + j] = a + b;
            R2[3 * i + j] = b - a;
        }
    }
    const double n = fast_rsqrt(U[2] * U[2] + U[5] * U[5] + U[8] * U[8] + 1e-300);
    t[0] = U[2] * n; t[1] = U[5] * n; t[2] = U[8] * n;
}

// Trace-method quaternion (w,x,y,z) with w >= 0; same four branches as the reference (m = R^T).
FEPE_HD void rot_to_quat(const double (&R)[9], double (&q)[4]) {
    // m[i][j] = R[j][i]
    const double m00 = R[0], m11 = R[4], m22 = R[8];
    const double m01 = R[3], m10 = R[1], m02 = R[6], m20 = R[2], m12 = R[7], m21 = R[5];
    double tr;
    if (m22 < 0.0) {
        if (m00 > m11) { tr = 1.0 + m00 - m11 - m22; q[0] = m12 - m21; q[1] = tr; q[2] = m01 + m10; q[3] = m20 + m02; }
        else           { tr = 1.0 - m00 + m11 - m22; q[0] = m20 - m02; q[1] = m01 + m10; q[2] = tr; q[3] = m12 + m21; }
    } else {
        if (m00 < -m11) { tr = 1.0 - m00 - m11 + m22; q[0] = m01 - m10; q[1] = m20 + m02; q[2] = m12 + m21; q[3] = tr; }
        else            { tr = 1.0 + m00 + m11 + m22; q[0] = tr; q[1] = m12 - m21; q[2] = m20 - m02; q[3] = m01 - m10; }
    }
    const double s = 0.5 * fast_rsqrt(tr);     // tr >= 1 on every branch
    const double sg = (q[0] < 0.0) ? -s : s;
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] *= sg;
}

// y = (M - mu I)^+ c on the complement of the unit eigenvector v (M symmetric 3x3 given by its six
// entries, mu its eigenvalue for v): solve (M - mu I + tau v v^T) y' = c - v (v.c) by the adjugate.
FEPE_HD void sym3_pinv_apply(double m00, double m01, double m02, double m11, double m12, double m22, double mu,
                             const double (&v)[3], const double (&c)[3], double (&y)[3]) {
    const double tau = (m00 + m11 + m22) * (1.0 / 3.0) + 1e-300;
    const double a00 = m00 - mu + tau * v[0] * v[0], a01 = m01 + tau * v[0] * v[1], a02 = m02 + tau * v[0] * v[2];
    const double a11 = m11 - mu + tau * v[1] * v[1], a12 = m12 + tau * v[1] * v[2], a22 = m22 - mu + tau * v[2] * v[2];
    const double vc = v[0] * c[0] + v[1] * c[1] + v[2] * c[2];
    const double r0 = c[0] - v[0] * vc, r1 = c[1] - v[1] * vc, r2 = c[2] - v[2] * vc;
    const double k00 = a11 * a22 - a12 * a12, k01 = a02 * a12 - a01 * a22, k02 = a01 * a12 - a02 * a11;
    const double k11 = a00 * a22 - a02 * a02, k12 = a01 * a02 - a00 * a12, k22 = a00 * a11 - a01 * a01;
    const double det = a00 * k00 + a01 * k01 + a02 * k02;
    const double id = 1.0 / ((fabs(det) > 1e-300) ? det : 1e-300);
    const double y0 = (k00 * r0 + k01 * r1 + k02 * r2) * id;
    const double y1 = (k01 * r0 + k11 * r1 + k12 * r2) * id;
    const double y2 = (k02 * r0 + k12 * r1 + k22 * r2) * id;
    const double vy = v[0] * y0 + v[1] * y1 + v[2] * y2;
    y[0] = y0 - v[0] * vy; y[1] = y1 - v[1] * vy; y[2] = y2 - v[2] * vy;
}

// Adjoint of the rank-2 projection F2 = F0 (I - v v^T), v = v3(F0):  given Abar = dL/dF2 returns dL/dF0.
//   dF2 = dF0 P - F0 (dv v^T + v dv^T),  dv = -(M - mu I)^+ (dF0^T F0 + F0^T dF0) v,  M = F0^T F0
//   => F0bar = Abar P + (F0 v) y^T + (F0 y) v^T   with  y = (M - mu I)^+ (F0^T Abar v + Abar^T F0 v).
FEPE_HD void rank2_project_adjoint(const double (&F0)[9], const double (&v)[3], const double (&Ab)[9],
                                   double (&F0b)[9]) {
    const double m00 = F0[0] * F0[0] + F0[3] * F0[3] + F0[6] * F0[6];
    const double m01 = F0[0] * F0[1] + F0[3] * F0[4] + F0[6] * F0[7];
    const double m02 = F0[0] * F0[2] + F0[3] * F0[5] + F0[6] * F0[8];
    const double m11 = F0[1] * F0[1] + F0[4] * F0[4] + F0[7] * F0[7];
    const double m12 = F0[1] * F0[2] + F0[4] * F0[5] + F0[7] * F0[8];
    const double m22 = F0[2] * F0[2] + F0[5] * F0[5] + F0[8] * F0[8];
    double Fv[3], Av[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Fv[r] = F0[3 * r] * v[0] + F0[3 * r + 1] * v[1] + F0[3 * r + 2] * v[2];
        Av[r] = Ab[3 * r] * v[0] + Ab[3 * r + 1] * v[1] + Ab[3 * r + 2] * v[2];
    }
    const double mu = Fv[0] * Fv[0] + Fv[1] * Fv[1] + Fv[2] * Fv[2];     // sigma3^2 = v^T M v
    double c[3], y[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        c[k] = (F0[k] * Av[0] + F0[3 + k] * Av[1] + F0[6 + k] * Av[2]) +
               (Ab[k] * Fv[0] + Ab[3 + k] * Fv[1] + Ab[6 + k] * Fv[2]);
    sym3_pinv_apply(m00, m01, m02, m11, m12, m22, mu, v, c, y);
    double Fy[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) Fy[r] = F0[3 * r] * y[0] + F0[3 * r + 1] * y[1] + F0[3 * r + 2] * y[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            F0b[3 * r + k] = Ab[3 * r + k] - Av[r] * v[k] + Fv[r] * y[k] + Fy[r] * v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// Adjoint of the pose head (what autograd does through torch.svd / _get_M2s / _R_to_q / _l2_error in the
// reference, train_good_utils.py:96-188).  All 3x3 row-major, fp64.
// ---------------------------------------------------------------------------------------------

// dL/dA from dL/dU, dL/dV for A = U diag(S) V^T (square, S may carry a signed / zero last entry):
//   dA = U [ (K o (U^T Ubar - Ubar^T U)) S + S (K o (V^T Vbar - Vbar^T V)) ] V^T,  K_ij = 1/(s_j^2 - s_i^2), i != j
// (the formula PyTorch's svd_backward uses when Sbar = 0).  Nearly equal singular values make K huge -- as in
// the reference; the denominators are only protected against an exact zero.
FEPE_HD void svd3_adjoint(const double (&U)[9], const double (&S)[3], const double (&V)[9], const double (&Ub)[9],
                          const double (&Vb)[9], double (&Ab)[9]) {
    double P[9], Q[9], T[9];
    mat3_mul_tn(U, Ub, P);      // U^T Ubar
    mat3_mul_tn(V, Vb, Q);      // V^T Vbar
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double t = 0.0;
            if (i != j) {
                double den = S[j] * S[j] - S[i] * S[i];
                if (fabs(den) < 1e-300) den = (den < 0.0) ? -1e-300 : 1e-300;
                const double k = 1.0 / den;
                t = k * (P[3 * i + j] - P[3 * j + i]) * S[j] + S[i] * k * (Q[3 * i + j] - Q[3 * j + i]);
            }
            T[3 * i + j] = t;
        }
    }
    double UT[9];
    mat3_mul(U, T, UT);
    mat3_mul_nt(UT, V, Ab);
}

// Adjoint of rot_to_quat: given qbar (4) returns Rbar (9).  Same branch as the forward.
FEPE_HD void rot_to_quat_adjoint(const double (&R)[9], const double (&qb)[4], double (&Rb)[9]) {
    // m[i][j] = R[j][i];  v = branch-dependent linear forms of m, tr = the "t" entry, q = sg * 0.5 v / sqrt(tr)
    const double m00 = R[0], m11 = R[4], m22 = R[8];
    const double m01 = R[3], m10 = R[1], m02 = R[6], m20 = R[2], m12 = R[7], m21 = R[5];
    double v[4], tr;
    int branch;
    if (m22 < 0.0) {
        if (m00 > m11) { branch = 0; tr = 1.0 + m00 - m11 - m22; v[0] = m12 - m21; v[1] = tr; v[2] = m01 + m10; v[3] = m20 + m02; }
        else           { branch = 1; tr = 1.0 - m00 + m11 - m22; v[0] = m20 - m02; v[1] = m01 + m10; v[2] = tr; v[3] = m12 + m21; }
    } else {
        if (m00 < -m11) { branch = 2; tr = 1.0 - m00 - m11 + m22; v[0] = m01 - m10; v[1] = m20 + m02; v[2] = m12 + m21; v[3] = tr; }
        else            { branch = 3; tr = 1.0 + m00 + m11 + m22; v[0] = tr; v[1] = m12 - m21; v[2] = m20 - m02; v[3] = m01 - m10; }
    }
    const double isq = 1.0 / sqrt(tr);
    const double sg = (v[0] < 0.0) ? -1.0 : 1.0;          // sign(q0) = sign(v0)
    // q_i = sg * 0.5 * v_i * tr^-1/2  =>  vbar_i = sg 0.5 isq qbar_i ;  trbar = -sg 0.25 tr^-3/2 sum_i v_i qbar_i
    double vb[4];
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { vb[i] = sg * 0.5 * isq * qb[i]; dot += v[i] * qb[i]; }
    const double trb = -sg * 0.25 * isq * isq * isq * dot;
    // accumulate mbar[i][j]
    double mb[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // index 3*i+j for m_ij
    const int I00 = 0, I01 = 1, I02 = 2, I10 = 3, I11 = 4, I12 = 5, I20 = 6, I21 = 7, I22 = 8;
    double tb = trb;
    if (branch == 0) {
        tb += vb[1]; mb[I12] += vb[0]; mb[I21] -= vb[0]; mb[I01] += vb[2]; mb[I10] += vb[2]; mb[I20] += vb[3]; mb[I02] += vb[3];
        mb[I00] += tb; mb[I11] -= tb; mb[I22] -= tb;
    } else if (branch == 1) {
        tb += vb[2]; mb[I20] += vb[0]; mb[I02] -= vb[0]; mb[I01] += vb[1]; mb[I10] += vb[1]; mb[I12] += vb[3]; mb[I21] += vb[3];
        mb[I00] -= tb; mb[I11] += tb; mb[I22] -= tb;
    } else if (branch == 2) {
        tb += vb[3]; mb[I01] += vb[0]; mb[I10] -= vb[0]; mb[I20] += vb[1]; mb[I02] += vb[1]; mb[I12] += vb[2]; mb[I21] += vb[2];
        mb[I00] -= tb; mb[I11] -= tb; mb[I22] += tb;
    } else {
        tb += vb[0]; mb[I12] += vb[1]; mb[I21] -= vb[1]; mb[I20] += vb[2]; mb[I02] -= vb[2]; mb[I01] += vb[3]; mb[I10] -= vb[3];
        mb[I00] += tb; mb[I11] += tb; mb[I22] += tb;
    }
    // R = m^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Rb[3 * i + j] = mb[3 * j + i];
}

// Adjoint of the whole head for one (layer, pair).  Inputs: E_cam (= E^T, what was decomposed), the GT pose, which
// candidates won (q_first, t_first) and the upstream gradients of the two L2 errors.  Output: dL/dE_cam.
FEPE_HD void pose_head_adjoint(const double (&Ec)[9], const double (&qg)[4], const double (&tg_unit)[3], bool q_first,
                               bool t_first, double gq, double gt, double (&Ecb)[9]) {
    double R1[9], R2[9], t[3], U[9], S[3], V[9];
    essential_decompose(Ec, R1, R2, t, U, S, V);
    // essential_decompose overwrote the third columns with u1 x u2 / v1 x v2: (U, S', V) is still a factorisation of
    // Ec with a SIGNED third value S'_3 = u3^T Ec v3 = +-S_3, which is what the adjoint formula needs.
    {
        const double w0 = Ec[0] * V[2] + Ec[1] * V[5] + Ec[2] * V[8];
        const double w1 = Ec[3] * V[2] + Ec[4] * V[5] + Ec[5] * V[8];
        const double w2 = Ec[6] * V[2] + Ec[7] * V[5] + Ec[8] * V[8];
        S[2] = U[2] * w0 + U[5] * w1 + U[8] * w2;
    }
    double Rb[9], ub3[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rb[i] = 0.0;
    {
        double q[4], Rsel[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rsel[i] = q_first ? R1[i] : R2[i];
        rot_to_quat(Rsel, q);
        double d[4], n2 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { d[i] = q[i] - qg[i]; n2 += d[i] * d[i]; }
        const double n = sqrt(n2);
        if (n > 0.0 && gq != 0.0) {
            double qb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qb[i] = gq * d[i] / n;
            rot_to_quat_adjoint(Rsel, qb, Rb);
        }
    }
    {
        const double sgn = t_first ? 1.0 : -1.0;
        double d[3], n2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { d[i] = sgn * t[i] - tg_unit[i]; n2 += d[i] * d[i]; }
        const double n = sqrt(n2);
        // t = u3 / |u3| with |u3| = 1:  u3bar = (I - t t^T) tbar
        double tb[3] = {0.0, 0.0, 0.0};
        if (n > 0.0 && gt != 0.0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) tb[i] = sgn * gt * d[i] / n;
        }
        const double tt = t[0] * tb[0] + t[1] * tb[1] + t[2] * tb[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) ub3[i] = tb[i] - t[i] * tt;
    }
    // R1 = U W V^T, R2 = U W^T V^T with U W = [u2, -u1, u3], U W^T = [-u2, u1, u3]:
    //   R = sum_k c_k(U) v_k^T  =>  Ubar, Vbar below (s = +1 for R1, -1 for R2 on the first two columns)
    const double s = q_first ? 1.0 : -1.0;
    double Ub[9], Vb[9];
    // G = Rbar (3x3).  R = s (u2 v1^T - u1 v2^T) + u3 v3^T
    //   u1bar = -s G v2, u2bar = s G v1, u3bar = G v3 ;  v1bar = s G^T u2, v2bar = -s G^T u1, v3bar = G^T u3
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double Gv1 = Rb[3 * r] * V[0] + Rb[3 * r + 1] * V[3] + Rb[3 * r + 2] * V[6];
        const double Gv2 = Rb[3 * r] * V[1] + Rb[3 * r + 1] * V[4] + Rb[3 * r + 2] * V[7];
        const double Gv3 = Rb[3 * r] * V[2] + Rb[3 * r + 1] * V[5] + Rb[3 * r + 2] * V[8];
        Ub[3 * r] = -s * Gv2; Ub[3 * r + 1] = s * Gv1; Ub[3 * r + 2] = Gv3 + ub3[r];
        const double Gtu1 = Rb[r] * U[0] + Rb[3 + r] * U[3] + Rb[6 + r] * U[6];
        const double Gtu2 = Rb[r] * U[1] + Rb[3 + r] * U[4] + Rb[6 + r] * U[7];
        const double Gtu3 = Rb[r] * U[2] + Rb[3 + r] * U[5] + Rb[6 + r] * U[8];
        Vb[3 * r] = s * Gtu2; Vb[3 * r + 1] = -s * Gtu1; Vb[3 * r + 2] = Gtu3;
    }
    svd3_adjoint(U, S, V, Ub, Vb, Ecb);
}

}  // namespace fepe
