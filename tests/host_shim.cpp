// Host build of the product's math header (pytorch-deepfepe_b200/csrc/fepe_math.cuh) so that the
// eigen / SVD routines can be checked against LAPACK on a machine without a GPU.  Test
// infrastructure: built on the fly by tests/test_math_host.py with g++.
#include "fepe_math.cuh"

extern "C" {

int shim_eig9(const double* g36, double* f, double* lambda) {
    double ff[9];
    double lam;
    int it = fepe::eig9_smallest(g36, ff, lam);
    for (int i = 0; i < 9; ++i) f[i] = ff[i];
    *lambda = lam;
    return it;
}

void shim_pinv(const double* g36, const double* f, double lambda, const double* rhs, double* z) {
    double ff[9], rr[9], zz[9];
    for (int i = 0; i < 9; ++i) { ff[i] = f[i]; rr[i] = rhs[i]; }
    fepe::eig9_pinv_apply(g36, ff, lambda, rr, zz);
    for (int i = 0; i < 9; ++i) z[i] = zz[i];
}

void shim_svd3(const double* A, double* U, double* S, double* V) {
    double a[9], u[9], s[3], v[9];
    for (int i = 0; i < 9; ++i) a[i] = A[i];
    fepe::svd3(a, u, s, v);
    for (int i = 0; i < 9; ++i) { U[i] = u[i]; V[i] = v[i]; }
    for (int i = 0; i < 3; ++i) S[i] = s[i];
}

void shim_svd3_direct(const double* A, double* U, double* S, double* V) {
    double a[9], u[9], s[3], v[9];
    for (int i = 0; i < 9; ++i) a[i] = A[i];
    fepe::svd3_direct(a, u, s, v);
    for (int i = 0; i < 9; ++i) { U[i] = u[i]; V[i] = v[i]; }
    for (int i = 0; i < 3; ++i) S[i] = s[i];
}

void shim_rank2(const double* F0, double* F2) {
    double a[9], f2[9], v[3], s3;
    for (int i = 0; i < 9; ++i) a[i] = F0[i];
    fepe::rank2_project(a, f2, v, s3);
    for (int i = 0; i < 9; ++i) F2[i] = f2[i];
}

// Host emulation of the multi-shift drivers eig9_smallest_warp (fepe_fit.cuh, 32 lanes) and eig9_smallest_cta
// (fepe_fit.cu, 128 lanes): same scalar pieces, the ballot / shuffle / shared-memory exchange replaced by
// loops over `nlanes` virtual lanes.
int shim_eig9_multishift_n(const double* g36, double* f, double* lambda, int nlanes) {
    fepe::Eig9Bracket b;
    if (nlanes < 3 || nlanes > 128 || !fepe::eig9_bracket_init(g36, b)) {
        for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
        *lambda = 0.0;
        return 0;
    }
    const double tiny = 1e-18 * b.tr;
    double x[9];
    fepe::eig9_start_vector(x);
    double rho = 0.0;
    int rounds = 0;
    while (rounds < 10) {
        double mu[128], rho_l[128], r_l[128], c_l[128], xl[128][9];
        int nneg[128];
        for (int lane = 0; lane < nlanes; ++lane) {
            mu[lane] = fepe::eig9_lane_shift(b, lane, nlanes);
            double xx[9];
            for (int i = 0; i < 9; ++i) xx[i] = x[i];
            fepe::eig9_lane_round(g36, mu[lane], tiny, 2, xx, nneg[lane], rho_l[lane], r_l[lane], c_l[lane]);
            for (int i = 0; i < 9; ++i) xl[lane][i] = xx[i];
        }
        ++rounds;
        int first_fail = nlanes;
        for (int lane = nlanes - 1; lane >= 0; --lane) if (nneg[lane] != 0) first_fail = lane;
        const int best = first_fail - 1;
        if (best < 0) { b.lo = b.lo * 64.0 - 1e-13 * b.tr; b.lo_heur = b.lo; continue; }
        const double mu_fail = (first_fail < nlanes) ? mu[first_fail] : -1.0;
        rho = rho_l[best];
        for (int i = 0; i < 9; ++i) x[i] = xl[best][i];
        if (fepe::eig9_bracket_update(b, mu[best], mu_fail, rho, r_l[best], c_l[best])) break;
    }
    double ff[9];
    fepe::canonical_sign9(x, ff);
    for (int i = 0; i < 9; ++i) f[i] = ff[i];
    *lambda = rho;
    return rounds;
}
int shim_eig9_multishift(const double* g36, double* f, double* lambda) {
    return shim_eig9_multishift_n(g36, f, lambda, 32);
}
int shim_eig9_multishift128(const double* g36, double* f, double* lambda) {
    return shim_eig9_multishift_n(g36, f, lambda, 128);
}

int shim_g36_index(int r, int c) { return fepe::g36_index(r, c); }

void shim_essential(const double* E, double* R1, double* R2, double* t, double* q1, double* q2) {
    double e[9], r1[9], r2[9], tt[3], u[9], s[3], v[9], qa[4], qb[4];
    for (int i = 0; i < 9; ++i) e[i] = E[i];
    fepe::essential_decompose(e, r1, r2, tt, u, s, v);
    fepe::rot_to_quat(r1, qa);
    fepe::rot_to_quat(r2, qb);
    for (int i = 0; i < 9; ++i) { R1[i] = r1[i]; R2[i] = r2[i]; }
    for (int i = 0; i < 3; ++i) t[i] = tt[i];
    for (int i = 0; i < 4; ++i) { q1[i] = qa[i]; q2[i] = qb[i]; }
}

void shim_rank2_adjoint(const double* F0, const double* Ab, double* F0b) {
    double a[9], ab[9], f2[9], v[3], s3, out[9];
    for (int i = 0; i < 9; ++i) { a[i] = F0[i]; ab[i] = Ab[i]; }
    fepe::rank2_project(a, f2, v, s3);
    fepe::rank2_project_adjoint(a, v, ab, out);
    for (int i = 0; i < 9; ++i) F0b[i] = out[i];
}

void shim_pose_adjoint(const double* Ec, const double* qg, const double* tg, int q_first, int t_first, double gq,
                       double gt, double* Ecb) {
    double e[9], q[4], t[3], out[9];
    for (int i = 0; i < 9; ++i) e[i] = Ec[i];
    for (int i = 0; i < 4; ++i) q[i] = qg[i];
    for (int i = 0; i < 3; ++i) t[i] = tg[i];
    fepe::pose_head_adjoint(e, q, t, q_first != 0, t_first != 0, gq, gt, out);
    for (int i = 0; i < 9; ++i) Ecb[i] = out[i];
}

void shim_quat(const double* R, double* q) {
    double r[9], qq[4];
    for (int i = 0; i < 9; ++i) r[i] = R[i];
    fepe::rot_to_quat(r, qq);
    for (int i = 0; i < 4; ++i) q[i] = qq[i];
}
}
