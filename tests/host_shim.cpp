// Host build of the product's math header (pytorch-deepfepe_b200/csrc/fepe_math.cuh) so that the
// eigen / SVD routines can be checked against LAPACK on a machine without a GPU.  Test
// infrastructure: built on the fly by tests/test_math_host.py with g++.
#include "fepe_math.cuh"
#include "fepe_fit_adjoint.cuh"
#include "fepe_recover.cuh"
#include "fepe_virt.cuh"
#include <vector>

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

void shim_refine_step(const double* g36, const double* f0, double lambda0, const double* g, double* f1) {
    double ff[9], gg[9], out[9];
    for (int i = 0; i < 9; ++i) { ff[i] = f0[i]; gg[i] = g[i]; }
    fepe::eig9_refine_step(g36, ff, lambda0, gg, out);
    for (int i = 0; i < 9; ++i) f1[i] = out[i];
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

// Whole weighted 8-point fit of ONE pair and its backward (weights AND coordinates) in fp64 on the host, built from the
// product's own pieces: eig9_smallest, rank2_project(+adjoint), eig9_pinv_apply (fepe_math.cuh) and row_adjoint,
// epi_adjoint, norm_adjoint (fepe_fit_adjoint.cuh).  The loops mirror the passes of fepe_fit_bwd_kernel.
// m: [N,4] (u1,v1,u2,v2) already in the frame Fit.forward receives; outputs: F [9], res [N], epi [N], gw [N], gm [N,4].
void shim_fit_pair_fwd_bwd(const double* m, const double* w, int N, double clamp_at, const double* gF,
                           const double* gres, const double* gepi, double* F_out, double* res, double* epi,
                           double* gw, double* gm) {
    using namespace fepe;
    double s[2], cx[2], cy[2];
    for (int im = 0; im < 2; ++im) {
        double sx = 0, sy = 0, sd = 0;
        for (int i = 0; i < N; ++i) { sx += m[4 * i + 2 * im]; sy += m[4 * i + 2 * im + 1]; }
        cx[im] = sx / N; cy[im] = sy / N;
        for (int i = 0; i < N; ++i) {
            const double du = m[4 * i + 2 * im] - cx[im], dv = m[4 * i + 2 * im + 1] - cy[im];
            sd += sqrt(du * du + dv * dv);
        }
        s[im] = 1.4142 / (sd / N);
    }
    double g36[36];
    for (int e = 0; e < 36; ++e) g36[e] = 0.0;
    for (int i = 0; i < N; ++i) {
        const double x1 = s[0] * (m[4 * i] - cx[0]), y1 = s[0] * (m[4 * i + 1] - cy[0]);
        const double x2 = s[1] * (m[4 * i + 2] - cx[1]), y2 = s[1] * (m[4 * i + 3] - cy[1]);
        const double sc = w[i] * w[i] / ((x1 * x1 + y1 * y1 + 1) * (x2 * x2 + y2 * y2 + 1));
        const double a[6] = {x2 * x2, x2 * y2, x2, y2 * y2, y2, 1.0};
        const double b[6] = {x1 * x1, x1 * y1, x1, y1 * y1, y1, 1.0};
        for (int u = 0; u < 6; ++u)
            for (int v = 0; v < 6; ++v) g36[u * 6 + v] += sc * a[u] * b[v];
    }
    double f[9], lambda, F2[9], v3[3], sigma3;
    eig9_smallest(g36, f, lambda);
    rank2_project(f, F2, v3, sigma3);
    // out = T2^T F2 T1
    const double T1[9] = {s[0], 0, -s[0] * cx[0], 0, s[0], -s[0] * cy[0], 0, 0, 1};
    const double T2[9] = {s[1], 0, -s[1] * cx[1], 0, s[1], -s[1] * cy[1], 0, 0, 1};
    double tmp[9], Fo[9];
    mat3_mul(F2, T1, tmp);
    mat3_mul_tn(T2, tmp, Fo);
    for (int k = 0; k < 9; ++k) F_out[k] = Fo[k];
    // pass 1 of the backward (+ forward outputs)
    double hs[9], ge[9];
    for (int k = 0; k < 9; ++k) { hs[k] = 0; ge[k] = 0; }
    std::vector<double> cbe(4 * static_cast<size_t>(N));
    for (int i = 0; i < N; ++i) {
        const double x1 = s[0] * (m[4 * i] - cx[0]), y1 = s[0] * (m[4 * i + 1] - cy[0]);
        const double x2 = s[1] * (m[4 * i + 2] - cx[1]), y2 = s[1] * (m[4 * i + 3] - cy[1]);
        const double inv = 1.0 / sqrt((x1 * x1 + y1 * y1 + 1) * (x2 * x2 + y2 * y2 + 1));
        const double a[3] = {x2, y2, 1.0}, b[3] = {x1, y1, 1.0};
        double dot = 0;
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) {
                dot += f[3 * j + k] * a[j] * b[k];
                hs[3 * j + k] += gres[i] * w[i] * inv * a[j] * b[k];
            }
        res[i] = w[i] * dot * inv;
        double cb[4];
        epi_adjoint<double>(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3], Fo, clamp_at, gepi[i], ge, cb);
        for (int k = 0; k < 4; ++k) cbe[4 * i + k] = cb[k];
        // forward value of the epipolar distance (same expressions as epi_adjoint)
        {
            const double u1 = m[4 * i], v1 = m[4 * i + 1], u2 = m[4 * i + 2], v2 = m[4 * i + 3];
            const double l10 = u2 * Fo[0] + v2 * Fo[3] + Fo[6], l11 = u2 * Fo[1] + v2 * Fo[4] + Fo[7],
                         l12 = u2 * Fo[2] + v2 * Fo[5] + Fo[8];
            const double l20 = Fo[0] * u1 + Fo[1] * v1 + Fo[2], l21 = Fo[3] * u1 + Fo[4] * v1 + Fo[5];
            const double dd = l10 * u1 + l11 * v1 + l12;
            const double d = fabs(dd) * (1.0 / (sqrt(l10 * l10 + l11 * l11) + 1e-6) + 1.0 / (sqrt(l20 * l20 + l21 * l21) + 1e-6));
            epi[i] = d < clamp_at ? d : clamp_at;
        }
    }
    double ob[9], Ab[9], X[9];
    for (int k = 0; k < 9; ++k) ob[k] = ge[k] + gF[k];
    {   // F2bar = T2 ob T1^T
        double t[9];
        mat3_mul(T2, ob, t);
        mat3_mul_nt(t, T1, Ab);
        (void)X;
    }
    double fb[9], z[9];
    rank2_project_adjoint(f, v3, Ab, fb);
    for (int k = 0; k < 9; ++k) fb[k] += hs[k];
    eig9_pinv_apply(g36, f, lambda, fb, z);
    // pass 2
    NormAdjointSums S{};
    std::vector<double> xbs(4 * static_cast<size_t>(N));
    for (int i = 0; i < N; ++i) {
        const double du1 = m[4 * i] - cx[0], dv1 = m[4 * i + 1] - cy[0], du2 = m[4 * i + 2] - cx[1], dv2 = m[4 * i + 3] - cy[1];
        double xb[4], wb;
        row_adjoint<double>(s[0] * du1, s[0] * dv1, s[1] * du2, s[1] * dv2, w[i], gres[i], f, z, wb, xb);
        gw[i] = wb;
        for (int k = 0; k < 4; ++k) xbs[4 * i + k] = xb[k];
        S.sx[0] += xb[0]; S.sy[0] += xb[1]; S.sd[0] += xb[0] * du1 + xb[1] * dv1;
        S.sx[1] += xb[2]; S.sy[1] += xb[3]; S.sd[1] += xb[2] * du2 + xb[3] * dv2;
        const double d1 = sqrt(du1 * du1 + dv1 * dv1), d2 = sqrt(du2 * du2 + dv2 * dv2);
        if (d1 > 0) { S.dx[0] += du1 / d1; S.dy[0] += dv1 / d1; }
        if (d2 > 0) { S.dx[1] += du2 / d2; S.dy[1] += dv2 / d2; }
    }
    NormAdjointCoef C;
    norm_adjoint(ob, F2, s, cx, cy, S, N, C);
    // pass 3
    for (int i = 0; i < N; ++i) {
        for (int im = 0; im < 2; ++im) {
            const double du = m[4 * i + 2 * im] - cx[im], dv = m[4 * i + 2 * im + 1] - cy[im];
            const double d = sqrt(du * du + dv * dv), id = d > 0 ? 1.0 / d : 0.0;
            gm[4 * i + 2 * im] = s[im] * xbs[4 * i + 2 * im] + C.A[im] * du * id + C.Bx[im] + cbe[4 * i + 2 * im];
            gm[4 * i + 2 * im + 1] = s[im] * xbs[4 * i + 2 * im + 1] + C.A[im] * dv * id + C.By[im] + cbe[4 * i + 2 * im + 1];
        }
    }
}

// Host run of the pieces of fepe_recover_pose_kernel in its own order (fepe_recover.cuh + essential_decompose):
// E [9], K [9], m [N,4] pixels (as floats converted to double, like the kernel), Rt_scene [16] ->
// out: R [9], t [3], counts [4], best, err_q, err_t; mask [N] of the winner.
void shim_recover_pose(const double* E, const double* K, const float* m, int N, double thresh, const double* Rt,
                       double* R_out, double* t_out, int* counts, int* best_out, double* errs, unsigned char* mask) {
    using namespace fepe;
    double e[9], R1[9], R2[9], t[3], U[9], S[3], V[9];
    for (int i = 0; i < 9; ++i) e[i] = E[i];
    essential_decompose(e, R1, R2, t, U, S, V);
    const double inv_f = 1.0 / K[0], ppx = K[2], ppy = K[5];
    std::vector<unsigned char> bits(static_cast<size_t>(N), 0);
    for (int c = 0; c < 4; ++c) {
        double P[12];
        recover_candidate(c, R1, R2, t, P);
        counts[c] = 0;
        for (int i = 0; i < N; ++i) {
            const double x1 = (static_cast<double>(m[4 * i]) - ppx) * inv_f, y1 = (static_cast<double>(m[4 * i + 1]) - ppy) * inv_f;
            const double x2 = (static_cast<double>(m[4 * i + 2]) - ppx) * inv_f, y2 = (static_cast<double>(m[4 * i + 3]) - ppy) * inv_f;
            double X[4];
            triangulate_dlt(x1, y1, x2, y2, P, X);
            if (cheirality_ok(X, P, thresh)) { bits[i] |= static_cast<unsigned char>(1u << c); counts[c] += 1; }
        }
    }
    const int best = recover_select(counts[0], counts[1], counts[2], counts[3]);
    *best_out = best;
    double Pb[12], R[9], tt[3];
    recover_candidate(best, R1, R2, t, Pb);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = Pb[4 * r + c];
        tt[r] = Pb[4 * r + 3];
    }
    for (int i = 0; i < 9; ++i) R_out[i] = R[i];
    for (int i = 0; i < 3; ++i) t_out[i] = tt[i];
    for (int i = 0; i < N; ++i) mask[i] = ((bits[i] >> best) & 1u) ? 255 : 0;
    double Rs[9], ts[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Rs[3 * r + c] = Rt[4 * r + c];
        ts[r] = Rt[4 * r + 3];
    }
    recover_errors(R, tt, Rs, ts, errs[0], errs[1]);
}
// Host run of fepe_gt_virt_kernel's per-thread work (fepe_virt.cuh), fp32 rounding of the outputs included.
int shim_correct_matches(const double* F, const float* pts1, const float* pts2, int P, float* out1, float* out2) {
    double f[9];
    for (int i = 0; i < 9; ++i) f[i] = F[i];
    int nan_count = 0;
    for (int i = 0; i < P; ++i) {
        double o[4];
        const bool ok = fepe::correct_match_pair(f, pts1[2 * i], pts1[2 * i + 1], pts2[2 * i], pts2[2 * i + 1], o);
        if (!ok) { ++nan_count; for (int k = 0; k < 4; ++k) o[k] = 0.0; }
        out1[2 * i] = static_cast<float>(o[0]); out1[2 * i + 1] = static_cast<float>(o[1]);
        out2[2 * i] = static_cast<float>(o[2]); out2[2 * i + 1] = static_cast<float>(o[3]);
    }
    return nan_count;
}

int shim_solve_poly6(const double* k, double* t_re) {
    double kk[7], tt[6];
    for (int i = 0; i < 7; ++i) kk[i] = k[i];
    const int n = fepe::solve_poly6_cv(kk, tt);
    for (int i = 0; i < 6; ++i) t_re[i] = tt[i];
    return n;
}

void shim_gt_from_motion(const float* K, const float* Rt, double* gt) {
    double k[9], rt[16], g[32], kinv[9];
    for (int i = 0; i < 9; ++i) k[i] = K[i];
    for (int i = 0; i < 16; ++i) rt[i] = Rt[i];
    fepe::gt_from_motion(k, rt, g, kinv);
    for (int i = 0; i < 32; ++i) gt[i] = g[i];
}

// Householder tridiagonalisation of the Gram matrix + the reflectors applied to the columns of the identity (Q).
void shim_tridiag9(const double* g36, double* ta, double* tb, double* Q) {
    double a[9], b[8], hv[28], ht[7];
    fepe::tridiag9(g36, a, b, hv, ht);
    for (int i = 0; i < 9; ++i) ta[i] = a[i];
    for (int i = 0; i < 8; ++i) tb[i] = b[i];
    for (int c = 0; c < 9; ++c) {
        double x[9];
        for (int i = 0; i < 9; ++i) x[i] = (i == c) ? 1.0 : 0.0;
        fepe::tridiag9_back(hv, ht, x);
        for (int i = 0; i < 9; ++i) Q[i * 9 + c] = x[i];
    }
}

// Host emulation of the multi-shift drivers on the TRIDIAGONAL form (eig9_smallest_cta / _warp with tri9_lane_round).
int shim_eig9_tri_n(const double* g36, double* f, double* lambda, int nlanes) {
    double ta[9], tb[8], hv[28], ht[7];
    fepe::tridiag9(g36, ta, tb, hv, ht);
    fepe::Eig9Bracket b;
    const double tr_g = fepe::tri9_normalise(ta, tb);
    if (nlanes < 3 || nlanes > 128 || !(tr_g > 0.0) || !fepe::tri9_bracket_init(ta, b)) {
        for (int i = 0; i < 9; ++i) f[i] = (i == 8) ? 1.0 : 0.0;
        *lambda = 0.0;
        return 0;
    }
    const double tiny = 1e-18 * b.tr;
    {
        double tb2[8];
        for (int i = 0; i < 8; ++i) tb2[i] = tb[i] * tb[i];
        const int probes = (nlanes >= 64) ? 4 : 5;
        fepe::tri9_probe_begin(b, nlanes);
        for (int sub = 0; sub < probes; ++sub) {
            int first_fail = nlanes;
            for (int lane = nlanes - 1; lane >= 0; --lane)
                if (fepe::tri9_sturm_count(ta, tb2, fepe::tri9_probe_shift(b, lane, nlanes, sub)) != 0) first_fail = lane;
            fepe::tri9_probe_update(b, first_fail, nlanes, sub);
        }
        fepe::tri9_probe_finish(b);
    }
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
            fepe::tri9_lane_round(ta, tb, mu[lane], tiny, 2, xx, nneg[lane], rho_l[lane], r_l[lane], c_l[lane]);
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
    fepe::tridiag9_back(hv, ht, x);
    double ff[9];
    fepe::canonical_sign9(x, ff);
    for (int i = 0; i < 9; ++i) f[i] = ff[i];
    *lambda = rho * tr_g;
    return rounds;
}
int shim_eig9_tri_serial(const double* g36, double* f, double* lambda) {
    double ff[9];
    double lam;
    int it = fepe::eig9_smallest_tri(g36, ff, lam);
    for (int i = 0; i < 9; ++i) f[i] = ff[i];
    *lambda = lam;
    return it;
}
int shim_eig9_tri32(const double* g36, double* f, double* lambda) { return shim_eig9_tri_n(g36, f, lambda, 32); }
int shim_eig9_tri64(const double* g36, double* f, double* lambda) { return shim_eig9_tri_n(g36, f, lambda, 64); }

}
