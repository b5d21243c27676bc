// Microbenchmark (development aid): cycles one warp spends in each piece of the tridiagonal eigen-solver, alone on its
// SM -- the latency chain of the latency kernel's solve phase.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../pytorch-deepfepe_b200/csrc/fepe_math.cuh"
using namespace fepe;

__global__ void __launch_bounds__(64, 6) k_phases(const double* g36_in, double* out, long long* cyc) {
    __shared__ double g36[36];
    __shared__ double hv_s[28];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 36) g36[threadIdx.x] = g36_in[threadIdx.x];
    __syncthreads();
    long long t[10];
    t[0] = clock64();
    double ta[9], tb[8], htau[7];
    {
        double hv[28];
        tridiag9(g36, ta, tb, hv, htau);
        if (threadIdx.x == 0) for (int i = 0; i < 28; ++i) hv_s[i] = hv[i];
    }
    __syncthreads();
    t[1] = clock64();
    Eig9Bracket b;
    const double tr_g = tri9_normalise(ta, tb);
    tri9_bracket_init(ta, b);
    double tb2[8];
    for (int i = 0; i < 8; ++i) tb2[i] = tb[i] * tb[i];
    t[2] = clock64();
    tri9_probe_begin(b, 32);
    for (int sub = 0; sub < 5; ++sub) {
        const int cnt = tri9_sturm_count(ta, tb2, tri9_probe_shift(b, lane, 32, sub));
        const unsigned bad = ~__ballot_sync(0xffffffffu, cnt == 0);
        tri9_probe_update(b, bad ? (__ffs(bad) - 1) : 32, 32, sub);
    }
    tri9_probe_finish(b);
    t[3] = clock64();
    double x[9];
    eig9_start_vector(x);
    const double mu = eig9_lane_shift(b, lane, 32);
    int nneg; double rho, r, c;
    t[4] = clock64();
    tri9_lane_round(ta, tb, mu, 1e-18, 2, x, nneg, rho, r, c);
    t[5] = clock64();
    const unsigned bad = ~__ballot_sync(0xffffffffu, nneg == 0);
    const int best = (bad ? (__ffs(bad) - 1) : 32) - 1;
    for (int i = 0; i < 9; ++i) x[i] = __shfl_sync(0xffffffffu, x[i], best < 0 ? 0 : best);
    rho = __shfl_sync(0xffffffffu, rho, best < 0 ? 0 : best);
    t[6] = clock64();
    tridiag9_back(hv_s, htau, x);
    double f[9];
    canonical_sign9(x, f);
    t[7] = clock64();
    double F2[9], v3[3], s3;
    rank2_project(f, F2, v3, s3);
    t[8] = clock64();
    double s = rho * tr_g + s3;
    for (int i = 0; i < 9; ++i) s += F2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) for (int i = 0; i < 8; ++i) cyc[i] = t[i + 1] - t[i];
}

int main() {
    double h[36];
    // a plausible Gram matrix: sum of 40 outer products (a a^T) (x) (b b^T)
    for (int i = 0; i < 36; ++i) h[i] = 0.0;
    unsigned st = 12345;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (st >> 8) * (1.0 / 16777216.0) - 0.5; };
    for (int n = 0; n < 40; ++n) {
        double a[3] = {rnd() * 2, rnd() * 2, 1.0}, bb[3] = {rnd() * 2, rnd() * 2, 1.0};
        double mA[6] = {a[0]*a[0], a[0]*a[1], a[0]*a[2], a[1]*a[1], a[1]*a[2], a[2]*a[2]};
        double mB[6] = {bb[0]*bb[0], bb[0]*bb[1], bb[0]*bb[2], bb[1]*bb[1], bb[1]*bb[2], bb[2]*bb[2]};
        for (int u = 0; u < 6; ++u) for (int v = 0; v < 6; ++v) h[u * 6 + v] += 0.025 * mA[u] * mB[v];
    }
    double *g, *out; long long* cyc;
    cudaMalloc(&g, sizeof(h)); cudaMalloc(&out, 148 * 6 * 64 * 8); cudaMalloc(&cyc, 80);
    cudaMemcpy(g, h, sizeof(h), cudaMemcpyHostToDevice);
    const char* names[8] = {"tridiag9 (+ reflectors to smem)", "normalise + bracket init", "5 Sturm probes (32 lanes)",
                            "start vector + lane shift", "lane round (LDL + 2 solves)", "best-lane broadcast",
                            "back-transform + sign", "rank-2 projection"};
    for (int grid : {1, 148 * 6}) {
        for (int rep = 0; rep < 3; ++rep) k_phases<<<grid, 64>>>(g, out, cyc);
        cudaDeviceSynchronize();
        long long hc[8];
        cudaMemcpy(hc, cyc, 64, cudaMemcpyDeviceToHost);
        printf("grid %d x 64 threads (%s):\n", grid, grid == 1 ? "one CTA alone" : "6 CTAs per SM, all SMs");
        long long tot = 0;
        for (int i = 0; i < 8; ++i) { printf("  %-34s %7lld cycles\n", names[i], hc[i]); tot += hc[i]; }
        printf("  %-34s %7lld cycles\n", "total", tot);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
