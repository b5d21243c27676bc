// Microbenchmark of the Gram inner body (development aid): how many SM cycles does one warp-iteration of
// the 44-instruction fp64 burst cost with 16 warps per SM, with and without its fp32 front end?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../pytorch-deepfepe_b200/csrc/fepe_fit_passes.cuh"
using namespace fepe;

#define ITER 1024
// MODE 0: fp64 burst only (inputs already doubles)   1: + 5 F2F   2: + full fp32 front end from registers
// MODE 3: as the real loop, data from shared memory
template <int MODE> __global__ void __launch_bounds__(512, 1) k_body(double* out, float seed, long long* cyc) {
    __shared__ float4 sp[1024];
    __shared__ float sw[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { sp[i] = make_float4(i * 1e-3f, seed, i * 2e-3f, -seed); sw[i] = 1.f + i * 1e-4f; }
    __syncthreads();
    PairMap m{1.1f, 0.9f, 1.05f, 0.95f, -0.5f, -0.4f, -0.6f, -0.3f};
    double acc[36];
    for (int i = 0; i < 36; ++i) acc[i] = 0.0;
    GramTerm c; c.x1 = seed; c.y1 = seed * 2; c.x2 = seed * 3; c.y2 = seed * 0.5; c.s = 1.0;
    float fx1 = seed, fy1 = seed * 2, fx2 = seed * 3, fy2 = seed * 0.5f, fs = 1.f;
    long long t0 = clock64();
    if (MODE == 3) {
        for (int rep = 0; rep < ITER / 16; ++rep) pass_gram(sp, sw, 1024, threadIdx.x & 63, 64, m, acc);   // re-zeroes acc: fine
    } else {
#pragma unroll 2
        for (int it = 0; it < ITER; ++it) {
            if (MODE == 1) { c.x1 = fx1; c.y1 = fy1; c.x2 = fx2; c.y2 = fy2; c.s = fs; fx1 += 1e-3f; fy1 += 1e-3f; fx2 -= 1e-3f; fy2 += 2e-3f; fs += 1e-4f; }
            if (MODE == 2) { c = gram_prepare(make_float4(fx1, fy1, fx2, fy2), fs, m); fx1 += 1e-3f; fy1 += 1e-3f; fx2 -= 1e-3f; fy2 += 2e-3f; fs += 1e-4f; }
            if (MODE == 0) { c.x1 += 1e-3; }
            gram_accumulate(c, acc);
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 36; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMallocManaged(&cyc, 8);
    const char* names[4] = {"fp64 burst only", "+5 F2F", "+fp32 front end (regs)", "real pass_gram from smem"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k_body<0><<<148, 512>>>(out, 0.37f, cyc);
            if (mode == 1) k_body<1><<<148, 512>>>(out, 0.37f, cyc);
            if (mode == 2) k_body<2><<<148, 512>>>(out, 0.37f, cyc);
            if (mode == 3) k_body<3><<<148, 512>>>(out, 0.37f, cyc);
            cudaDeviceSynchronize();
        }
        // 16 warps per SM, ITER iterations each: cycles per warp-iteration per SM = cyc / (16 * ITER); x4 = per SMSP
        printf("%-28s %8lld cycles: %.1f cycles per warp-iteration per SMSP (fp64 floor 88 at 2 cycles/instr), %.0f SM-cycles per 1000 correspondences\n",
               names[mode], *cyc, (double)*cyc / (4.0 * ITER), (double)*cyc / (16.0 * ITER) * 31.25);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
