// Pipe throughput / latency microbenchmarks on the target GPU (development aid, not product code).
// Prints ops/clk/SM for several instruction types, measured with clock64 on a full-chip launch.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITER 2048

template <int ILP> __global__ void k_dfma(double* out, double a, double b, long long* cyc) {
    double v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> __global__ void k_ffma(float* out, float a, float b, long long* cyc) {
    float v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> __global__ void k_f2d(double* out, float a, long long* cyc) {
    float v[ILP]; double acc[ILP];
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x * 1e-3f + i; acc[i] = 0; }
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { double d = (double)v[i]; v[i] = __double_as_longlong(d) & 0xffff ? v[i] * a : v[i]; acc[i] = d; }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; ++i) s += acc[i] + v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// two-float (double-single) accumulate: s += a*b with error-free transforms, fp32 only
template <int ILP> __global__ void k_dsfma(float* out, float a, float b, long long* cyc) {
    float hi[ILP], lo[ILP];
    for (int i = 0; i < ILP; ++i) { hi[i] = threadIdx.x * 1e-3f + i; lo[i] = 0.f; }
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            float p = a * b; float pe = fmaf(a, b, -p);        // two-prod
            float s = hi[i] + p; float bb = s - hi[i];          // two-sum
            float e = (hi[i] - (s - bb)) + (p - bb);
            hi[i] = s; lo[i] += e + pe; a += 1e-7f;
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += hi[i] + lo[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dmma(double* out, long long* cyc) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_drcp(double* out, double a, long long* cyc) {
    double v[4]; for (int i = 0; i < 4; ++i) v[i] = 1.5 + threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = 1.0 / (v[i] + a);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v[0] + v[1] + v[2] + v[3];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F> void run(const char* name, F launch, double ops_per_thread_iter, int threads, int sms) {
    long long* cyc; cudaMallocManaged(&cyc, 8); *cyc = 0;
    launch(cyc); cudaDeviceSynchronize(); launch(cyc); cudaDeviceSynchronize();
    double per_sm_clk = ops_per_thread_iter * ITER * threads / (double)*cyc;
    printf("%-28s threads/SM=%4d  cycles=%8lld  -> %7.2f lane-ops/clk/SM (%.1f cyc per warp-instr per SMSP-warp)\n", name, threads, *cyc,
           per_sm_clk, (double)*cyc / (ops_per_thread_iter * ITER));
    cudaFree(cyc);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d, clock %d kHz\n", sms, clk);
    double* dout; cudaMalloc(&dout, sizeof(double) * sms * 1024);
    float* fout = (float*)dout;
    for (int threads : {32, 128, 256, 512, 1024}) {
        run("DFMA ilp1 (latency)", [&](long long* c) { k_dfma<1><<<sms, threads>>>(dout, 1.0000001, 1e-9, c); }, 1, threads, sms);
        run("DFMA ilp8", [&](long long* c) { k_dfma<8><<<sms, threads>>>(dout, 1.0000001, 1e-9, c); }, 8, threads, sms);
        run("FFMA ilp1 (latency)", [&](long long* c) { k_ffma<1><<<sms, threads>>>(fout, 1.0000001f, 1e-9f, c); }, 1, threads, sms);
        run("FFMA ilp8", [&](long long* c) { k_ffma<8><<<sms, threads>>>(fout, 1.0000001f, 1e-9f, c); }, 8, threads, sms);
        run("F2F.F64.F32 ilp8", [&](long long* c) { k_f2d<8><<<sms, threads>>>(dout, 1.0000001f, c); }, 8, threads, sms);
        run("two-float fma ilp4", [&](long long* c) { k_dsfma<4><<<sms, threads>>>(fout, 1.0000001f, 0.999f, c); }, 4, threads, sms);
        run("DMMA m8n8k4 x4 (per mma)", [&](long long* c) { k_dmma<<<sms, threads>>>(dout, c); }, 4, threads, sms);
        run("fp64 1/x ilp4", [&](long long* c) { k_drcp<<<sms, threads>>>(dout, 1e-3, c); }, 4, threads, sms);
    }
    return 0;
}
