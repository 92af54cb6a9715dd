// Latency / throughput probes for the panel design (fp64 pipe, conversions, barriers) on B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
__global__ void dfma_lat(double* out, long long* cyc, double x) {
    double a = x, b = x * 0.5, c = 1.0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) c = fma(a, c, b);
    long long t1 = clock64();
    out[threadIdx.x] = c; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void ffma_lat(float* out, long long* cyc, float x) {
    float a = x, b = x * 0.5f, c = 1.0f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) c = fmaf(a, c, b);
    long long t1 = clock64();
    out[threadIdx.x] = c; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// throughput: 8 independent chains per thread, many warps
__global__ void dfma_tput(double* out, long long* cyc, double x) {
    double c[8]; for (int j = 0; j < 8; ++j) c[j] = x + j;
    double a = x, b = 0.25;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < N_IT / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = fma(a, c[j], b);
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void ffma_tput(float* out, long long* cyc, float x) {
    float c[8]; for (int j = 0; j < 8; ++j) c[j] = x + j;
    float a = x, b = 0.25f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < N_IT / 8; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = fmaf(a, c[j], b);
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 8; ++j) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void cvt_tput(double* out, long long* cyc, float x) {
    float f[8]; for (int j = 0; j < 8; ++j) f[j] = x + j;
    double s = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < N_IT / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { double d = (double)f[j]; f[j] = __int_as_float(__float_as_int(f[j]) + 1); s += d; }
    }
    __syncthreads();
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void rsqrt_chain(double* out, long long* cyc, double x) {
    double a = x;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) {
        double y = (double)rsqrtf((float)a);
        y = y * (1.5 - 0.5 * a * y * y);
        a = a * y + 1.0;   // dependent
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void bar_cost(long long* cyc) {
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void shfl_lat(double* out, long long* cyc, double x) {
    double a = x + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) a = __shfl_sync(0xffffffffu, a, (i * 7) & 31) + 1.0;
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* d; float* f; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&f, 1 << 24); cudaMalloc(&c, 4096 * 8);
    long long h[4096];
    auto rd = [&](int n) { cudaDeviceSynchronize(); cudaMemcpy(h, c, n * 8, cudaMemcpyDeviceToHost); long long m = 0; for (int i = 0; i < n; ++i) m = h[i] > m ? h[i] : m; return m; };
    for (int rep = 0; rep < 2; ++rep) {
    dfma_lat<<<1, 32>>>(d, c, 1.0000001); printf("DFMA latency      %.2f cyc\n", rd(1) / (double)N_IT);
    ffma_lat<<<1, 32>>>(f, c, 1.0000001f); printf("FFMA latency      %.2f cyc\n", rd(1) / (double)N_IT);
    for (int warps : {4, 8, 16, 32}) {
        dfma_tput<<<148, warps * 32>>>(d, c, 1.0000001);
        long long cy = rd(148); printf("DFMA tput %2d warps: %.1f DFMA/clk/SM\n", warps, (double)N_IT * 8 * warps * 32 / cy);
        ffma_tput<<<148, warps * 32>>>(f, c, 1.0000001f);
        cy = rd(148); printf("FFMA tput %2d warps: %.1f FFMA/clk/SM\n", warps, (double)N_IT * 8 * warps * 32 / cy);
    }
    cvt_tput<<<148, 512>>>(d, c, 1.5f); { long long cy = rd(148); printf("F2F.F64.F32+DADD tput 16 warps: %.1f /clk/SM\n", (double)N_IT * 512 / cy); }
    rsqrt_chain<<<1, 32>>>(d, c, 2.0); printf("rsqrt(f32 seed)+1NR+dep DFMA chain: %.1f cyc/iter\n", rd(1) / 256.0);
    for (int t : {128, 256, 512, 1024}) { bar_cost<<<1, t>>>(c); printf("__syncthreads %4d thr: %.1f cyc\n", t, rd(1) / 256.0); }
    shfl_lat<<<1, 32>>>(d, c, 1.0); printf("shfl(double)+DADD dep: %.1f cyc\n", rd(1) / 256.0);
    }
    return 0;
}
