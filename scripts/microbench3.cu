// FFMA vs FFMA2 (fma.rn.f32x2) throughput on B200, alone and mixed with shared-memory loads.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0: FFMA, 1: FFMA2, 2: FFMA + LDS.128 per 8 FMA, 3: FFMA2 + LDS.128 per 8 FMA
__global__ void k(float* out, long long* cyc, float x) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(x, x * 0.5f, x * 0.25f, 1.f);
    float2 acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = make_float2(x + j, x - j);
    float2 a = make_float2(x * 0.5f + threadIdx.x, x * 0.5f + threadIdx.x);
    float4 b = make_float4(x, x * 0.5f, x * 0.25f, 1.f);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (MODE >= 2) b = sm[(threadIdx.x + 32 * g + i) & 1023];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = g * 4 + j;
                const float2 bb = (j & 1) ? make_float2(b.z, b.w) : make_float2(b.x, b.y);
                if (MODE & 1) acc[idx] = __ffma2_rn(a, bb, acc[idx]);
                else { acc[idx].x = fmaf(a.x, bb.x, acc[idx].x); acc[idx].y = fmaf(a.y, bb.y, acc[idx].y); }
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < 16; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, float* d, long long* c) {
    long long h[148];
    for (int warps : {4, 8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(d, c, 1.0000001f);
        cudaDeviceSynchronize(); cudaMemcpy(h, c, 148 * 8, cudaMemcpyDeviceToHost);
        long long m = 0; for (int i = 0; i < 148; ++i) m = h[i] > m ? h[i] : m;
        printf("%-28s %2d warps: %.1f FMA/clk/SM\n", name, warps, 256.0 * 32 * warps * 32 / m);
    }
}
int main() {
    float* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 4096 * 8);
    run<0>("FFMA", d, c); run<1>("FFMA2", d, c);
    run<2>("FFMA + LDS.128 per 8 FMA", d, c); run<3>("FFMA2 + LDS.128 per 8 FMA", d, c);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
