// DMMA (mma.sync.m8n8k4.f64) throughput on B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_tput(double* out, long long* cyc, double x) {
    double c[8][2];
    for (int j = 0; j < 8; ++j) { c[j][0] = x + j; c[j][1] = x - j; }
    double a = x * 0.5 + threadIdx.x, b = x * 0.25;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < 512; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 1 << 24); cudaMalloc(&c, 4096 * 8);
    long long h[4096];
    for (int warps : {1, 4, 8, 16}) {
        dmma_tput<<<148, warps * 32>>>(d, c, 1.0000001);
        cudaDeviceSynchronize(); cudaMemcpy(h, c, 148 * 8, cudaMemcpyDeviceToHost);
        long long m = 0; for (int i = 0; i < 148; ++i) m = h[i] > m ? h[i] : m;
        printf("DMMA m8n8k4 %2d warps: %.1f DFMA-equiv/clk/SM (%.2f cyc per mma per warp-slot)\n", warps,
               512.0 * 8 * warps * 256 / m, (double)m / (512.0 * 8));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
