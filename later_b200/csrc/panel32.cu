// 32-column panel entry points of the reference's QR/include/LATER_QR.h:20-25:
//
//   mgs_kernel2<<<blocks, dim3(32, 32)>>>   QR of every 256-row block of an m x 32 strip on its own
//                                            (reference QR/panel.cu:246-325), block b's R at rows 32 b
//   mgs_caqr_panel_256x32                    QR of a whole m x 32 strip (reference QR/panel.cu:65-134)
//
// Inside RGSQRF these are subsumed by the 128-column Gram/Cholesky panel (panel.cu); they exist so
// that callers of the reference's panel interface (test/test_mgs_panel.cu, test/test_caqr_panel.cu)
// find them.  Same method as the wide panel, in small: Gram matrix with exact fp32 x fp32 products
// summed in fp64, Cholesky in fp64, Q = A R^-1 by forward substitution in fp32 - one pass for the
// Gram matrix and one for Q instead of the reference's CAQR tree (2 depth - 1 launches, each
// reading and writing the strip twice).
#include "../../include/later_b200.h"
#include "../../include/LATER_QR.h"

#include <algorithm>

#include "context.h"
#include "launch.cuh"

namespace lb {
namespace {

constexpr int SW = 32;          // strip width
constexpr int BR = 256;         // rows per block (the reference's block height)
constexpr int LDT = SW + 1;     // padded row of the staged tile: row-wise and column-wise conflict-free

// Stages rows [row0, row0 + mm) x nn columns of A into T[r][c] (zero padded to BR x SW).  Consecutive
// threads move consecutive rows of one column: coalesced global accesses, conflict-free in T.
template <int NT>
__device__ __forceinline__ void tile_load(float (*T)[LDT], const float* __restrict__ A, long lda, long row0,
                                          int mm, int nn, int tid) {
#pragma unroll
    for (int idx = tid; idx < BR * SW; idx += NT) {
        const int r = idx % BR, c = idx / BR;
        T[r][c] = (r < mm && c < nn) ? A[row0 + r + (long)c * lda] : 0.f;
    }
}

// acc += sum_r T[r][y] T[r][x]: exact products, fp64 sums.  (Within a warp T[r][y] is a broadcast and
// T[r][x] hits 32 consecutive banks.)
__device__ __forceinline__ double tile_gram(const float (*T)[LDT], int x, int y, double acc) {
#pragma unroll 8
    for (int r = 0; r < BR; ++r) acc = fma((double)T[r][y], (double)T[r][x], acc);
    return acc;
}

// In-place Cholesky G = R^T R of the leading nn x nn block by ONE warp; lane i owns row i of the
// lower triangle.  Leaves R (upper, fp32) in Rs[k][j], j >= k, zeros elsewhere, and 1 / R(k, k) in
// rinv.  A non-positive pivot is clamped (the reference's MGS divides by the vanishing norm instead,
// reference QR/panel.cu:286-290).
__device__ void chol32_warp(double (*G)[SW + 1], float (*Rs)[SW], float* rinv, int nn, int lane) {
    for (int k = 0; k < nn; ++k) {
        double piv = G[k][k];
        if (!(piv > 0.0)) piv = 1e-300;
        const double rs = rsqrt(piv);
        const double lik = lane >= k ? G[lane][k] * rs : 0.0;      // L(i, k)
        __syncwarp();
        if (lane >= k && lane < nn) G[lane][k] = lik;
        __syncwarp();
        if (lane > k && lane < nn)
            for (int j = k + 1; j <= lane; ++j) G[lane][j] = fma(-lik, G[j][k], G[lane][j]);
        __syncwarp();
        if (lane == k) rinv[k] = (float)rs;
    }
    for (int k = 0; k < SW; ++k)                                   // R(k, j) = L(j, k)
        Rs[k][lane] = (k < nn && lane < nn && lane >= k) ? (float)G[lane][k] : 0.f;
    if (lane >= nn) rinv[lane] = 0.f;
    __syncwarp();
}

// Row r of the tile: q_j = (a_j - sum_{k<j} q_k R(k, j)) / R(j, j), in registers.
__device__ __forceinline__ void row_substitute(float (*T)[LDT], const float (*Rs)[SW], const float* rinv,
                                               int r, int nn) {
    float q[SW];
#pragma unroll
    for (int j = 0; j < SW; ++j) q[j] = T[r][j];
#pragma unroll
    for (int k = 0; k < SW; ++k) {
        const float qk = q[k] * rinv[k];
        q[k] = qk;
#pragma unroll
        for (int j = k + 1; j < SW; ++j) q[j] = fmaf(-qk, Rs[k][j], q[j]);
    }
#pragma unroll
    for (int j = 0; j < SW; ++j) T[r][j] = j < nn ? q[j] : 0.f;
}

template <int NT>
__device__ __forceinline__ void tile_store(const float (*T)[LDT], float* __restrict__ A, long lda, long row0,
                                           int mm, int nn, int tid) {
#pragma unroll
    for (int idx = tid; idx < BR * SW; idx += NT) {
        const int r = idx % BR, c = idx / BR;
        if (r < mm && c < nn) A[row0 + r + (long)c * lda] = T[r][c];
    }
}

struct Shared32 {
    float T[BR][LDT];            // 33 KiB
    double G[SW][SW + 1];
    float Rs[SW][SW];
    float rinv[SW];
};

// Partial Gram matrix of the blocks this CTA owns: part[cta][y][x].
__global__ void __launch_bounds__(1024, 1)
gram32_kernel(const float* __restrict__ A, long lda, int m, int n, double* __restrict__ part) {
    __shared__ float T[BR][LDT];
    const int x = threadIdx.x, y = threadIdx.y;
    double acc = 0.0;
    const int nblocks = (m + BR - 1) / BR;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const long row0 = (long)b * BR;
        tile_load<1024>(T, A, lda, row0, (int)min((long)BR, m - row0), n, y * 32 + x);
        __syncthreads();
        acc = tile_gram(T, x, y, acc);
        __syncthreads();
    }
    part[(long)blockIdx.x * SW * SW + y * SW + x] = acc;
}

// One CTA: fixed-order sum of the partials, Cholesky, R to the caller (zeros below the diagonal)
// and the factors the apply kernel needs.
__global__ void __launch_bounds__(1024, 1)
chol32_kernel(const double* __restrict__ part, int nparts, int n, float* __restrict__ R, long ldr,
              float* __restrict__ fac) {
    __shared__ double G[SW][SW + 1];
    __shared__ float Rs[SW][SW];
    __shared__ float rinv[SW];
    const int x = threadIdx.x, y = threadIdx.y;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += part[(long)c * SW * SW + y * SW + x];
    G[y][x] = s;
    __syncthreads();
    if (y == 0) chol32_warp(G, Rs, rinv, n, x);
    __syncthreads();
    if (x < n && y < n) R[x + (long)y * ldr] = x <= y ? Rs[x][y] : 0.f;
    fac[y * SW + x] = Rs[y][x];
    if (y == 0) fac[SW * SW + x] = rinv[x];
}

__global__ void __launch_bounds__(1024, 1)
apply32_kernel(float* __restrict__ A, long lda, int m, int n, const float* __restrict__ fac) {
    __shared__ float T[BR][LDT];
    __shared__ float Rs[SW][SW];
    __shared__ float rinv[SW];
    const int x = threadIdx.x, y = threadIdx.y, tid = y * 32 + x;
    Rs[y][x] = fac[y * SW + x];
    if (y == 0) rinv[x] = fac[SW * SW + x];
    const int nblocks = (m + BR - 1) / BR;
    for (int b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const long row0 = (long)b * BR;
        const int mm = (int)min((long)BR, m - row0);
        tile_load<1024>(T, A, lda, row0, mm, n, tid);
        __syncthreads();
        if (tid < BR) row_substitute(T, Rs, rinv, tid, n);
        __syncthreads();
        tile_store<1024>(T, A, lda, row0, mm, n, tid);
        __syncthreads();
    }
}

}  // namespace
}  // namespace lb

using namespace lb;

// The QR of one 256-row block by one thread block of NT threads: R_b goes to rows 32 b of RR.
template <int NT>
__device__ __forceinline__ void block_qr32(int m, int n, float* AA, int lda, float* RR, int ldr) {
    __shared__ Shared32 s;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, lane = tid & 31;
    const long row0 = (long)blockIdx.x * BR;
    const int mm = (int)min((long)BR, (long)m - row0);
    if (mm <= 0) return;
    const int nn = min(min(n, SW), mm);          // columns actually factored (the reference's mnmin)
    tile_load<NT>(s.T, AA, lda, row0, mm, nn, tid);
    __syncthreads();
    for (int e = tid; e < SW * SW; e += NT) s.G[e >> 5][e & 31] = tile_gram(s.T, e & 31, e >> 5, 0.0);
    __syncthreads();
    if (tid < 32) chol32_warp(s.G, s.Rs, s.rinv, nn, lane);
    __syncthreads();
    for (int r = tid; r < BR; r += NT) row_substitute(s.T, s.Rs, s.rinv, r, nn);
    __syncthreads();
    tile_store<NT>(s.T, AA, lda, row0, mm, nn, tid);
    for (int e = tid; e < SW * SW; e += NT) {
        const int x = e & 31, y = e >> 5;
        if (x < nn && y < nn) RR[(long)blockIdx.x * SW + x + (long)y * ldr] = x <= y ? s.Rs[x][y] : 0.f;
    }
}

// One block per 256-row block of the strip: the block's own QR.  Launch configurations and argument
// meaning are the reference's: mgs_kernel2<<<blocks, dim3(32, 32)>>> (QR/panel.cu:246-325; launched by
// test/test_mgs_panel.cu:26 and QR/panel.cu:79,92,112) and the older mgs_kernel<<<blocks, 256>>>
// (QR/panel.cu:136-235, test/test_mgs_panel.cu:38).
__global__ void __launch_bounds__(1024) mgs_kernel2(int m, int n, float* AA, int lda, float* RR, int ldr) {
    block_qr32<1024>(m, n, AA, lda, RR, ldr);
}
__global__ void __launch_bounds__(256) mgs_kernel(int m, int n, float* AA, int lda, float* RR, int ldr) {
    block_qr32<256>(m, n, AA, lda, RR, ldr);
}

extern "C" int later_b200_panel32_qr(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* R, int ldr) {
    if (!ctx) return LATER_B200_EINVAL;
    if (n != SW) return fail(ctx, LATER_B200_EINVAL, "panel width must be 32");
    if (!A || !R || m < 1 || lda < m || ldr < std::min(m, n))
        return fail(ctx, LATER_B200_EINVAL, "bad panel arguments");
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    if (m <= BR) {                              // (the reference's recursion leaf, QR/panel.cu:74-79)
        mgs_kernel2<<<1, dim3(32, 32), 0, ctx->stream>>>(m, n, A, lda, R, ldr);
        ctx->launches = 1;
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(ctx, e, "mgs_kernel2");
        return 0;
    }
    const int nblocks = (m + BR - 1) / BR;
    const int grid = std::min(nblocks, ctx->num_sms);
    const size_t bytes = (size_t)grid * SW * SW * sizeof(double) + (SW * SW + SW) * sizeof(float);
    if ((e = ctx->arena.reserve(bytes + 4096)) != cudaSuccess) {
        cuda_fail(ctx, e, "workspace reserve");
        return LATER_B200_ENOMEM;
    }
    ctx->arena.reset();
    ctx->plan.valid = false;
    double* part = static_cast<double*>(ctx->arena.alloc((size_t)grid * SW * SW * sizeof(double)));
    float* fac = static_cast<float*>(ctx->arena.alloc((SW * SW + SW) * sizeof(float)));
    gram32_kernel<<<grid, dim3(32, 32), 0, ctx->stream>>>(A, lda, m, n, part);
    chol32_kernel<<<1, dim3(32, 32), 0, ctx->stream>>>(part, grid, n, R, ldr, fac);
    apply32_kernel<<<grid, dim3(32, 32), 0, ctx->stream>>>(A, lda, m, n, fac);
    ctx->launches = 3;
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(ctx, e, "panel32");
    return 0;
}
