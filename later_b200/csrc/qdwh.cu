// QDWH polar iteration: the reference's in-repo CALLER of later_rgsqrf (reference
// EVD/later_qdwh_polar.cu:24-110).  Every iteration factors the stacked 2n x n matrix
// [sqrt(c) X; I] = [Q1; Q2] R with RGSQRF (reference :79) and forms
//     X <- (a - b/c) / sqrt(c) * Q1 Q2^T + (b/c) X,     X <- (X + X^T) / 2      (reference :94-100)
// with the dynamically weighted Halley coefficients a, b, c of the reference (:60-66).
//
// What is different underneath: the factorisation is the tcgen05 RGSQRF of this library, replayed
// from its cached graph from the third iteration on (same buffers every time); the product Q1 Q2^T
// reads Q1 straight from the fp16 shadow RGSQRF leaves behind (the reference casts both blocks again,
// :88-89) and runs on the tcgen05 GEMM kernel with the scale folded into its epilogue (the reference:
// cublasGemmEx, :94-98); scaling, stacking and the copy of the previous iterate are one elementwise
// pass instead of four (:71, :76, :77, :80); the symmetrisation reads and writes different buffers (the
// reference's generateNewU, :10-21, synchronises only inside a thread block while other blocks
// overwrite the entries it reads).
#include "../../include/later_b200.h"

#include <algorithm>
#include <cmath>
#include <vector>

#include "context.h"
#include "tc_gemm.cuh"

namespace lb {
namespace {

constexpr int kPartials = 592;

__global__ void sumsq_kernel(const float* __restrict__ X, long ldx, int n, double* __restrict__ part) {
    __shared__ double sh[256];
    double s = 0.0;
    const long total = (long)n * n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const double v = X[idx % n + (idx / n) * ldx];
        s += v * v;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

// First iteration (take_from_top = 0): X *= xscale, top = sc * X.  Later ones: X <- top (the iterate the
// previous step left there), top = sc * top.  Always: bottom = I.
__global__ void stack_kernel(float* __restrict__ X, long ldx, float* __restrict__ B, long ldb, int n, float xscale,
                             float sc, int take_from_top) {
    const long total = (long)n * n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        float* top = B + i + (long)j * ldb;
        const float x = take_from_top ? *top : X[i + (long)j * ldx] * xscale;
        X[i + (long)j * ldx] = x;
        *top = x * sc;
        top[n] = i == j ? 1.f : 0.f;
    }
}

// Bt[k + j n] = Qh[(row0 + j) + k ldq]: the K-major operand B(k, j) = Q2(j, k) of Q1 Q2^T.
__global__ void transpose_half_kernel(const __half* __restrict__ Qh, long ldq, int row0, int n, __half* __restrict__ Bt) {
    __shared__ __half t[32][33];
    const int j0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)                 // read: j contiguous
        t[r][threadIdx.x] = Qh[(row0 + j0 + threadIdx.x) + (long)(k0 + r) * ldq];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y)                 // write: k contiguous
        Bt[(k0 + threadIdx.x) + (long)(j0 + r) * n] = t[threadIdx.x][r];
}

__global__ void set_scalar_kernel(float* p, float v) { *p = v; }

// top(i,j) = ((W(i,j) + W(j,i)) + beta (X(i,j) + X(j,i))) / 2; partial sums of (top - X)^2.
__global__ void new_iterate_kernel(const float* __restrict__ W, const float* __restrict__ X, long ldx,
                                   float* __restrict__ B, long ldb, int n, float beta, double* __restrict__ part) {
    __shared__ double sh[256];
    double s = 0.0;
    const long total = (long)n * n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n), j = (int)(idx / n);
        const float xij = X[i + (long)j * ldx], xji = X[j + (long)i * ldx];
        const float v = 0.5f * ((W[i + (long)j * n] + W[j + (long)i * n]) + beta * (xij + xji));
        B[i + (long)j * ldb] = v;
        const double d = (double)v - (double)xij;
        s += d * d;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

}  // namespace
}  // namespace lb

using namespace lb;

extern "C" int later_b200_qdwh_polar(later_b200_ctx* ctx, int n, float* X, int ldx, float* B, int ldb,
                                     float smin_est, int max_iter, int* iters) {
    if (!ctx) return LATER_B200_EINVAL;
    if (!X || !B) return fail(ctx, LATER_B200_EINVAL, "null matrix pointer");
    if (n < 128 || n % 128 != 0 || ((n / 128) & (n / 128 - 1)) != 0)
        return fail(ctx, LATER_B200_EINVAL, "n must be 128 * 2^k");
    if (ldx < n || ldb < 2 * n) return fail(ctx, LATER_B200_EINVAL, "leading dimension too small");
    if (max_iter <= 0) max_iter = 10;                         // reference EVD/later_qdwh_polar.cu:48
    if (!(smin_est > 0.f)) smin_est = 0.0002070391384f;       // the reference's constant (:37)
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = ctx->stream;

    // scratch outside the factorisation's arena: R (discarded), W = Q1 Q2^T, partial sums, one scalar
    const size_t nn = (size_t)n * n;
    const size_t need = 2 * nn * sizeof(float) + kPartials * sizeof(double) + 256;
    if (ctx->aux_bytes < need) {
        if (ctx->aux) cudaFree(ctx->aux);
        ctx->aux = nullptr; ctx->aux_bytes = 0;
        if ((e = cudaMalloc(&ctx->aux, need)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc scratch");
        ctx->aux_bytes = need;
    }
    float* Rw = static_cast<float*>(ctx->aux);
    float* W = Rw + nn;
    double* part = reinterpret_cast<double*>(W + nn);
    float* dscal = reinterpret_cast<float*>(part + kPartials);
    std::vector<double> hpart(kPartials);
    auto reduce_host = [&](double* out) -> int {
        cudaError_t ee = cudaMemcpyAsync(hpart.data(), part, kPartials * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (ee == cudaSuccess) ee = cudaStreamSynchronize(st);
        if (ee != cudaSuccess) return cuda_fail(ctx, ee, "qdwh reduction");
        double s = 0.0;
        for (int i = 0; i < kPartials; ++i) s += hpart[i];     // fixed order
        *out = s;
        return 0;
    };
    const int ew_grid = (int)std::min<size_t>((nn + 255) / 256, 148 * 16);

    // alpha = 1 / ||X||_F  (reference :26-31)
    sumsq_kernel<<<kPartials, 256, 0, st>>>(X, ldx, n, part);
    double ss = 0.0;
    int rc = reduce_host(&ss);
    if (rc) return rc;
    if (!(ss > 0.0) || !std::isfinite(ss)) return fail(ctx, LATER_B200_EINVAL, "qdwh: zero or non-finite input");
    const float alpha = (float)(1.0 / std::sqrt(ss));

    float L = smin_est / std::sqrt((float)n);                 // (:38)
    const float eps = 2e-4f;                                  // (:7)
    const float tol1 = 10.f * eps / 2.f, tol3 = std::pow(tol1, 1.0f / 3.0f);
    int it = 0;
    long launches = 0;
    for (; it < max_iter; ++it) {
        if (it > 0) {
            double d2 = 0.0;
            if ((rc = reduce_host(&d2)) != 0) return rc;      // ||X_k - X_{k-1}||_F (:53-57)
            if (std::sqrt(d2) < tol3 && 1.0f - L < tol1) break;
        }
        // dynamically weighted Halley coefficients (:60-66)
        const float L2 = L * L;
        const float dd = std::pow(4.0f * (1 - L2) / (L2 * L2), 1.0f / 3.0f);
        const float sqd = std::sqrt(1 + dd);
        const float a = sqd + std::sqrt(8 - 4 * dd + 8 * (2 - L2) / (L2 * sqd)) / 2;
        const float b = (a - 1.0f) * (a - 1.0f) / 4.0f;
        const float c = a + b - 1.0f;
        // (clamped: L is a lower bound of the smallest singular value of the iterate, never above 1; in
        // fp32 the update can land on 1 + ulp, which turns the next dd into NaN - the reference's
        // iteration dies that way from its fifth step on when its stopping test does not fire)
        L = std::fmin(L * (a + b * L2) / (1.0f + c * L2), 1.0f);
        const float sqrtc = std::sqrt(c);

        stack_kernel<<<ew_grid, 256, 0, st>>>(X, ldx, B, ldb, n, it == 0 ? alpha : 1.f, sqrtc, it > 0 ? 1 : 0);
        if ((rc = later_b200_rgsqrf(ctx, 2 * n, n, B, ldb, Rw, n)) != 0) return rc;      // (:79)
        launches += ctx->launches;
        auto& p = ctx->plan;
        transpose_half_kernel<<<dim3(n / 32, n / 32), dim3(32, 8), 0, st>>>(p.Qh, p.ldh, n, n, p.Wh);
        set_scalar_kernel<<<1, 1, 0, st>>>(dscal, (a - b / c) / sqrtc);
        // W = (a - b/c) / sqrt(c) * Q1 Q2^T on the tcgen05 kernel: A operand = rows 0 .. n-1 of the shadow
        // (MN-major), B operand = the transposed copy of Q2 (K-major), scale in the epilogue
        CUtensorMap qmap, bmap;
        const int bn = n >= 256 ? 256 : 128;
        HalfMatrix qm{p.Qh, 2 * n, n, p.ldh}, bm{p.Wh, n, n, n};
        if ((e = make_tensor_map_f16(&qmap, qm, 64, 64)) != cudaSuccess ||
            (e = make_tensor_map_f16(&bmap, bm, 64, bn)) != cudaSuccess)
            return cuda_fail(ctx, e, "tensor map encode");
        TcGemmParams g;
        tc_fill_update(g, bn, 0, n, 0, n, 0, n, W, n, nullptr, 0);
        g.dscale = dscal;
        if ((e = tc_gemm_launch(st, ctx->num_sms, true, bn, EPI_STORE, qmap, bmap, g)) != cudaSuccess)
            return cuda_fail(ctx, e, "qdwh gemm");
        new_iterate_kernel<<<kPartials, 256, 0, st>>>(W, X, ldx, B, ldb, n, b / c, part);
        launches += 5;
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(ctx, e, "qdwh launch");
    }
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(ctx, e, "sync");
    ctx->launches = launches;
    if (iters) *iters = it;
    return 0;
}
