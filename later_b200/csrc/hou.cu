// Recursive Householder QR in WY form (reference QR/later_rhouqr.cu:21-201,
// QR/later_bhouqr.cu): A = Q R with Q = I - W Y^T; A <- Y (unit lower trapezoidal), W, R out.
// later_ormqr / later_ormqr2 (ormqr.cu) turn the pair into the explicit Q.
//
// Structure of the reference, kept: halve the columns; factor the left half; A2 <- Q1^T A2 =
// A2 - Y1 (W1^T A2); factor the trailing block A22; R12 = A12, A12 <- 0; W2 <- Q1 W2 = W2 - W1 (Y1^T W2).
// What is different underneath:
//   * the 32-column leaf is the Gram/Cholesky strip factorisation of panel32.cu (explicit Q, R with
//     positive diagonal) instead of the Householder CAQR tree (reference QR/panel.cu:341-558), followed
//     by the reconstruction of the Householder vectors from that Q (Ballard et al., "Reconstructing
//     Householder vectors from TSQR"; the reference does the same with cuSOLVER getrf + two cuBLAS trsm,
//     QR/later_rhouqr.cu:239-277): LU WITHOUT pivoting of S - Q1 gives Y1 = L and U = T Y1^T S, then
//     Y2 = -Q2 U^-1 and W = ([I; 0] - Q S) Y1^-T.  The signs S are chosen during the elimination so
//     that every pivot is >= 1 in magnitude (the reference takes S = I, which is safe for its
//     Householder panel, whose Q1 has a negative diagonal, but not for a Q with positive diagonal);
//     R <- S R accordingly.
//   * the four products per node run on the tcgen05 kernels of tc_gemm.cu on fp16 shadows of Y and W
//     (the reference: cublasGemmEx on separately cast copies for n/2 > 128, fp32 Sgemm below).
#include "../../include/later_b200.h"

#include <algorithm>

#include "context.h"
#include "launch.cuh"
#include "tc_gemm.cuh"

namespace lb {
namespace {

constexpr int HB = 32;        // leaf width (the reference's NMIN, QR/later_rhouqr.cu:7)

inline long round_up(long x, long a) { return (x + a - 1) / a * a; }

// One warp: modified LU of S - Q1 (32 x 32, lane i = row i).  In: Q1 = top block of the strip (in A), the
// strip's R.  Out: A top <- Y1 = L (unit lower, explicit ones and zeros), fac = {L, U, S}, W top <-
// ([I] - Q1 S) L^-T, R <- S R.
__global__ void __launch_bounds__(32)
hou_lu32_kernel(float* __restrict__ A, long lda, float* __restrict__ W, long ldw, float* __restrict__ R, long ldr,
                float* __restrict__ fac) {
    __shared__ float M[HB][HB + 1], L[HB][HB + 1], V[HB][HB + 1];
    __shared__ float S[HB];
    const int i = threadIdx.x;
    for (int j = 0; j < HB; ++j) {
        const float q = A[i + (long)j * lda];
        M[i][j] = -q;
        V[i][j] = q;                       // (turned into [I] - Q1 S once S is known)
        L[i][j] = i == j ? 1.f : 0.f;
    }
    __syncwarp();
    for (int j = 0; j < HB; ++j) {
        if (i == j) {
            const float s = M[j][j] >= 0.f ? 1.f : -1.f;      // |pivot| = |M_jj| + 1 >= 1
            S[j] = s;
            M[j][j] += s;
        }
        __syncwarp();
        const float piv = M[j][j];
        if (i > j) {
            const float l = M[i][j] / piv;
            L[i][j] = l;
            for (int k = j + 1; k < HB; ++k) M[i][k] = fmaf(-l, M[j][k], M[i][k]);
        }
        __syncwarp();
    }
    // U = upper part of M (rows are final once eliminated); W top: row i of ([I] - Q1 S) L^-T
    float w[HB];
    for (int j = 0; j < HB; ++j) w[j] = (i == j ? 1.f : 0.f) - V[i][j] * S[j];
    for (int j = 0; j < HB; ++j)                               // w L^T = v: w_j = v_j - sum_{k<j} w_k L(j,k)
        for (int k = 0; k < j; ++k) w[j] = fmaf(-w[k], L[j][k], w[j]);
    for (int j = 0; j < HB; ++j) {
        W[i + (long)j * ldw] = w[j];
        A[i + (long)j * lda] = L[i][j];
        fac[i + j * HB] = L[i][j];                             // column-major 32 x 32
        fac[HB * HB + i + j * HB] = i <= j ? M[i][j] : 0.f;    // U
        R[i + (long)j * ldr] = i <= j ? S[i] * R[i + (long)j * ldr] : 0.f;
    }
    fac[2 * HB * HB + i] = S[i];
}

// Rows below the top block, one per thread: W row = (-q S) L^-T, Y row = (-q) U^-1.
__global__ void __launch_bounds__(128)
hou_rows32_kernel(float* __restrict__ A, long lda, float* __restrict__ W, long ldw, int rows,
                  const float* __restrict__ fac) {
    __shared__ float L[HB][HB + 1], U[HB][HB + 1], S[HB], uinv[HB];
    for (int e = threadIdx.x; e < HB * HB; e += blockDim.x) {
        L[e % HB][e / HB] = fac[e];
        U[e % HB][e / HB] = fac[HB * HB + e];
    }
    if (threadIdx.x < HB) S[threadIdx.x] = fac[2 * HB * HB + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < HB) uinv[threadIdx.x] = 1.f / U[threadIdx.x][threadIdx.x];
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float q[HB], w[HB];
#pragma unroll
    for (int j = 0; j < HB; ++j) {
        q[j] = -A[r + (long)j * lda];
        w[j] = q[j] * S[j];
    }
#pragma unroll
    for (int j = 0; j < HB; ++j) {
#pragma unroll
        for (int k = 0; k < j; ++k) w[j] = fmaf(-w[k], L[j][k], w[j]);     // w L^T = -q S
    }
#pragma unroll
    for (int j = 0; j < HB; ++j) {                                         // y U = -q
#pragma unroll
        for (int k = 0; k < j; ++k) q[j] = fmaf(-q[k], U[k][j], q[j]);
        q[j] *= uinv[j];
    }
#pragma unroll
    for (int j = 0; j < HB; ++j) {
        W[r + (long)j * ldw] = w[j];
        A[r + (long)j * lda] = q[j];
    }
}

// D(i, j) = fp16(S(i, j)) over a rows x cols block (column-major both).
__global__ void cast_block_kernel(const float* __restrict__ S, long lds, int rows, int cols, __half* __restrict__ D,
                                  long ldd) {
    const long total = (long)rows * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        D[i + (long)j * ldd] = __float2half_rn(S[i + (long)j * lds]);
    }
}

// R12 <- A12, A12 <- 0 (fp32 and its fp16 shadow).
__global__ void extract_r12_kernel(float* __restrict__ A12, long lda, __half* __restrict__ Ah, long ldh, int h, int nb,
                                   float* __restrict__ R12, long ldr) {
    const long total = (long)h * nb;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % h), j = (int)(idx / h);
        R12[i + (long)j * ldr] = A12[i + (long)j * lda];
        A12[i + (long)j * lda] = 0.f;
        Ah[i + (long)j * ldh] = __float2half_rn(0.f);
    }
}

inline int ew_grid(long total) { return (int)std::min<long>((total + 255) / 256, 148L * 16); }

struct Hou {
    later_b200_ctx* ctx;
    int m, n;
    float* A; long lda;
    float* W; long ldw;
    float* R; long ldr;
    __half* Sh; long ldh;      // fp16 shadows: columns [0, n) of A / Y, [n, 2n) of W
    float* T; __half* Th;      // node product (h x h), fp32 and fp16
    float* part; float* fac;
    void* split_scratch;       // planes of the fp32-faithful products
    bool all_fp32;             // later_bhouqr's arithmetic: every product fp32 (reference QR/later_bhouqr.cu)
    CUtensorMap q128, q64;
    cudaError_t err = cudaSuccess;
    int rc = 0;
    long launches = 0;

    void check(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }
    bool ok() const { return err == cudaSuccess && rc == 0; }

    void cast(const float* S, long lds, int rows, int cols, __half* D) {
        cast_block_kernel<<<ew_grid((long)rows * cols), 256, 0, ctx->stream>>>(S, lds, rows, cols, D, ldh);
        launches += 1;
    }

    // C2 -= X1 (Z1^T C2) over rows c0 .. m - 1: X1, Z1 = columns [c0, c0 + h) of the shadow parts xs / zs,
    // C2 = columns [c0 + h, c0 + 2h) of C (fp32, ldc), whose shadow part is cs.
    void project(int c0, int h, int zs, int xs, float* C, long ldc, int cs) {
        if (!ok()) return;
        cudaStream_t st = ctx->stream;
        const int rows = m - c0, cb = c0 + h;
        // The reference multiplies in fp32 where n/2 <= 128 (QR/later_rhouqr.cu:83) and everywhere in
        // later_bhouqr: those products run as fp32-faithful split-precision triples here; the others
        // on fp16 operands as the reference's cublasGemmEx calls do (:106-137, :185-213).
        if (h <= 128 || all_fp32) {
            const float* Z = (zs == 0 ? A : W) + c0 + (long)c0 * (zs == 0 ? lda : ldw);
            const float* X = (xs == 0 ? A : W) + c0 + (long)c0 * (xs == 0 ? lda : ldw);
            float* C2 = C + c0 + (long)cb * ldc;
            rc = split_project(ctx, rows, h, h, Z, zs == 0 ? lda : ldw, X, xs == 0 ? lda : ldw, C2, ldc, split_scratch,
                               &launches);
            if (rc == 0) cast(C2, ldc, rows, h, Sh + c0 + (long)(cs + cb) * ldh);     // its fp16 shadow
            return;
        }
        const int splits = choose_gram_splits(ctx->num_sms, h, h, 128, rows);
        check(tc_gram(st, ctx->num_sms, q128, q128, 128, c0, rows, zs + c0, h, cs + cb, h, T, h, Th, h, part, splits));
        CUtensorMap tmap;
        HalfMatrix tm{Th, h, h, h};
        check(make_tensor_map_f16(&tmap, tm, 64, 128));
        check(tc_update(st, ctx->num_sms, q64, tmap, 128, c0, rows, xs + c0, h, 0, h, C + c0 + (long)cb * ldc, ldc,
                        Sh + c0 + (long)(cs + cb) * ldh, ldh, true));
        launches += splits > 1 ? 3 : 2;
    }

    void leaf(int c0) {
        if (!ok()) return;
        const int rows = m - c0;
        float* As = A + c0 + (long)c0 * lda;
        float* Ws = W + c0 + (long)c0 * ldw;
        float* Rs = R + c0 + (long)c0 * ldr;
        if ((rc = later_b200_panel32_qr(ctx, rows, HB, As, (int)lda, Rs, (int)ldr)) != 0) return;
        launches += ctx->launches;
        hou_lu32_kernel<<<1, 32, 0, ctx->stream>>>(As, lda, Ws, ldw, Rs, ldr, fac);
        if (rows > HB)
            hou_rows32_kernel<<<(rows - HB + 127) / 128, 128, 0, ctx->stream>>>(As + HB, lda, Ws + HB, ldw, rows - HB, fac);
        launches += 2;
        cast(As, lda, rows, HB, Sh + c0 + (long)c0 * ldh);
        cast(Ws, ldw, rows, HB, Sh + c0 + (long)(n + c0) * ldh);
    }

    void qr(int c0, int w, bool merge_w) {
        if (!ok()) return;
        if (w <= HB) { leaf(c0); return; }
        const int h = w / 2;
        qr(c0, h, true);
        project(c0, h, /*Z1 = W1*/ n, /*X1 = Y1*/ 0, A, lda, /*C2 = A2*/ 0);          // A2 <- Q1^T A2
        qr(c0 + h, h, true);
        if (!ok()) return;
        extract_r12_kernel<<<ew_grid((long)h * h), 256, 0, ctx->stream>>>(
            A + c0 + (long)(c0 + h) * lda, lda, Sh + c0 + (long)(c0 + h) * ldh, ldh, h, h,
            R + c0 + (long)(c0 + h) * ldr, ldr);
        launches += 1;
        if (merge_w) project(c0, h, /*Z1 = Y1*/ 0, /*X1 = W1*/ n, W, ldw, /*C2 = W2*/ n);  // W2 <- Q1 W2
    }
};

}  // namespace
}  // namespace lb

using namespace lb;

extern "C" int later_b200_rhouqr(later_b200_ctx* ctx, int m, int n, float* A, int lda, float* W, int ldw, float* R,
                                 int ldr, int merge_top) {
    if (!ctx) return LATER_B200_EINVAL;
    if (!A || !W || !R) return fail(ctx, LATER_B200_EINVAL, "null matrix pointer");
    if (n < HB || n % HB != 0 || ((n / HB) & (n / HB - 1)) != 0) return fail(ctx, LATER_B200_EINVAL, "n must be 32 * 2^k");
    if (m < n || m % 8 != 0) return fail(ctx, LATER_B200_EINVAL, "m must be >= n and a multiple of 8");
    if (lda < m || ldw < m || ldr < n) return fail(ctx, LATER_B200_EINVAL, "leading dimension too small");
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");

    Hou hq{};
    hq.ctx = ctx; hq.m = m; hq.n = n;
    hq.A = A; hq.lda = lda; hq.W = W; hq.ldw = ldw; hq.R = R; hq.ldr = ldr;
    hq.ldh = round_up(m, 8);
    const int hmax = std::max(HB, n / 2);
    size_t part_floats = 0;
    for (int h = HB; h * 2 <= n; h *= 2) {
        const int s = choose_gram_splits(ctx->num_sms, h, h, 128, m);
        if (s > 1) part_floats = std::max(part_floats, (size_t)s * h * h);
    }
    const size_t sh_bytes = round_up((size_t)hq.ldh * 2 * n * sizeof(__half), 256);
    const size_t t_bytes = round_up((size_t)hmax * hmax * sizeof(float), 256);
    const size_t th_bytes = round_up((size_t)hmax * hmax * sizeof(__half), 256);
    const int h_split = merge_top ? n / 2 : std::min(128, n / 2);     // widest fp32-faithful product
    const size_t split_bytes = n > HB ? round_up(split_project_scratch_bytes(m, h_split, h_split), 256) : 0;
    const size_t need = sh_bytes + t_bytes + th_bytes + round_up(part_floats * sizeof(float), 256) + split_bytes + 16384;
    if (ctx->aux_bytes < need) {
        if (ctx->aux) cudaFree(ctx->aux);
        ctx->aux = nullptr; ctx->aux_bytes = 0;
        if ((e = cudaMalloc(&ctx->aux, need)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc scratch");
        ctx->aux_bytes = need;
    }
    uint8_t* base = static_cast<uint8_t*>(ctx->aux);
    hq.Sh = reinterpret_cast<__half*>(base);
    hq.T = reinterpret_cast<float*>(base + sh_bytes);
    hq.Th = reinterpret_cast<__half*>(base + sh_bytes + t_bytes);
    hq.part = reinterpret_cast<float*>(base + sh_bytes + t_bytes + th_bytes);
    hq.split_scratch = base + sh_bytes + t_bytes + th_bytes + round_up(part_floats * sizeof(float), 256);
    hq.all_fp32 = merge_top != 0;
    hq.fac = reinterpret_cast<float*>(base + need - 16384);
    HalfMatrix sm{hq.Sh, m, 2 * n, hq.ldh};
    if ((e = make_tensor_map_f16(&hq.q128, sm, 64, 128)) != cudaSuccess ||
        (e = make_tensor_map_f16(&hq.q64, sm, 64, 64)) != cudaSuccess)
        return cuda_fail(ctx, e, "tensor map encode");
    // W's blocks above the block diagonal are never produced (the reference relies on zero-initialised
    // memory, test/test_qr.cu:109); R's strictly lower triangle reads as zero
    cudaStream_t st = ctx->stream;
    if ((e = cudaMemset2DAsync(W, (size_t)ldw * sizeof(float), 0, (size_t)m * sizeof(float), n, st)) != cudaSuccess ||
        (e = cudaMemset2DAsync(R, (size_t)ldr * sizeof(float), 0, (size_t)n * sizeof(float), n, st)) != cudaSuccess ||
        (e = cudaMemsetAsync(hq.Sh, 0, sh_bytes, st)) != cudaSuccess)
        return cuda_fail(ctx, e, "clear");
    hq.cast(A, lda, m, n, hq.Sh);                 // the trailing columns enter the products as they are
    hq.qr(0, n, merge_top != 0);
    ctx->launches = hq.launches;
    ctx->plan.valid = false;
    if (hq.rc) return hq.rc;
    if (hq.err != cudaSuccess) return cuda_fail(ctx, hq.err, "rhouqr");
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(ctx, e, "rhouqr launch");
    return 0;
}
