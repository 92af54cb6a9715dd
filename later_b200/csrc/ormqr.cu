// later_ormqr / later_ormqr2: explicit Q = I - W Y^T from a Householder WY pair
// (reference QR/later_ormqr.cu:18-85).
//
// The reference does this with three fp32 cuBLAS products, an m x n cudaMalloc'd temporary, a
// device-to-device copy and a cuBLAS handle created and destroyed per call.  Here the same three
// products run on the tcgen05 kernels of tc_gemm.cu in split precision so the result stays
// fp32-faithful: every operand x is scaled by a power of two s (max|x| s in [2^13, 2^14)) and
// split as x s = hi + lo with hi = fp16(x s), lo = fp16(x s - hi)  (|x s - hi - lo| <= 2^-22 max);
// a product A B is accumulated as  Ahi Bhi + Ahi Blo + Alo Bhi  in fp32 (fp16 x fp16 products are
// exact in fp32), unscaled in the epilogue.  The dropped Alo Blo term is O(2^-22) relative.
// W is overwritten in place from the epilogue (the A operand is read from its fp16 planes), so
// there is no temporary and no copy.
#include "../../include/later_b200.h"

#include <algorithm>
#include <cstdlib>

#include "context.h"
#include "tc_gemm.cuh"

namespace lb {
namespace {

inline long round_up(long x, long a) { return (x + a - 1) / a * a; }
constexpr int BK = 64;   // k-block of the tcgen05 kernels (tc_gemm.cu)

__global__ void maxabs_kernel(const float* __restrict__ X, long ld, int rows, int cols,
                              unsigned* __restrict__ slot) {
    float mx = 0.f;
    const long total = (long)rows * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        mx = fmaxf(mx, fabsf(X[i + (long)j * ld]));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(mx));  // order-preserving for >= 0
}

// scales[0] = s (power of two bringing max|x| into [2^13, 2^14)); scales[1] = 1 / s.
__global__ void make_scale_kernel(unsigned* slot, float* scales) {
    const float mx = __uint_as_float(*slot);
    float s = 1.f;
    if (mx > 0.f && isfinite(mx)) {
        int e;
        frexpf(mx, &e);            // mx = f * 2^e, f in [0.5, 1)
        s = ldexpf(1.f, 14 - e);   // mx * s in [2^13, 2^14)
    }
    scales[0] = s;
    scales[1] = 1.f / s;
    *slot = 0u;  // ready for the next use
}

// unscale[0] = 1 / (sa * sb)
__global__ void combine_scale_kernel(const float* sa, const float* sb, float* out) {
    out[0] = sa[1] * sb[1];
}

// hi/lo fp16 planes of s * X.  transpose=1 writes plane(j, i) = X(i, j).
__global__ void split_kernel(const float* __restrict__ X, long ld, int rows, int cols,
                             const float* __restrict__ scales, __half* __restrict__ hi,
                             __half* __restrict__ lo, long ldp, int transpose) {
    const float s = scales[0];
    const long total = (long)rows * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        const float v = X[i + (long)j * ld] * s;
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const long o = transpose ? (j + (long)i * ldp) : (i + (long)j * ldp);
        hi[o] = h;
        lo[o] = l;
    }
}

__global__ void eye_kernel(float* __restrict__ A, long lda, int rows, int cols) {
    const long total = (long)rows * cols;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        A[i + (long)j * lda] = (i == j) ? 1.f : 0.f;
    }
}

inline int ew_grid(long total) { return (int)std::min<long>((total + 255) / 256, 148L * 16); }

struct Planes {
    __half* hi;
    __half* lo;
    long ld;
};

struct Ormqr {
    int kChunk = 2048;
    later_b200_ctx* ctx;
    cudaStream_t st;
    cudaError_t err = cudaSuccess;
    long launches = 0;
    unsigned* slot;
    float* sW; float* sY; float* sK; float* unscale;  // device scalars (2 floats each)

    void check(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }

    void scale_of(const float* X, long ld, int rows, int cols, float* scales) {
        maxabs_kernel<<<ew_grid((long)rows * cols), 256, 0, st>>>(X, ld, rows, cols, slot);
        make_scale_kernel<<<1, 1, 0, st>>>(slot, scales);
        launches += 2;
    }
    void split(const float* X, long ld, int rows, int cols, const float* scales, Planes p,
               bool transpose) {
        split_kernel<<<ew_grid((long)rows * cols), 256, 0, st>>>(X, ld, rows, cols, scales, p.hi,
                                                                 p.lo, p.ld, transpose ? 1 : 0);
        launches += 1;
    }
    // C (op)= unscale * (Ahi Bhi + Ahi Blo + Alo Bhi) with the operand layouts of `a_mn_major`.
    void product3(bool a_mn_major, int bn, const HalfMatrix& Ahi, const HalfMatrix& Alo,
                  const HalfMatrix& Bhi, const HalfMatrix& Blo, TcGemmParams p, int first_epi,
                  int next_epi) {
        CUtensorMap ah, al, bh, bl;
        const int a_box_outer = a_mn_major ? 64 : 128;
        check(make_tensor_map_f16(&ah, Ahi, 64, a_box_outer));
        check(make_tensor_map_f16(&al, Alo, 64, a_box_outer));
        check(make_tensor_map_f16(&bh, Bhi, 64, bn));
        check(make_tensor_map_f16(&bl, Blo, 64, bn));
        if (err != cudaSuccess) return;
        p.dscale = unscale;
        // The tensor core adds into its fp32 accumulator with truncation, so the error of one long
        // accumulation grows linearly with the number of MMA steps (measured at K = 32768: 1.1e-5
        // of max|Q| in one piece, 4.8e-6 / 2.3e-6 / 1.2e-6 in chunks of 8192 / 4096 / 2048, against
        // 1.0e-6 for the reference's fp32 FMA chain).  The hi*hi term is therefore cut into chunks
        // of kChunk along K; each chunk accumulates in TMEM and is added to C in fp32
        // round-to-nearest by the epilogue.  The two cross terms are 2^-11 smaller, and so is their
        // truncation error: they run over the whole K.
        const int kb_all = p.kb_total;
        const int kb_chunk = kChunk / BK;
        for (int kb0 = 0; kb0 < kb_all; kb0 += kb_chunk) {
            TcGemmParams q = p;
            q.kb_total = q.kb_per_split = std::min(kb_chunk, kb_all - kb0);
            if (a_mn_major) q.a_c1 += kb0 * BK; else q.a_c0 += kb0 * BK;
            q.b_c0 += kb0 * BK;
            check(tc_gemm_launch(st, ctx->num_sms, a_mn_major, bn, kb0 == 0 ? first_epi : next_epi, ah,
                                 bh, q));
            launches += 1;
        }
        check(tc_gemm_launch(st, ctx->num_sms, a_mn_major, bn, next_epi, ah, bl, p));
        check(tc_gemm_launch(st, ctx->num_sms, a_mn_major, bn, next_epi, al, bh, p));
        launches += 2;
    }
};

int ormqr_impl(later_b200_ctx* ctx, int m, int n, float* W, int ldw, const float* Y, int ldy,
               bool merge_step) {
    if (!ctx) return LATER_B200_EINVAL;
    if (!W || !Y) return fail(ctx, LATER_B200_EINVAL, "null matrix pointer");
    if (n <= 0 || n % 128 != 0) return fail(ctx, LATER_B200_EINVAL, "n must be a multiple of 128");
    if (merge_step && n % 256 != 0)
        return fail(ctx, LATER_B200_EINVAL, "later_ormqr needs n to be a multiple of 256");
    if (m < n || m % 8 != 0) return fail(ctx, LATER_B200_EINVAL, "m must be >= n and a multiple of 8");
    if (ldw < m || ldy < m) return fail(ctx, LATER_B200_EINVAL, "leading dimension too small");
    DeviceGuard guard(ctx->device);
    cudaError_t e = guard.error();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");

    const long ldp = round_up(m, 8);
    const int h = n / 2;
    const size_t plane = (size_t)ldp * n * sizeof(__half);
    const size_t yt_plane = (size_t)n * n * sizeof(__half);
    const size_t k_plane = (size_t)h * h * sizeof(__half);
    const size_t total = 4 * round_up(plane, 256) + 2 * round_up(yt_plane, 256) +
                         2 * round_up(k_plane, 256) + round_up((size_t)h * h * sizeof(float), 256) +
                         8192;
    if ((e = ctx->arena.reserve(total)) != cudaSuccess) {
        cuda_fail(ctx, e, "workspace reserve");
        return LATER_B200_ENOMEM;
    }
    ctx->arena.reset();
    ctx->plan.valid = false;
    Planes Wp{(__half*)ctx->arena.alloc(plane), (__half*)ctx->arena.alloc(plane), ldp};
    Planes Yp{(__half*)ctx->arena.alloc(plane), (__half*)ctx->arena.alloc(plane), ldp};
    Planes Ytp{(__half*)ctx->arena.alloc(yt_plane), (__half*)ctx->arena.alloc(yt_plane), n};
    Planes Kp{(__half*)ctx->arena.alloc(k_plane), (__half*)ctx->arena.alloc(k_plane), h};
    float* work = (float*)ctx->arena.alloc((size_t)h * h * sizeof(float));
    float* scal = (float*)ctx->arena.alloc(64 * sizeof(float));
    if (!Wp.hi || !Wp.lo || !Yp.hi || !Yp.lo || !Ytp.hi || !Ytp.lo || !Kp.hi || !Kp.lo || !work ||
        !scal)
        return fail(ctx, LATER_B200_ENOMEM, "workspace carve failed");

    Ormqr o{};
    o.ctx = ctx; o.st = ctx->stream;
    o.kChunk = ctx->opts.ormqr_kchunk;
    o.slot = reinterpret_cast<unsigned*>(scal + 32);
    o.sW = scal; o.sY = scal + 2; o.sK = scal + 4; o.unscale = scal + 6;
    o.check(cudaMemsetAsync(scal, 0, 64 * sizeof(float), o.st));

    const int bn_n = n >= 256 ? 256 : 128;
    auto hm = [](const __half* p, int rows, int cols, long ld) { return HalfMatrix{p, rows, cols, ld}; };

    if (merge_step) {
        // (i) work = Y1^T W2 ; W2 -= W1 work        (reference QR/later_ormqr.cu:27-45)
        const int bn_h = h >= 256 ? 256 : 128;
        o.scale_of(W, ldw, m, n, o.sW);
        o.scale_of(Y, ldy, m, h, o.sY);
        o.split(W, ldw, m, n, o.sW, Wp, false);
        o.split(Y, ldy, m, h, o.sY, Yp, false);
        combine_scale_kernel<<<1, 1, 0, o.st>>>(o.sY, o.sW, o.unscale);
        TcGemmParams p;
        tc_fill_gram(p, bn_h, 0, m, 0, h, h, h, work, h, nullptr, 0);
        // A operand = Y1 planes (columns 0..h of Y), B operand = W2 planes (columns h..n of W)
        o.product3(false, bn_h, hm(Yp.hi, m, h, ldp), hm(Yp.lo, m, h, ldp), hm(Wp.hi, m, n, ldp),
                   hm(Wp.lo, m, n, ldp), p, EPI_STORE, EPI_ADD);
        o.scale_of(work, h, h, h, o.sK);
        o.split(work, h, h, h, o.sK, Kp, false);
        combine_scale_kernel<<<1, 1, 0, o.st>>>(o.sW, o.sK, o.unscale);
        tc_fill_update(p, bn_h, 0, m, 0, h, 0, h, W + (long)h * ldw, ldw, nullptr, 0);
        o.product3(true, bn_h, hm(Wp.hi, m, n, ldp), hm(Wp.lo, m, n, ldp), hm(Kp.hi, h, h, h),
                   hm(Kp.lo, h, h, h), p, EPI_SUB, EPI_SUB);
        o.launches += 2;
    }
    // (ii) W <- I - W Y(0:n, 0:n)^T                 (reference QR/later_ormqr.cu:47-60, :76-84)
    o.scale_of(W, ldw, m, n, o.sW);
    o.scale_of(Y, ldy, n, n, o.sY);
    o.split(W, ldw, m, n, o.sW, Wp, false);
    o.split(Y, ldy, n, n, o.sY, Ytp, true);  // B(k, j) = Y(j, k): K-major planes of Y_n^T
    combine_scale_kernel<<<1, 1, 0, o.st>>>(o.sW, o.sY, o.unscale);
    eye_kernel<<<ew_grid((long)m * n), 256, 0, o.st>>>(W, ldw, m, n);
    o.launches += 2;
    {
        TcGemmParams p;
        tc_fill_update(p, bn_n, 0, m, 0, n, 0, n, W, ldw, nullptr, 0);
        o.product3(true, bn_n, hm(Wp.hi, m, n, ldp), hm(Wp.lo, m, n, ldp), hm(Ytp.hi, n, n, n),
                   hm(Ytp.lo, n, n, n), p, EPI_SUB, EPI_SUB);
    }
    ctx->launches = o.launches;
    if (o.err != cudaSuccess) return cuda_fail(ctx, o.err, "ormqr");
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(ctx, e, "ormqr launch");
    return 0;
}

}  // namespace

// C = A * B with every matrix fp32 column-major (A: M x K, B: K x N), fp32-faithful: the same
// split-precision tcgen05 products as later_ormqr.  scratch must hold split_gemm_scratch_bytes.
size_t split_gemm_scratch_bytes(int M, int N, int K) {
    return 2 * round_up((size_t)round_up(M, 8) * K * sizeof(__half), 256) +
           2 * round_up((size_t)round_up(K, 8) * N * sizeof(__half), 256) + 1024;
}

int split_gemm_nn(later_b200_ctx* ctx, int M, int N, int K, const float* A, long lda, const float* B, long ldb,
                  float* C, long ldc, void* scratch, long* launches) {
    if (M % 8 != 0 || K % 8 != 0) return fail(ctx, LATER_B200_EINVAL, "split_gemm: M and K must be multiples of 8");
    uint8_t* base = static_cast<uint8_t*>(scratch);
    const long ldpa = round_up(M, 8), ldpb = round_up(K, 8);
    const size_t pa = round_up((size_t)ldpa * K * sizeof(__half), 256), pb = round_up((size_t)ldpb * N * sizeof(__half), 256);
    Planes Ap{(__half*)base, (__half*)(base + pa), ldpa};
    Planes Bp{(__half*)(base + 2 * pa), (__half*)(base + 2 * pa + pb), ldpb};
    float* scal = reinterpret_cast<float*>(base + 2 * pa + 2 * pb);
    Ormqr o{};
    o.ctx = ctx; o.st = ctx->stream;
    o.kChunk = ctx->opts.ormqr_kchunk;
    o.slot = reinterpret_cast<unsigned*>(scal + 32);
    o.sW = scal; o.sY = scal + 2; o.sK = scal + 4; o.unscale = scal + 6;
    o.check(cudaMemsetAsync(scal, 0, 64 * sizeof(float), o.st));
    o.scale_of(A, lda, M, K, o.sW);
    o.scale_of(B, ldb, K, N, o.sY);
    o.split(A, lda, M, K, o.sW, Ap, false);
    o.split(B, ldb, K, N, o.sY, Bp, false);
    combine_scale_kernel<<<1, 1, 0, o.st>>>(o.sW, o.sY, o.unscale);
    o.launches += 1;
    const int bn = N >= 256 ? 256 : 128;
    TcGemmParams p;
    tc_fill_update(p, bn, 0, M, 0, K, 0, N, C, ldc, nullptr, 0);
    o.product3(true, bn, HalfMatrix{Ap.hi, M, K, ldpa}, HalfMatrix{Ap.lo, M, K, ldpa}, HalfMatrix{Bp.hi, K, N, ldpb},
               HalfMatrix{Bp.lo, K, N, ldpb}, p, EPI_STORE, EPI_ADD);
    if (launches) *launches += o.launches;
    if (o.err != cudaSuccess) return cuda_fail(ctx, o.err, "split_gemm");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "split_gemm launch");
    return 0;
}

// C2 -= X (Z^T C2) with every matrix fp32 column-major and `rows` rows (Z, X: rows x h; C2: rows x nb),
// fp32-faithful: the same two split-precision product triples as the merge step of later_ormqr.
size_t split_project_scratch_bytes(int rows, int h, int nb) {
    const size_t ldp = round_up(rows, 8);
    return 4 * round_up(ldp * h * sizeof(__half), 256) + 2 * round_up(ldp * nb * sizeof(__half), 256) +
           2 * round_up((size_t)round_up(h, 8) * nb * sizeof(__half), 256) + round_up((size_t)h * nb * sizeof(float), 256) +
           1024;
}

int split_project(later_b200_ctx* ctx, int rows, int h, int nb, const float* Z, long ldz, const float* X, long ldx,
                  float* C2, long ldc, void* scratch, long* launches) {
    if (rows % 8 != 0) return fail(ctx, LATER_B200_EINVAL, "split_project: rows must be a multiple of 8");
    uint8_t* base = static_cast<uint8_t*>(scratch);
    const long ldp = round_up(rows, 8), ldk = round_up(h, 8);
    const size_t pz = round_up((size_t)ldp * h * sizeof(__half), 256), pc = round_up((size_t)ldp * nb * sizeof(__half), 256);
    const size_t pk = round_up((size_t)ldk * nb * sizeof(__half), 256);
    Planes Zp{(__half*)base, (__half*)(base + pz), ldp};
    Planes Xp{(__half*)(base + 2 * pz), (__half*)(base + 3 * pz), ldp};
    Planes Cp{(__half*)(base + 4 * pz), (__half*)(base + 4 * pz + pc), ldp};
    Planes Kp{(__half*)(base + 4 * pz + 2 * pc), (__half*)(base + 4 * pz + 2 * pc + pk), ldk};
    float* work = reinterpret_cast<float*>(base + 4 * pz + 2 * pc + 2 * pk);
    float* scal = reinterpret_cast<float*>(base + 4 * pz + 2 * pc + 2 * pk + round_up((size_t)h * nb * sizeof(float), 256));
    Ormqr o{};
    o.ctx = ctx; o.st = ctx->stream;
    o.kChunk = ctx->opts.ormqr_kchunk;
    o.slot = reinterpret_cast<unsigned*>(scal + 32);
    o.sW = scal; o.sY = scal + 2; o.sK = scal + 4; o.unscale = scal + 6;
    o.check(cudaMemsetAsync(scal, 0, 64 * sizeof(float), o.st));
    const int bn = nb >= 256 ? 256 : 128;
    auto hm = [](const __half* p, int r, int c, long ld) { return HalfMatrix{p, r, c, ld}; };
    // work = Z^T C2
    o.scale_of(C2, ldc, rows, nb, o.sW);
    o.scale_of(Z, ldz, rows, h, o.sY);
    o.split(C2, ldc, rows, nb, o.sW, Cp, false);
    o.split(Z, ldz, rows, h, o.sY, Zp, false);
    combine_scale_kernel<<<1, 1, 0, o.st>>>(o.sY, o.sW, o.unscale);
    TcGemmParams p;
    tc_fill_gram(p, bn, 0, rows, 0, h, 0, nb, work, h, nullptr, 0);
    o.product3(false, bn, hm(Zp.hi, rows, h, ldp), hm(Zp.lo, rows, h, ldp), hm(Cp.hi, rows, nb, ldp),
               hm(Cp.lo, rows, nb, ldp), p, EPI_STORE, EPI_ADD);
    // C2 -= X work
    o.scale_of(work, h, h, nb, o.sK);
    o.split(work, h, h, nb, o.sK, Kp, false);
    o.scale_of(X, ldx, rows, h, o.sW);
    o.split(X, ldx, rows, h, o.sW, Xp, false);
    combine_scale_kernel<<<1, 1, 0, o.st>>>(o.sW, o.sK, o.unscale);
    tc_fill_update(p, bn, 0, rows, 0, h, 0, nb, C2, ldc, nullptr, 0);
    o.product3(true, bn, hm(Xp.hi, rows, h, ldp), hm(Xp.lo, rows, h, ldp), hm(Kp.hi, h, nb, ldk),
               hm(Kp.lo, h, nb, ldk), p, EPI_SUB, EPI_SUB);
    o.launches += 2;
    if (launches) *launches += o.launches;
    if (o.err != cudaSuccess) return cuda_fail(ctx, o.err, "split_project");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "split_project launch");
    return 0;
}

}  // namespace lb

extern "C" {

int later_b200_ormqr(later_b200_ctx* ctx, int m, int n, float* W, int ldw, const float* Y, int ldy) {
    return lb::ormqr_impl(ctx, m, n, W, ldw, Y, ldy, true);
}

int later_b200_ormqr2(later_b200_ctx* ctx, int m, int n, float* W, int ldw, const float* Y,
                      int ldy) {
    return lb::ormqr_impl(ctx, m, n, W, ldw, Y, ldy, false);
}

}  // extern "C"
