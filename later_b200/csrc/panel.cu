// 128-column tall-skinny panel QR (see panel.cuh).  Four kernels per panel:
//   gram128_f64_kernel    one pass over the panel: all column norms and dot products, exact
//                         fp32 x fp32 products accumulated in fp64, per-CTA partial Gram blocks
//   gram128_reduce_kernel fixed-order sum of the per-CTA partials (deterministic)
//   chol128_kernel        R = chol(G) in fp64 (one CTA, register-resident trailing matrix, one
//                         block barrier per column), emits R (fp32) and its 32x32 blocks
//   apply128_kernel       Q = A R^-1, one matrix row per thread, forward substitution (project
//                         out earlier 32-column blocks, then solve against the diagonal block);
//                         writes Q (fp32, in place) and its fp16 shadow
#include "panel.cuh"

#include <algorithm>
#include <cstdint>

namespace lb {
namespace {

constexpr int PW = kPanelWidth;        // 128
constexpr int GB = 8;                  // Gram register block (GB x GB doubles per thread)
constexpr int NBLK = PW / GB;          // 16 blocks per dimension
constexpr int NTRI = NBLK * (NBLK + 1) / 2;  // 136 upper-triangular blocks
constexpr int GRAM_THREADS = 160;      // 5 warps, 136 of them compute
constexpr int GRAM_ROWS = 32;          // rows staged per chunk
constexpr int GRAM_ELEMS = NTRI * GB * GB;   // 8704 doubles per partial

// Upper-triangular block index t -> (bi, bj), bi <= bj, row-major enumeration.
__host__ __device__ inline void tri_coords(int t, int& bi, int& bj) {
    int b = 0, rem = t;
    while (rem >= NBLK - b) { rem -= NBLK - b; ++b; }
    bi = b;
    bj = b + rem;
}
__host__ __device__ inline int tri_index(int bi, int bj) {  // bi <= bj
    return bi * NBLK - bi * (bi - 1) / 2 + (bj - bi);
}

// ---------------------------------------------------------------------------------------------
// Partial Gram matrix of the rows this CTA owns.  part layout: [cta][e = i*8+j][t] (t fastest).
__global__ void __launch_bounds__(GRAM_THREADS, 2)
gram128_f64_kernel(const float* __restrict__ A, long lda, int m, double* __restrict__ part) {
    __shared__ __align__(16) double As[GRAM_ROWS][PW];  // 32 KiB
    const int t = threadIdx.x;
    int bi = 0, bj = 0;
    if (t < NTRI) tri_coords(t, bi, bj);

    double acc[GB][GB];
#pragma unroll
    for (int i = 0; i < GB; ++i)
#pragma unroll
        for (int j = 0; j < GB; ++j) acc[i][j] = 0.0;

    const int nchunks = (m + GRAM_ROWS - 1) / GRAM_ROWS;
    const bool vec_ok = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    constexpr int VEC_PER_CHUNK = GRAM_ROWS * PW / 4;                       // 1024 float4
    constexpr int VEC_PER_THREAD = (VEC_PER_CHUNK + GRAM_THREADS - 1) / GRAM_THREADS;  // 7
    float4 pre[VEC_PER_THREAD];

    auto prefetch = [&](int chunk) {
        const int r0 = chunk * GRAM_ROWS;
#pragma unroll
        for (int k = 0; k < VEC_PER_THREAD; ++k) {
            const int idx = t + k * GRAM_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < VEC_PER_CHUNK) {
                const int col = idx >> 3;           // 8 float4 per column (32 rows)
                const int r = r0 + ((idx & 7) << 2);
                const float* src = A + r + (long)col * lda;
                if (vec_ok && r + 3 < m) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    if (r < m) v.x = src[0];
                    if (r + 1 < m) v.y = src[1];
                    if (r + 2 < m) v.z = src[2];
                    if (r + 3 < m) v.w = src[3];
                }
            }
            pre[k] = v;
        }
    };

    int chunk = blockIdx.x;
    if (chunk < nchunks) prefetch(chunk);
    for (; chunk < nchunks; chunk += gridDim.x) {
#pragma unroll
        for (int k = 0; k < VEC_PER_THREAD; ++k) {
            const int idx = t + k * GRAM_THREADS;
            if (idx < VEC_PER_CHUNK) {
                const int col = idx >> 3;
                const int r = (idx & 7) << 2;
                As[r][col] = (double)pre[k].x;
                As[r + 1][col] = (double)pre[k].y;
                As[r + 2][col] = (double)pre[k].z;
                As[r + 3][col] = (double)pre[k].w;
            }
        }
        __syncthreads();
        if (chunk + (int)gridDim.x < nchunks) prefetch(chunk + gridDim.x);  // in flight during math
        if (t < NTRI) {
#pragma unroll 2
            for (int r = 0; r < GRAM_ROWS; ++r) {
                double ai[GB], aj[GB];
                const double2* pi = reinterpret_cast<const double2*>(&As[r][bi * GB]);
                const double2* pj = reinterpret_cast<const double2*>(&As[r][bj * GB]);
#pragma unroll
                for (int q = 0; q < GB / 2; ++q) {
                    const double2 vi = pi[q], vj = pj[q];
                    ai[2 * q] = vi.x; ai[2 * q + 1] = vi.y;
                    aj[2 * q] = vj.x; aj[2 * q + 1] = vj.y;
                }
#pragma unroll
                for (int i = 0; i < GB; ++i)
#pragma unroll
                    for (int j = 0; j < GB; ++j) acc[i][j] = fma(ai[i], aj[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    if (t < NTRI) {
        double* dst = part + (long)blockIdx.x * GRAM_ELEMS + t;
#pragma unroll
        for (int i = 0; i < GB; ++i)
#pragma unroll
            for (int j = 0; j < GB; ++j) dst[(i * GB + j) * NTRI] = acc[i][j];
    }
}

__global__ void gram128_reduce_kernel(const double* __restrict__ part, int nparts,
                                      double* __restrict__ G) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= GRAM_ELEMS) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 0;
    for (; c + 3 < nparts; c += 4) {  // four independent chains, combined in a fixed order
        s0 += part[(long)c * GRAM_ELEMS + idx];
        s1 += part[(long)(c + 1) * GRAM_ELEMS + idx];
        s2 += part[(long)(c + 2) * GRAM_ELEMS + idx];
        s3 += part[(long)(c + 3) * GRAM_ELEMS + idx];
    }
    for (; c < nparts; ++c) s0 += part[(long)c * GRAM_ELEMS + idx];
    G[idx] = (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_nr(double x) {
    double y = (double)rsqrtf((float)x);
    y = y * (1.5 - 0.5 * x * y * y);
    y = y * (1.5 - 0.5 * x * y * y);
    return y;
}

// Element (i, j), i >= j, of the symmetric Gram matrix stored as packed upper blocks.
__device__ __forceinline__ double gram_elem(const double* __restrict__ G, int i, int j) {
    const int bi = j / GB, bj = i / GB;  // upper block (row block of j, column block of i)
    return G[((j % GB) * GB + (i % GB)) * NTRI + tri_index(bi, bj)];
}

// Scratch produced for apply128_kernel.
struct PanelFactors {
    float Roff[6][32][32];   // off-diagonal 32x32 blocks R(ib, jb), ib < jb, [k][c] row-major
    float Rdiag[4][32][32];  // diagonal blocks R(b, b), [k][c] row-major; only c >= k is defined
    float rinv[128];         // 1 / R(c, c)
};
__host__ __device__ inline int off_index(int ib, int jb) {  // ib < jb < 4
    return ib == 0 ? (jb - 1) : (ib == 1 ? (jb + 1) : 5);
}

// One CTA of 1024 threads = 32 warps.  Thread (warp = tx, lane = ty) owns the strided elements
// (i, j) = (ty + 32a, tx + 32b) of the lower triangle, in registers.  Column k belongs to warp
// k % 32, so the pivot is broadcast with a shuffle and each column costs one block barrier.
__global__ void __launch_bounds__(1024, 1)
chol128_kernel(const double* __restrict__ G, float* __restrict__ R, long ldr,
               PanelFactors* __restrict__ fac, int* __restrict__ info) {
    __shared__ double col[2][PW];
    const int tx = threadIdx.x >> 5;   // column residue / warp
    const int ty = threadIdx.x & 31;   // row residue / lane

    double a[4][4];
#pragma unroll
    for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
            const int i = ty + 32 * ia, j = tx + 32 * jb;
            a[ia][jb] = (i >= j) ? gram_elem(G, i, j) : 0.0;
        }
    // zero the strictly lower triangle of the caller's R block
    for (int e = threadIdx.x; e < PW * PW; e += 1024) {
        const int i = e % PW, j = e / PW;
        if (i > j) R[i + (long)j * ldr] = 0.f;
    }

#pragma unroll 1
    for (int k = 0; k < PW; ++k) {
        const int kb = k >> 5, kr = k & 31;
        if (tx == kr) {
            // pivot lives in lane kr, register a[kb][kb]
            double akk = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == kb) akk = a[q][q];
            akk = __shfl_sync(0xffffffffu, akk, kr);
            if (!(akk > 0.0)) {  // breakdown: panel numerically rank deficient
                if (ty == 0) atomicExch(info, k + 1);
                akk = 1e-300;
            }
            const double rs = rsqrt_nr(akk);
#pragma unroll
            for (int ia = 0; ia < 4; ++ia) {
                const int i = ty + 32 * ia;
                double aik = 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q == kb) aik = a[ia][q];
                if (i >= k) {
                    const double l = (i == k) ? akk * rs : aik * rs;
                    col[k & 1][i] = l;
                    R[k + (long)i * ldr] = (float)l;               // R = L^T
                    if (ia == kb) fac->Rdiag[kb][kr][ty] = (float)l;   // same 32-block
                    else fac->Roff[off_index(kb, ia)][kr][ty] = (float)l;
                } else {
                    col[k & 1][i] = 0.0;
                }
            }
            if (ty == kr) fac->rinv[k] = (float)rs;
        }
        __syncthreads();
        // rank-1 update of the trailing lower triangle
        double ci[4], cj[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ci[q] = col[k & 1][ty + 32 * q];
            cj[q] = col[k & 1][tx + 32 * q];
        }
#pragma unroll
        for (int ia = 0; ia < 4; ++ia)
#pragma unroll
            for (int jb = 0; jb < 4; ++jb)
                if (ia >= jb) a[ia][jb] = fma(-ci[ia], cj[jb], a[ia][jb]);
        // (entries of finished columns keep being updated; they are never read again)
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int APPLY_ROWS = 128;  // rows (= threads) per CTA
struct ApplySmem {
    float Q[PW][APPLY_ROWS];   // staged row block, column-major (conflict-free per-lane access)
    PanelFactors fac;
};

__global__ void __launch_bounds__(APPLY_ROWS)
apply128_kernel(float* __restrict__ A, long lda, int m, const PanelFactors* __restrict__ fac,
                __half* __restrict__ Qh, long ldqh) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ApplySmem& s = *reinterpret_cast<ApplySmem*>(smem_raw);
    const int t = threadIdx.x;
    const int row = blockIdx.x * APPLY_ROWS + t;
    const bool ok = row < m;

    {   // factors -> smem (float4 copies)
        const float4* src = reinterpret_cast<const float4*>(fac);
        float4* dst = reinterpret_cast<float4*>(&s.fac);
        for (int i = t; i < (int)(sizeof(PanelFactors) / 16); i += APPLY_ROWS) dst[i] = src[i];
    }
    for (int c = 0; c < PW; ++c) s.Q[c][t] = ok ? A[row + (long)c * lda] : 0.f;
    __syncthreads();

#pragma unroll 1
    for (int jb = 0; jb < 4; ++jb) {
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = s.Q[jb * 32 + c][t];
        // project out the finished blocks: a_jb -= q_ib * R(ib, jb)
#pragma unroll 1
        for (int ib = 0; ib < jb; ++ib) {
            const float (*Rb)[32] = s.fac.Roff[off_index(ib, jb)];
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const float qk = s.Q[ib * 32 + k][t];
                const float4* rr = reinterpret_cast<const float4*>(Rb[k]);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 r4 = rr[q];
                    acc[4 * q] = fmaf(-qk, r4.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(-qk, r4.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(-qk, r4.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(-qk, r4.w, acc[4 * q + 3]);
                }
            }
        }
        // normalise against the diagonal block by forward substitution:
        // q_k = a_k / R(k,k); a_c -= q_k R(k,c) for c > k
        const float (*Rd)[32] = s.fac.Rdiag[jb];
        float qv[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float qk = acc[k] * s.fac.rinv[jb * 32 + k];
            qv[k] = qk;
#pragma unroll
            for (int q = (k + 1) / 4; q < 8; ++q) {
                const float4 r4 = reinterpret_cast<const float4*>(Rd[k])[q];
                if (4 * q > k) acc[4 * q] = fmaf(-qk, r4.x, acc[4 * q]);
                if (4 * q + 1 > k) acc[4 * q + 1] = fmaf(-qk, r4.y, acc[4 * q + 1]);
                if (4 * q + 2 > k) acc[4 * q + 2] = fmaf(-qk, r4.z, acc[4 * q + 2]);
                if (4 * q + 3 > k) acc[4 * q + 3] = fmaf(-qk, r4.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            s.Q[jb * 32 + c][t] = qv[c];
            if (ok) {
                A[row + (long)(jb * 32 + c) * lda] = qv[c];
                if (Qh) Qh[row + (long)(jb * 32 + c) * ldqh] = __float2half_rn(qv[c]);
            }
        }
    }
}

int gram_grid(int m, int num_sms) {
    const int nchunks = (m + GRAM_ROWS - 1) / GRAM_ROWS;
    return std::max(1, std::min(nchunks, 2 * num_sms));
}

struct ScratchLayout {
    size_t part_off, g_off, fac_off, info_off, total;
};
ScratchLayout scratch_layout(int m, int num_sms) {
    ScratchLayout L{};
    size_t off = 0;
    L.part_off = off; off += (size_t)gram_grid(m, num_sms) * GRAM_ELEMS * sizeof(double);
    L.g_off = off;    off += (size_t)GRAM_ELEMS * sizeof(double);
    off = (off + 255) & ~(size_t)255;
    L.fac_off = off;  off += sizeof(PanelFactors);
    off = (off + 255) & ~(size_t)255;
    L.info_off = off; off += 256;
    L.total = off;
    return L;
}

}  // namespace

size_t panel_scratch_bytes(int m, int num_sms) { return scratch_layout(m, num_sms).total; }

cudaError_t panel_init() {
    return cudaFuncSetAttribute(apply128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(ApplySmem));
}

cudaError_t panel_qr128(cudaStream_t stream, int num_sms, int m, float* A, long lda, float* R,
                        long ldr, __half* Qh, long ldqh, void* scratch) {
    const ScratchLayout L = scratch_layout(m, num_sms);
    uint8_t* base = static_cast<uint8_t*>(scratch);
    double* part = reinterpret_cast<double*>(base + L.part_off);
    double* G = reinterpret_cast<double*>(base + L.g_off);
    PanelFactors* fac = reinterpret_cast<PanelFactors*>(base + L.fac_off);
    int* info = reinterpret_cast<int*>(base + L.info_off);
    const int ggrid = gram_grid(m, num_sms);

    gram128_f64_kernel<<<ggrid, GRAM_THREADS, 0, stream>>>(A, lda, m, part);
    gram128_reduce_kernel<<<(GRAM_ELEMS + 255) / 256, 256, 0, stream>>>(part, ggrid, G);
    chol128_kernel<<<1, 1024, 0, stream>>>(G, R, ldr, fac, info);
    apply128_kernel<<<(m + APPLY_ROWS - 1) / APPLY_ROWS, APPLY_ROWS, sizeof(ApplySmem), stream>>>(
        A, lda, m, fac, Qh, ldqh);
    return cudaGetLastError();
}

}  // namespace lb
