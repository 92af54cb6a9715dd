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
constexpr int GRAM_GROUP_THREADS = 160;  // 5 warps, 136 of them own a block
constexpr int GRAM_GROUPS = 2;
constexpr int GRAM_THREADS = GRAM_GROUP_THREADS * GRAM_GROUPS;
constexpr int GRAM_ROWS = 16;          // rows staged per chunk
constexpr int GRAM_BLK = 10;           // smem doubles per 8-column block (8 + 2 pad)
constexpr int GRAM_LDS = NBLK * GRAM_BLK;   // 160 doubles per staged row
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
// Partial Gram matrix of the rows this CTA owns.  One CTA per SM, two thread groups of 160 (136 of
// them own an 8x8 block of the upper triangle each); group g takes rows r = g (mod 2) of every
// staged 16-row chunk, the two groups' sums are combined in a fixed order at the end.
// part layout: [cta][e = i*8+j][t] (t fastest).
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram128_f64_kernel(const float* __restrict__ A, long lda, int m, double* __restrict__ part) {
    extern __shared__ __align__(16) uint8_t gram_smem[];
    // staged rows; every 8-double block is padded to 10 doubles (80 B) so that the 16-byte loads of
    // 8 lanes with consecutive block indices fall into 8 distinct bank groups
    double (*As)[GRAM_LDS] = reinterpret_cast<double (*)[GRAM_LDS]>(gram_smem);  // [GRAM_ROWS][160]
    double* comb = reinterpret_cast<double*>(gram_smem);                  // reused at the end
    const int tid = threadIdx.x;
    const int grp = tid / GRAM_GROUP_THREADS;
    const int t = tid - grp * GRAM_GROUP_THREADS;
    int bi = 0, bj = 0;
    if (t < NTRI) tri_coords(t, bi, bj);

    double acc[GB][GB];
#pragma unroll
    for (int i = 0; i < GB; ++i)
#pragma unroll
        for (int j = 0; j < GB; ++j) acc[i][j] = 0.0;

    const int nchunks = (m + GRAM_ROWS - 1) / GRAM_ROWS;
    const bool vec_ok = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    constexpr int VEC_PER_CHUNK = GRAM_ROWS * PW / 4;                                   // 512 float4
    constexpr int VEC_PER_THREAD = (VEC_PER_CHUNK + GRAM_THREADS - 1) / GRAM_THREADS;   // 2
    constexpr int VEC_PER_COL = GRAM_ROWS / 4;                                          // 4
    float4 pre[VEC_PER_THREAD];

    auto prefetch = [&](int chunk) {
        const int r0 = chunk * GRAM_ROWS;
#pragma unroll
        for (int k = 0; k < VEC_PER_THREAD; ++k) {
            const int idx = tid + k * GRAM_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < VEC_PER_CHUNK) {
                const int col = idx / VEC_PER_COL;
                const int r = r0 + ((idx % VEC_PER_COL) << 2);
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
            const int idx = tid + k * GRAM_THREADS;
            if (idx < VEC_PER_CHUNK) {
                const int col = idx / VEC_PER_COL;
                const int r = (idx % VEC_PER_COL) << 2;
                const int pc = (col >> 3) * GRAM_BLK + (col & 7);
                As[r][pc] = (double)pre[k].x;
                As[r + 1][pc] = (double)pre[k].y;
                As[r + 2][pc] = (double)pre[k].z;
                As[r + 3][pc] = (double)pre[k].w;
            }
        }
        __syncthreads();
        if (chunk + (int)gridDim.x < nchunks) prefetch(chunk + gridDim.x);  // in flight during math
        if (t < NTRI) {
#pragma unroll 2
            for (int r = grp; r < GRAM_ROWS; r += GRAM_GROUPS) {
                double ai[GB], aj[GB];
                const double2* pi = reinterpret_cast<const double2*>(&As[r][bi * GRAM_BLK]);
                const double2* pj = reinterpret_cast<const double2*>(&As[r][bj * GRAM_BLK]);
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
    // combine the groups (group 0 + group 1, fixed order) and emit the CTA's partial
    if (grp == 1 && t < NTRI) {
#pragma unroll
        for (int i = 0; i < GB; ++i)
#pragma unroll
            for (int j = 0; j < GB; ++j) comb[(i * GB + j) * NTRI + t] = acc[i][j];
    }
    __syncthreads();
    if (grp == 0 && t < NTRI) {
        double* dst = part + (long)blockIdx.x * GRAM_ELEMS + t;
#pragma unroll
        for (int i = 0; i < GB; ++i)
#pragma unroll
            for (int j = 0; j < GB; ++j)
                dst[(i * GB + j) * NTRI] = acc[i][j] + comb[(i * GB + j) * NTRI + t];
    }
}

// G[e] = sum over CTAs of part[c][e], summed in a fixed order: 32 slices of the CTA index per entry,
// each slice sequential, then the slices in order.  One CTA handles 32 consecutive entries.
__global__ void __launch_bounds__(1024)
gram128_reduce_kernel(const double* __restrict__ part, int nparts, double* __restrict__ G) {
    __shared__ double sh[32][33];
    const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + e;
    double s = 0.0;
    if (idx < GRAM_ELEMS)
        for (int c = sl; c < nparts; c += 32) s += part[(long)c * GRAM_ELEMS + idx];
    sh[sl][e] = s;
    __syncthreads();
    if (sl == 0 && idx < GRAM_ELEMS) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) tot += sh[k][e];
        G[idx] = tot;
    }
}

// ---------------------------------------------------------------------------------------------
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

// R = chol(G)^T in fp64, one CTA of 256 threads arranged 16 x 16: thread (tx = column residue,
// ty = row residue) keeps the elements (i, j) = (ty + 16 ia, tx + 16 jb), ia >= jb, of the trailing
// lower triangle in registers (36 doubles).  Column c is owned by the 16 threads with tx = c % 16 -
// half a warp - so the pivot is broadcast with a shuffle and each column costs ONE block barrier.
// Look-ahead: in the step that applies column c-1 every thread updates block column c/16 first, the
// owner of column c then derives it (pivot, reciprocal square root, scale) while the other warps
// are still busy with the rest of the rank-1 update.
__device__ __forceinline__ double rsqrt_f64(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);          // 1 - x y^2
    return fma(0.5 * y, e, y);                 // one Newton step: ~2^-45 relative
}

struct CholOut {
    float* R; long ldr;
    PanelFactors* fac;
    int* info;
};

// Derives column c (pivot, reciprocal square root, scaling) and publishes it in col[c & 1].
// Executed by the whole warp that contains the owner half-warp (tx == c % 16).
template <int CB>
__device__ __forceinline__ void chol_emit_column(double (&a)[8][8], int c, int tx, int ty,
                                                 double (*col)[PW], float* rinv_s, const CholOut& o) {
    const int cr = c & 15;
    double piv = a[CB][CB];
    piv = __shfl_sync(0xffffffffu, piv, ((cr & 1) << 4) + cr);   // lane with tx == cr, ty == cr
    if (tx != cr) return;
    if (!(piv > 0.0)) {                        // breakdown: numerically rank-deficient panel
        if (ty == 0) atomicExch(o.info, c + 1);
        piv = 1e-300;
    }
    const double rs = rsqrt_f64(piv);
#pragma unroll
    for (int ia = 0; ia < 8; ++ia) {
        const int i = ty + 16 * ia;
        double l = 0.0;
        if (ia >= CB && i >= c) l = (i == c) ? piv * rs : a[ia][CB] * rs;
        col[c & 1][i] = l;
    }
    if (ty == cr) rinv_s[c] = (float)rs;
}

// Writes column k of L (= row k of R) to the caller's R and to the factor blocks, from col[k & 1].
// Done by the warps that do NOT own the next column, so it stays off the critical path.
__device__ __forceinline__ void chol_output_column(int k, int q, const double (*col)[PW],
                                                   const CholOut& o) {
    if (q >= PW) return;
    const int i = q;
    if (i < k) { o.R[k + (long)i * o.ldr] = 0.f; return; }        // strictly lower part of R
    const float l = (float)col[k & 1][i];
    o.R[k + (long)i * o.ldr] = l;                                  // R(k, i) = L(i, k)
    const int rb = k >> 5, rr = k & 31, ib = i >> 5;
    if (ib == rb) o.fac->Rdiag[rb][rr][i & 31] = l;
    else o.fac->Roff[off_index(rb, ib)][rr][i & 31] = l;
}

template <int CB>
__device__ __forceinline__ void chol_block_column(double (&a)[8][8], int tx, int ty,
                                                  double (*col)[PW], float* rinv_s, const CholOut& o) {
    const int warp = tx >> 1;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int cr = (CB == 0 ? 1 : 0); cr < 16; ++cr) {
        const int c = CB * 16 + cr, k = c - 1;
        __syncthreads();                       // column k is in col[k & 1]
        double ci[8], cj[8];
#pragma unroll
        for (int q = CB; q < 8; ++q) {
            ci[q] = col[k & 1][ty + 16 * q];
            cj[q] = col[k & 1][tx + 16 * q];
        }
        // block column CB first (it contains column c) ...
#pragma unroll
        for (int ia = CB; ia < 8; ++ia) a[ia][CB] = fma(-ci[ia], cj[CB], a[ia][CB]);
        // ... so that its owner can already derive column c
        const int owner = cr >> 1;
        if (warp == owner) chol_emit_column<CB>(a, c, tx, ty, col, rinv_s, o);
        else chol_output_column(k, (((warp - owner - 1) & 7) << 5) + lane, col, o);
        // rest of the rank-1 update
#pragma unroll
        for (int jb = CB + 1; jb < 8; ++jb)
#pragma unroll
            for (int ia = jb; ia < 8; ++ia) a[ia][jb] = fma(-ci[ia], cj[jb], a[ia][jb]);
    }
}

__global__ void __launch_bounds__(256, 1)
chol128_kernel(const double* __restrict__ G, float* __restrict__ R, long ldr,
               PanelFactors* __restrict__ fac, int* __restrict__ info) {
    __shared__ double col[2][PW];
    __shared__ float rinv_s[PW];
    const int tx = threadIdx.x >> 4;   // column residue
    const int ty = threadIdx.x & 15;   // row residue
    const CholOut o{R, ldr, fac, info};

    double a[8][8];
#pragma unroll
    for (int ia = 0; ia < 8; ++ia)
#pragma unroll
        for (int jb = 0; jb < 8; ++jb) {
            const int i = ty + 16 * ia, j = tx + 16 * jb;
            a[ia][jb] = (ia >= jb && i >= j) ? gram_elem(G, i, j) : 0.0;
        }
    if ((tx >> 1) == 0) chol_emit_column<0>(a, 0, tx, ty, col, rinv_s, o);   // column 0: no update
    chol_block_column<0>(a, tx, ty, col, rinv_s, o);
    chol_block_column<1>(a, tx, ty, col, rinv_s, o);
    chol_block_column<2>(a, tx, ty, col, rinv_s, o);
    chol_block_column<3>(a, tx, ty, col, rinv_s, o);
    chol_block_column<4>(a, tx, ty, col, rinv_s, o);
    chol_block_column<5>(a, tx, ty, col, rinv_s, o);
    chol_block_column<6>(a, tx, ty, col, rinv_s, o);
    chol_block_column<7>(a, tx, ty, col, rinv_s, o);
    __syncthreads();
    chol_output_column(PW - 1, threadIdx.x, col, o);               // last column
    if (threadIdx.x < PW) fac->rinv[threadIdx.x] = rinv_s[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
constexpr int APPLY_ROWS = 64;   // rows (= threads) per CTA
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
#pragma unroll 32
    for (int c = 0; c < PW; ++c) s.Q[c][t] = ok ? A[row + (long)c * lda] : 0.f;   // 32 loads in flight
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
#pragma unroll 8
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
    return std::max(1, std::min(nchunks, num_sms));
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

constexpr int GRAM_SMEM = GRAM_ELEMS * (int)sizeof(double);   // combine buffer (> 16x160 staging tile)

cudaError_t panel_init() {
    cudaError_t e = cudaFuncSetAttribute(gram128_f64_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
    if (e != cudaSuccess) return e;
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

    gram128_f64_kernel<<<ggrid, GRAM_THREADS, GRAM_SMEM, stream>>>(A, lda, m, part);
    gram128_reduce_kernel<<<(GRAM_ELEMS + 31) / 32, 1024, 0, stream>>>(part, ggrid, G);
    chol128_kernel<<<1, 256, 0, stream>>>(G, R, ldr, fac, info);
    apply128_kernel<<<(m + APPLY_ROWS - 1) / APPLY_ROWS, APPLY_ROWS, sizeof(ApplySmem), stream>>>(
        A, lda, m, fac, Qh, ldqh);
    return cudaGetLastError();
}

}  // namespace lb
