// 128-column tall-skinny panel QR (see panel.cuh).  Four kernels per panel:
//   gram128_f64_kernel    one pass over the panel: all column norms and dot products, exact
//                         fp32 x fp32 products accumulated in fp64 on the fp64 tensor path
//                         (mma.sync.m8n8k4.f64), per-CTA partial Gram blocks
//   gram128_reduce_kernel fixed-order sum of the per-CTA partials (deterministic)
//   chol128b_kernel       R = chol(G) in fp64 on one CTA, blocked by 32 columns: one warp factors each
//                         diagonal block (two columns per step), the others solve, update (fp64
//                         tensor path) and publish R block-row by block-row; barriers only
//   apply128_kernel       Q = A R^-1 by forward substitution, right-looking, two threads per matrix row; launched
//                         with programmatic dependent launch so that it runs CONCURRENTLY with the
//                         Cholesky kernel and consumes each 32-row block of R as soon as its flag is
//                         raised; writes Q (fp32, in place) and its fp16 shadow
#include "panel.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace lb {
namespace {

using namespace ptx;

constexpr int PW = kPanelWidth;        // 128
constexpr int GW = 32;                 // Gram block owned by one warp: GW x GW (4 x 4 DMMA tiles)
constexpr int NGB = PW / GW;           // 4 blocks per dimension
constexpr int NTRI = NGB * (NGB + 1) / 2;    // 10 upper-triangular blocks = 10 warps
constexpr int GRAM_THREADS = NTRI * 32;      // 320
constexpr int GRAM_ROWS = 16;          // rows staged per chunk (4 DMMA k-steps)
constexpr int GRAM_LDS = PW + 4;       // 132 doubles per staged row: the 4 rows of a k-step land in
                                       // 4 different bank groups (row stride = 32 B mod 128 B)
constexpr int GRAM_ELEMS = NTRI * GW * GW;   // 10240 doubles per partial: [block][i][j]

// Upper-triangular block index t -> (bi, bj), bi <= bj, row-major enumeration.
__host__ __device__ inline void tri_coords(int t, int& bi, int& bj) {
    int b = 0, rem = t;
    while (rem >= NGB - b) { rem -= NGB - b; ++b; }
    bi = b;
    bj = b + rem;
}
__host__ __device__ inline int tri_index(int bi, int bj) {  // bi <= bj
    return bi * NGB - bi * (bi - 1) / 2 + (bj - bi);
}

// Factors handed from the Cholesky kernel to the apply kernel.  Columns inside a 32-wide block are
// stored permuted, perm(c) = (c % 4) * 8 + c / 4, so that the apply thread that owns the columns
// c = p (mod 4) reads its 8 entries of a row with two 16-byte loads.
struct PanelFactors {
    float Roff[6][32][32];   // off-diagonal blocks R(ib, jb), ib < jb: [k][perm(c)]
    float Rdiag[4][32][32];  // diagonal blocks R(b, b): [k][perm(c)], only c >= k is defined
    float rinv[128];         // 1 / R(c, c)
    int flag[4];             // flag[b] = 1 once block-row b of R (and rinv) is complete
    int pad[28];
};
__host__ __device__ inline int off_index(int ib, int jb) {  // ib < jb < 4
    return ib == 0 ? (jb - 1) : (ib == 1 ? (jb + 1) : 5);
}
__host__ __device__ inline int perm32(int c) { return ((c & 3) << 3) | (c >> 2); }

// ---------------------------------------------------------------------------------------------
// Partial Gram matrix of the rows this CTA owns, on the fp64 tensor path: warp w owns the 32 x 32
// block (bi, bj), bi <= bj, of G as 4 x 4 accumulator tiles of mma.sync.m8n8k4.f64 (DMMA runs at the
// DFMA rate on B200, 64 FMA/clk/SM measured, but needs 4x fewer shared-memory wavefronts and 16x
// fewer issue slots than an 8x8 register-blocked DFMA loop).  Products of fp32 inputs are exact in
// fp64.  part layout: [cta][block][i][j].
__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram128_f64_kernel(const float* __restrict__ A, long lda, int m, double* __restrict__ part,
                   const int* __restrict__ cond) {
    __shared__ __align__(16) double As[GRAM_ROWS][GRAM_LDS];   // 16.5 KiB
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    int bi, bj;
    tri_coords(warp, bi, bj);
    const int fr = lane & 3;       // row of the k-step this lane feeds (fragment K index)
    const int fc = lane >> 2;      // column inside an 8-wide tile (fragment M / N index)

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int nchunks = (m + GRAM_ROWS - 1) / GRAM_ROWS;
    const bool vec_ok = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    constexpr int VEC_PER_CHUNK = GRAM_ROWS * PW / 4;                                   // 512 float4
    constexpr int VEC_PER_THREAD = (VEC_PER_CHUNK + GRAM_THREADS - 1) / GRAM_THREADS;   // 2
    float4 pre[VEC_PER_THREAD];

    // element idx of a chunk: column = idx % 128 (consecutive lanes -> consecutive smem words),
    // row quad = idx / 128
    auto prefetch = [&](int chunk) {
        const int r0 = chunk * GRAM_ROWS;
#pragma unroll
        for (int k = 0; k < VEC_PER_THREAD; ++k) {
            const int idx = tid + k * GRAM_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < VEC_PER_CHUNK) {
                const int col = idx & (PW - 1);
                const int r = r0 + ((idx >> 7) << 2);
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

    pdl_trigger();
    pdl_wait();      // the panel is written by the preceding update kernel
    if (cond && *cond == 0) return;   // fallback launch of a panel that did not need it (whole grid)
    int chunk = blockIdx.x;
    if (chunk < nchunks) prefetch(chunk);
    for (; chunk < nchunks; chunk += gridDim.x) {
#pragma unroll
        for (int k = 0; k < VEC_PER_THREAD; ++k) {
            const int idx = tid + k * GRAM_THREADS;
            if (idx < VEC_PER_CHUNK) {
                const int col = idx & (PW - 1);
                const int r = (idx >> 7) << 2;
                As[r][col] = (double)pre[k].x;
                As[r + 1][col] = (double)pre[k].y;
                As[r + 2][col] = (double)pre[k].z;
                As[r + 3][col] = (double)pre[k].w;
            }
        }
        __syncthreads();
        if (chunk + (int)gridDim.x < nchunks) prefetch(chunk + gridDim.x);  // in flight during math
#pragma unroll
        for (int ks = 0; ks < GRAM_ROWS / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) a[t] = As[ks * 4 + fr][bi * GW + t * 8 + fc];
            if (bi == bj) {
#pragma unroll
                for (int t = 0; t < 4; ++t) b[t] = a[t];
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) b[t] = As[ks * 4 + fr][bj * GW + t * 8 + fc];
            }
            if (bi == bj) {
                // diagonal block: only its upper tiles are ever read (G is symmetric)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = i; j < 4; ++j) dmma_884(acc[i][j], a[i], b[j]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma_884(acc[i][j], a[i], b[j]);
            }
        }
        __syncthreads();
    }
    // accumulator tile (i, j): lane holds G[8 i + lane / 4][8 j + 2 (lane % 4) + {0, 1}] of the block
    double* dst = part + (long)blockIdx.x * GRAM_ELEMS + (long)warp * GW * GW;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(&dst[(8 * i + fc) * GW + 8 * j + 2 * fr]) =
                make_double2(acc[i][j][0], acc[i][j][1]);
}

// G[e] = sum over CTAs of part[c][e], summed in a fixed order: 32 slices of the CTA index per entry,
// each slice sequential, then the slices in order.  One CTA handles 32 consecutive entries.
// Also clears the block-row flags of the factor hand-over for this panel.
__global__ void __launch_bounds__(1024)
gram128_reduce_kernel(const double* __restrict__ part, int nparts, double* __restrict__ G,
                      int* __restrict__ flags, const int* __restrict__ cond) {
    __shared__ double sh[32][33];
    const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + e;
    pdl_trigger();
    pdl_wait();
    if (cond && *cond == 0) return;
    if (blockIdx.x == 0 && threadIdx.x < 4) flags[threadIdx.x] = 0;
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
// Element (i, j), i >= j, of the symmetric Gram matrix stored as upper 32 x 32 blocks [block][r][c].
__device__ __forceinline__ double gram_elem(const double* __restrict__ G, int i, int j) {
    const int bi = j / GW, bj = i / GW;  // upper block: row block of j, column block of i
    return G[tri_index(bi, bj) * GW * GW + (j % GW) * GW + (i % GW)];
}

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
    int* info;          // status words of the factorisation (PanelInfo)
    int col0;           // global index of the panel's first column
    int check_redo;     // 1: this is the first attempt of an integer-Gram panel - a small or bad pivot
                        //    raises info[INFO_REDO] instead of being reported
    double tau;         // pivot-ratio threshold of that check
};

// ---------------------------------------------------------------------------------------------
// R = chol(G)^T in fp64, one CTA, blocked (32-column blocks), synchronised with barriers only.
//
// Per block step b:
//   warp 0          factors the 32 x 32 diagonal block entirely by itself: lane i owns row i in
//                   registers, two columns per step (the two reciprocal square roots independent of
//                   each other), the column pair broadcast to the other lanes through 512 bytes of
//                   shared memory behind a __syncwarp: ~425 cycles per pair (a chain of ~10 dependent
//                   fp64 operations at ~20 cycles each plus the shuffle / shared-memory round trip), where
//                   round 1's register-resident kernel with its cross-warp publish / poll ring needed ~840.
//                   (That kernel is gone: the two were equally fast end to end - the trailing updates on
//                   a single SM's fp64 pipe bound both - but compute-sanitizer racecheck cannot model its
//                   ld.acquire / st.release protocol and reported hazards; this one is clean.)
//   warps 1 .. 11   meanwhile finish step b - 1: the rest of its trailing update and its output (32 rows
//                   of R, the factor blocks of the apply kernel, 1/diag, the block-row flag)
//   all             triangular solve of the rows below (one row per thread), then the update of the NEXT
//                   diagonal block only, so that warp 0 can go on at once
// Everything lives in shared memory (trailing matrix 129 KB, two block columns of L 65 KB).
constexpr int CH2_THREADS = 384;
constexpr int CH2_LDA = PW + 1;        // doubles per row of the trailing matrix (odd: rows hit distinct banks)
constexpr int CH2_LDL = PW + 4;        // doubles per k-row of a block column of L: 16-byte aligned pairs, and the
                                       // 4 k x 8 i elements of a DMMA fragment land in distinct banks
struct Chol2Smem {
    double As[PW][CH2_LDA];            // trailing matrix, As[i][j], i >= j
    double Lt[2][GW][CH2_LDL];         // block column b of L in Lt[b & 1]: Lt[.][k][i] = L(i, 32 b + k)
    double2 pair[2][GW];               // diagonal-block factorisation: the current column pair, per row
    double rs[PW];                     // 1 / L(c, c)
    double gdiag[PW];                  // G(c, c)
    float Rs[GW][PW + 1];              // output staging: block-row b of R in fp32, Rs[k][i] = L(i, 32 b + k)
    int bad;                           // 1 + local column of the first non-positive pivot (0 = none)
#ifdef LB_CHOL_TRACE
    long long tr[48];                  // phase timestamps (thread 0 / thread 32)
#endif
};
#ifdef LB_CHOL_TRACE
#define CH2_MARK(slot) do { if (tid == 0) s.tr[slot] = clock64(); } while (0)
#define CH2_MARK1(slot) do { if (tid == 32) s.tr[slot] = clock64(); } while (0)
#else
#define CH2_MARK(slot) do {} while (0)
#define CH2_MARK1(slot) do {} while (0)
#endif

// Warp 0: Cholesky of the diagonal block b.  Lane i owns row 32 b + i.
__device__ __forceinline__ void chol2_diag_block(Chol2Smem& s, int b, int lane) {
    const int o = GW * b;
    double (*Lt)[CH2_LDL] = s.Lt[b & 1];
    double a[GW];
#pragma unroll
    for (int j = 0; j < GW; ++j) a[j] = j <= lane ? s.As[o + lane][o + j] : 0.0;
#pragma unroll
    for (int kk = 0; kk < GW; kk += 2) {
        const double piv0 = __shfl_sync(0xffffffffu, a[kk], kk);             // (c, c)
        const double a10 = __shfl_sync(0xffffffffu, a[kk], kk + 1);          // (c+1, c)
        const double a11 = __shfl_sync(0xffffffffu, a[kk + 1], kk + 1);      // (c+1, c+1)
        // piv1 = a11 - a10^2 / piv0 = d / piv0: rsqrt(piv1) = rsqrt(d) sqrt(piv0), so the two reciprocal
        // square roots start together
        const double d = fma(a11, piv0, -a10 * a10);
        const bool bad0 = !(piv0 > 0.0), bad1 = !(d > 0.0);
        const double p0 = bad0 ? 1e-300 : piv0;
        const double rs0 = rsqrt_f64(p0);
        const double rd = rsqrt_f64(bad1 ? 1e-300 : d);
        const double rs1 = rd * (p0 * rs0);
        const double w = a10 * rs0 * rs0;                                    // L(c+1, c) / L(c, c)
        double l0 = a[kk] * rs0;
        double l1 = fma(-a[kk], w, a[kk + 1]) * rs1;
        if (lane < kk) l0 = 0.0;
        if (lane < kk + 1) l1 = 0.0;
        s.pair[(kk >> 1) & 1][lane] = make_double2(l0, l1);
        Lt[kk][o + lane] = l0;
        Lt[kk + 1][o + lane] = l1;
        if (lane == 0) {
            s.rs[o + kk] = rs0;
            s.rs[o + kk + 1] = rs1;
            if (bad0 || bad1) atomicCAS(&s.bad, 0, o + kk + (bad0 ? 1 : 2));
        }
        __syncwarp();
#pragma unroll
        for (int j = kk + 2; j < GW; ++j) {
            const double2 lj = s.pair[(kk >> 1) & 1][j];                     // broadcast
            a[j] = fma(-l1, lj.y, fma(-l0, lj.x, a[j]));
        }
    }
}

// Row i (below block b) of the block column: x L_bb^T = A(i, block b), right-looking in registers.
__device__ __forceinline__ void chol2_trsm_row(Chol2Smem& s, int b, int i) {
    const int o = GW * b;
    double (*Lt)[CH2_LDL] = s.Lt[b & 1];
    double x[GW];
#pragma unroll
    for (int j = 0; j < GW; ++j) x[j] = s.As[i][o + j];
#pragma unroll
    for (int k = 0; k < GW; ++k) {
        const double xk = x[k] * s.rs[o + k];
        x[k] = xk;
        // L(o + j, o + k), j > k: broadcast loads, two entries each
        if ((k & 1) == 0) x[k + 1] = fma(-xk, Lt[k][o + k + 1], x[k + 1]);
#pragma unroll
        for (int j = (k + 2) & ~1; j < GW; j += 2) {
            const double2 l = *reinterpret_cast<const double2*>(&Lt[k][o + j]);
            x[j] = fma(-xk, l.x, x[j]);
            x[j + 1] = fma(-xk, l.y, x[j + 1]);
        }
    }
#pragma unroll
    for (int k = 0; k < GW; ++k) Lt[k][i] = x[k];
}

// One 8 x 8 tile of the trailing update on the fp64 tensor path: A(i0.., j0..) -= L(i0.., :) L(j0.., :)^T over the
// 32 columns of block column b (8 x mma.m8n8k4; fragments straight from Lt: A[r = lane / 4][k = lane % 4],
// B[k = lane % 4][n = lane / 4], C[r = lane / 4][2 (lane % 4) + {0, 1}]).
__device__ __forceinline__ void chol2_tile_update(Chol2Smem& s, const double (*Lt)[CH2_LDL], int i0, int j0, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    double* c = &s.As[i0 + fr][j0 + 2 * fk];
    double acc[2] = {-c[0], -c[1]};                    // accumulate -A + L L^T, negate at the end
#pragma unroll
    for (int k = 0; k < GW; k += 4) dmma_884(acc, Lt[k + fk][i0 + fr], Lt[k + fk][j0 + fr]);
    c[0] = -acc[0];
    c[1] = -acc[1];
}

__device__ __forceinline__ void chol2_tile_update2(Chol2Smem& s, const double (*Lt)[CH2_LDL], int i0, int j0, int i1,
                                                   int j1, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    double* c0 = &s.As[i0 + fr][j0 + 2 * fk];
    double* c1 = &s.As[i1 + fr][j1 + 2 * fk];
    double acc0[2] = {-c0[0], -c0[1]}, acc1[2] = {-c1[0], -c1[1]};
#pragma unroll
    for (int k = 0; k < GW; k += 4) {
        dmma_884(acc0, Lt[k + fk][i0 + fr], Lt[k + fk][j0 + fr]);
        dmma_884(acc1, Lt[k + fk][i1 + fr], Lt[k + fk][j1 + fr]);
    }
    c0[0] = -acc0[0]; c0[1] = -acc0[1];
    c1[0] = -acc1[0]; c1[1] = -acc1[1];
}

// The next diagonal block only: A(i, j) -= sum_k L(i, k) L(j, k), i, j in block b + 1, j <= i: the ten lower
// 8 x 8 tiles, one per warp.
__device__ __forceinline__ void chol2_update_next_diag(Chol2Smem& s, int b, int warp, int lane) {
    if (warp >= 10) return;
    const int o = GW * (b + 1);
    int ti = 0, t = warp;
    while (t > ti) { t -= ti + 1; ++ti; }              // warp -> (ti, tj), tj <= ti
    chol2_tile_update(s, s.Lt[b & 1], o + 8 * ti, o + 8 * t, lane);
}

// The rest of step b's trailing update (blocks (bi, bj), bi > b + 1, b + 1 <= bj <= bi), by warps 1 .. 11,
// 8 x 8 tiles dealt round-robin (tiles above the diagonal of a diagonal block are skipped).
__device__ __forceinline__ void chol2_update_rest(Chol2Smem& s, int b, int w, int lane) {
    const double (*Lt)[CH2_LDL] = s.Lt[b & 1];
    int unit = 0, pend_i = -1, pend_j = 0;             // two tiles at a time: independent DMMA chains
    for (int bi = b + 2; bi < NGB; ++bi)
        for (int bj = b + 1; bj <= bi; ++bj)
            for (int ti = 0; ti < 4; ++ti)
                for (int tj = 0; tj < 4; ++tj) {
                    if (bi == bj && tj > ti) continue;
                    if (unit++ % 11 != w) continue;
                    const int i0 = GW * bi + 8 * ti, j0 = GW * bj + 8 * tj;
                    if (pend_i < 0) { pend_i = i0; pend_j = j0; continue; }
                    chol2_tile_update2(s, Lt, pend_i, pend_j, i0, j0, lane);
                    pend_i = -1;
                }
    if (pend_i >= 0) chol2_tile_update(s, Lt, pend_i, pend_j, lane);
}

// Block-row b of R and everything the apply kernel needs from it, by the threads t = 0 .. nt - 1 of the
// warps that do not factor the next diagonal block; ends with the block-row flag.  The block row is
// rounded to fp32 into a staging array first (lanes along i: conflict-free), so that the global writes
// can run along k (128 contiguous bytes per warp) without bank conflicts.
__device__ __forceinline__ void chol2_output(Chol2Smem& s, int b, int t, int nt, const CholOut& o, int barrier_id) {
    const int ob = GW * b;
    const double (*Lt)[CH2_LDL] = s.Lt[b & 1];
    for (int e = t; e < GW * PW; e += nt) {
        const int i = e & (PW - 1), k = e >> 7;
        s.Rs[k][i] = i < ob + k ? 0.f : (float)Lt[k][i];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(barrier_id), "r"(nt) : "memory");
    // R(k, i) = L(i, k), zero for i < k: lanes along k
    for (int e = t; e < GW * PW; e += nt) {
        const int k = e & 31, i = e >> 5;
        o.R[ob + k + (long)i * o.ldr] = s.Rs[k][i];
    }
    // factor blocks [k][perm(c)]: lanes along i
    for (int e = t; e < GW * (PW - ob); e += nt) {
        const int c = e & 31, k = (e >> 5) & 31, j = b + (e >> 10);          // column block j >= b of R
        const float v = s.Rs[k][GW * j + c];
        if (j == b) o.fac->Rdiag[b][k][perm32(c)] = v;
        else o.fac->Roff[off_index(b, j)][k][perm32(c)] = v;
    }
    if (t < GW) o.fac->rinv[ob + t] = (float)s.rs[ob + t];
    // Publish: the barrier orders every thread's writes before thread 0's release store (release is
    // cumulative over what the barrier made visible to it), so no thread needs a fence of its own -
    // a __threadfence() per thread here stalled the whole CTA for thousands of cycles.
    asm volatile("bar.sync %0, %1;" ::"r"(barrier_id), "r"(nt) : "memory");
    if (t == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(&o.fac->flag[b]), "r"(1) : "memory");
}

__global__ void __launch_bounds__(CH2_THREADS, 1)
chol128b_kernel(const double* __restrict__ G, float* __restrict__ R, long ldr, PanelFactors* __restrict__ fac,
                int* __restrict__ info, int col0, int check_redo, double tau, const int* __restrict__ cond) {
    extern __shared__ __align__(16) uint8_t chol2_raw[];
    Chol2Smem& s = *reinterpret_cast<Chol2Smem*>(chol2_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const CholOut o{R, ldr, fac, info, col0, check_redo, tau};
    if (tid == 0) s.bad = 0;
    CH2_MARK(0);
    pdl_wait();      // G and the cleared flags come from the reduce kernel
    CH2_MARK(1);
    if (cond) {      // fallback launch: runs only if the first attempt asked for it
        if (*cond == 0) return;
        if (tid == 0) { atomicAdd(&info[INFO_FALLBACKS], 1); atomicOr(&info[INFO_FLAGS], 2); }
    }
    // Only now may the dependent apply grid start: it synchronises on fac->flag[] (not on the
    // completion of this grid), so the flags must already have been cleared by the reduce kernel.
    pdl_trigger();
    // G (upper 32 x 32 blocks [block][r][c]) -> lower triangle of As
    {
        constexpr int NLD = (GRAM_ELEMS + CH2_THREADS - 1) / CH2_THREADS;     // 27 loads per thread, all in flight
        double v[NLD];
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int idx = tid + q * CH2_THREADS;
            v[q] = idx < GRAM_ELEMS ? G[idx] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int idx = tid + q * CH2_THREADS;
            if (idx < GRAM_ELEMS) {
                // block t = idx >> 10 of the upper triangle -> (bi, bj), from two packed tables
                const int t4 = (idx >> 10) * 4;
                const int bi = (int)((0x3221110000ull >> t4) & 15), bj = (int)((0x3323213210ull >> t4) & 15);
                const int j = GW * bi + ((idx >> 5) & 31), i = GW * bj + (idx & 31);
                if (i >= j) {
                    s.As[i][j] = v[q];
                    if (i == j) s.gdiag[i] = v[q];
                }
            }
        }
    }
    __syncthreads();
    CH2_MARK(2);
#pragma unroll 1
    for (int b = 0; b < NGB; ++b) {
        if (warp == 0) {
            chol2_diag_block(s, b, lane);
            CH2_MARK(3 + 8 * b);
        } else if (b > 0) {
            // (output first: the apply kernel, which is already running, gets its block row ~4000
            // cycles earlier; the update only has to be done by the next barrier)
            chol2_output(s, b - 1, tid - 32, CH2_THREADS - 32, o, 1);
            CH2_MARK1(4 + 8 * b);
            chol2_update_rest(s, b - 1, warp - 1, lane);
            CH2_MARK1(5 + 8 * b);
        }
        __syncthreads();
        CH2_MARK(6 + 8 * b);
        const int below = GW * (b + 1);
        if (below + tid < PW) chol2_trsm_row(s, b, below + tid);
#ifdef LB_CHOL_TRACE
        asm volatile("" ::: "memory");
        if (tid == 0) s.tr[36 + b] = clock64();
        if (tid == 32) s.tr[40 + b] = clock64();
        if (tid == 64) s.tr[44 + b] = clock64();
        asm volatile("" ::: "memory");
#endif
        __syncthreads();
        CH2_MARK(7 + 8 * b);
        if (b + 1 < NGB) chol2_update_next_diag(s, b, warp, lane);
        __syncthreads();
        CH2_MARK(8 + 8 * b);
    }
    chol2_output(s, NGB - 1, tid, CH2_THREADS, o, 2);
    CH2_MARK(35);
#ifdef LB_CHOL_TRACE
    __syncthreads();
    if (tid == 0) {
        const long long t0 = s.tr[0];
        printf("chol2: pdl_wait %lld load %lld |", s.tr[1] - t0, s.tr[2] - s.tr[1]);
        for (int b = 0; b < NGB; ++b) {
            const long long start = b == 0 ? s.tr[2] : s.tr[8 * b];
            printf(" b%d: diag %lld", b, s.tr[3 + 8 * b] - start);
            if (b > 0) printf(" (out %lld rest %lld)", s.tr[4 + 8 * b] - start, s.tr[5 + 8 * b] - s.tr[4 + 8 * b]);
            printf(" barA %lld trsm %lld upd %lld |", s.tr[6 + 8 * b] - start, s.tr[7 + 8 * b] - s.tr[6 + 8 * b],
                   s.tr[8 + 8 * b] - s.tr[7 + 8 * b]);
        }
        printf(" final out %lld total %lld\n", s.tr[35] - s.tr[32], s.tr[35] - t0);
        for (int b = 0; b < 3; ++b)
            printf("  trsm b%d: warp0 %lld warp1 %lld warp2 %lld (from barA)\n", b, s.tr[36 + b] - s.tr[6 + 8 * b],
                   s.tr[40 + b] - s.tr[6 + 8 * b], s.tr[44 + b] - s.tr[6 + 8 * b]);
    }
#endif
    // Status of the panel.
    if (warp == 0) {
        double min_ratio = 1.0;
#pragma unroll
        for (int q = 0; q < PW / 32; ++q) {
            const double rs = s.rs[lane + 32 * q], gd = s.gdiag[lane + 32 * q];
            // piv / G_kk = 1 / (rs^2 G_kk); a zero or non-finite G_kk counts as breakdown
            min_ratio = fmin(min_ratio, (gd > 0.0 && gd < 1e300) ? 1.0 / (rs * rs * gd) : 0.0);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) min_ratio = fmin(min_ratio, __shfl_xor_sync(0xffffffffu, min_ratio, off));
        if (lane == 0) {
            const int bad = s.bad;
            if (check_redo) {
                info[INFO_REDO] = (bad != 0 || !(min_ratio >= tau)) ? 1 : 0;
            } else {
                if (bad != 0) atomicCAS(&info[INFO_BAD_COLUMN], 0, col0 + bad);
                int e = 0;
                if (min_ratio > 0.0) { frexp(min_ratio, &e); e = 1 - e; } else { e = 2047; }
                atomicMax(&info[INFO_COND_LOG2], e);      // ~ -log2(min ratio), rounded up
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Q = A R^-1, right-looking.  128 rows per CTA, 256 threads: thread (rg, p) owns rows 2 rg, 2 rg + 1 and,
// in each of the four 32-column blocks, the columns c = p (mod 4) - the whole 2 x 32 share of the panel
// lives in registers from the first stage to the last.  Stage jb waits for block-row jb of R (flag raised
// by the Cholesky kernel, which is still running), turns block jb into Q by forward substitution against
// the diagonal block (the owner of column k broadcasting q_k to the three other parts of its quad), writes
// it out, and projects it out of the blocks to its right at once (a_j -= q_jb R(jb, j), j > jb).  When the
// last flag goes up only the 32 x 32 substitution of block 3 is left - the Cholesky kernel is the critical
// path of the panel and this is what it is followed by.  Every accumulator sees the same fused
// multiply-adds in the same order as in a left-looking sweep (blocks ascending, k ascending).
constexpr int APPLY_ROWS = 128;
constexpr int APPLY_THREADS = 2 * APPLY_ROWS;    // (APPLY_ROWS / 2) row pairs x 4 column parts
constexpr int APPLY_LDQ = APPLY_ROWS + 16;       // 144 floats: parts p = 0..3 land 16 banks apart
struct ApplySmem {
    float Q[32][APPLY_LDQ];    // the block just finished (each quad reads back only its own rows)
    float Rb[4][32][32];       // block-row jb of R: [0] = diagonal block, [d] = R(jb, jb + d); [k][perm(c)]
    float rinv[32];
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// acc[d] -= q R(jb, jb + d) for the ND blocks to the right of the one just finished
template <int ND>
__device__ __forceinline__ void apply_project(float (&acc)[4][2][8], const ApplySmem& s, int rg, int p) {
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const float2 qv = *reinterpret_cast<const float2*>(&s.Q[k][2 * rg]);
#pragma unroll
        for (int d = 1; d <= ND; ++d) {
            const float4 ra = *reinterpret_cast<const float4*>(&s.Rb[d][k][p * 8]);
            const float4 rb = *reinterpret_cast<const float4*>(&s.Rb[d][k][p * 8 + 4]);
            const float rr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc[d][0][q] = fmaf(-qv.x, rr[q], acc[d][0][q]);
                acc[d][1][q] = fmaf(-qv.y, rr[q], acc[d][1][q]);
            }
        }
    }
}

__global__ void __launch_bounds__(APPLY_THREADS, 2)
apply128_kernel(float* __restrict__ A, long lda, int m, const PanelFactors* __restrict__ fac,
                __half* __restrict__ Qh, long ldqh, const int* __restrict__ only_if) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ApplySmem& s = *reinterpret_cast<ApplySmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int rg = tid >> 2, p = tid & 3;    // row pair, column part
    const int lane = tid & 31;
    const int grow = blockIdx.x * APPLY_ROWS + 2 * rg;   // first global row of this thread
    const int nvalid = min(2, max(0, m - grow));         // rows of this thread inside the matrix
    const bool rows_ok = nvalid == 2;
    const bool vec_a = (lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 7) == 0);
    const bool vec_h = Qh && (ldqh % 2 == 0) && ((reinterpret_cast<uintptr_t>(Qh) & 3) == 0);
    pdl_trigger();   // (no pdl_wait: this grid synchronises with the Cholesky grid through flags)
    if (only_if) {   // conditional launch behind the tensor-core apply of a tall panel: runs iff the panel
                     // was factored again from the fp64 Gram matrix (every flag has long been raised)
        pdl_wait();
        if (*only_if == 0) return;
    }

    // acc[0] is always the block being finished, acc[1..] the blocks to its right
    float acc[4][2][8];                          // [block][row][q]: column 4 q + p of the block
#pragma unroll 1
    for (int jb = 0; jb < 4; ++jb) {
        // wait until the Cholesky kernel has published block-row jb of R
        if (tid == 0) {
            // (invariant: the Cholesky grid raises flag[b] only after every write of block-row b, and writes
            // nothing the apply reads after flag[3]; bounded, so that a protocol bug is a launch failure and
            // not a hung GPU)
            unsigned long long spins = 0;
            while (ld_acquire(&fac->flag[jb]) == 0) {
                __nanosleep(64);
                if (++spins > (unsigned long long)(LB_SPIN_LIMIT)) __trap();
            }
        }
        __syncthreads();   // also: everybody has finished reading the previous stage's Rb
        if (jb == 0) {
            // The panel itself is read only now, behind the first flag: the Cholesky kernel has waited for
            // the whole chain of kernels before it (Gram <- update), this grid has not.  A warp touches four
            // columns x 64 contiguous bytes per request.
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float* src = A + grow + (long)(b * 32 + 4 * q + p) * lda;
                    float2 v = make_float2(0.f, 0.f);
                    if (rows_ok && vec_a) {
                        v = *reinterpret_cast<const float2*>(src);
                    } else {
                        if (nvalid > 0) v.x = src[0];
                        if (nvalid > 1) v.y = src[1];
                    }
                    acc[b][0][q] = v.x; acc[b][1][q] = v.y;
                }
        }
        for (int e = tid; e < (4 - jb) * 256; e += APPLY_THREADS) {
            const int d = e >> 8, w4 = e & 255;           // block jb + d of the block-row, float4 index inside it
            const float4* src = reinterpret_cast<const float4*>(
                d == 0 ? &fac->Rdiag[jb][0][0] : &fac->Roff[off_index(jb, jb + d)][0][0]);
            reinterpret_cast<float4*>(&s.Rb[d][0][0])[w4] = __ldcg(src + w4);
        }
        if (tid < 32) s.rinv[tid] = __ldcg(&fac->rinv[jb * 32 + tid]);
        __syncthreads();

        // forward substitution against the diagonal block; column k = 4 kq + kp is owned by part kp
#pragma unroll
        for (int kq = 0; kq < 8; ++kq) {
            float4 da[4], db[4];
            float rv[4];
#pragma unroll
            for (int kp = 0; kp < 4; ++kp) {
                da[kp] = *reinterpret_cast<const float4*>(&s.Rb[0][4 * kq + kp][p * 8]);
                db[kp] = *reinterpret_cast<const float4*>(&s.Rb[0][4 * kq + kp][p * 8 + 4]);
                rv[kp] = s.rinv[4 * kq + kp];
            }
#pragma unroll
            for (int kp = 0; kp < 4; ++kp) {
                const float rr[8] = {da[kp].x, da[kp].y, da[kp].z, da[kp].w,
                                     db[kp].x, db[kp].y, db[kp].z, db[kp].w};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float qk = acc[0][i][kq] * rv[kp];
                    qk = __shfl_sync(0xffffffffu, qk, (lane & ~3) | kp);
                    if (p == kp) acc[0][i][kq] = qk;
#pragma unroll
                    for (int q = kq; q < 8; ++q) {
                        // column 4 q + p is updated iff it lies to the right of column 4 kq + kp
                        if (q > kq || p > kp) acc[0][i][q] = fmaf(-qk, rr[q], acc[0][i][q]);
                    }
                }
            }
        }
        // finished block: to global (fp32 in place + fp16 shadow) and, for the projections, to shared memory
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = jb * 32 + 4 * q + p;
            const float2 v = make_float2(acc[0][0][q], acc[0][1][q]);
            if (jb < 3) *reinterpret_cast<float2*>(&s.Q[4 * q + p][2 * rg]) = v;
            if (nvalid > 0) {
                float* dst = A + grow + (long)c * lda;
                if (rows_ok && vec_a) *reinterpret_cast<float2*>(dst) = v;
                else { dst[0] = v.x; if (nvalid > 1) dst[1] = v.y; }
                if (Qh) {
                    __half* hd = Qh + grow + (long)c * ldqh;
                    const __half2 h01 = __floats2half2_rn(v.x, v.y);
                    if (rows_ok && vec_h) *reinterpret_cast<__half2*>(hd) = h01;
                    else { hd[0] = __low2half(h01); if (nvalid > 1) hd[1] = __high2half(h01); }
                }
            }
        }
        if (jb == 3) break;
        __syncwarp();      // the four parts of a row pair sit in one warp
        if (jb == 0) apply_project<3>(acc, s, rg, p);
        else if (jb == 1) apply_project<2>(acc, s, rg, p);
        else apply_project<1>(acc, s, rg, p);
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[b][i][q] = acc[b + 1][i][q];
    }
}

int gram_grid(int m, int num_sms) {
    const int nchunks = (m + GRAM_ROWS - 1) / GRAM_ROWS;
    return std::max(1, std::min(nchunks, num_sms));
}

struct ScratchLayout {
    size_t part_off, g_off, fac_off, info_off, tc_off, colmax_off, total;
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
    L.tc_off = off;   off += sizeof(TcApplyFactors);
    L.colmax_off = off; off += (size_t)PW * kColmaxParts * sizeof(float);
    L.total = off;
    return L;
}


}  // namespace

size_t panel_scratch_bytes(int m, int num_sms) { return scratch_layout(m, num_sms).total; }

float* panel_colmax_scratch(void* scratch, int m, int num_sms) {
    return reinterpret_cast<float*>(static_cast<uint8_t*>(scratch) + scratch_layout(m, num_sms).colmax_off);
}

cudaError_t panel_init() {
    cudaError_t e = cudaFuncSetAttribute(apply128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(ApplySmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chol128b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Chol2Smem));
    return e != cudaSuccess ? e : tc_apply_init();
}

static cudaError_t launch_chol(cudaStream_t stream, const PanelOpts&, const double* G, float* R, long ldr,
                               PanelFactors* fac, int* info, int col0, int check_redo, double tau, const int* cond) {
    return launch_pdl(chol128b_kernel, dim3(1), dim3(CH2_THREADS), sizeof(Chol2Smem), stream, G, R, ldr, fac, info,
                      col0, check_redo, tau, cond);
}

bool panel_uses_i8_gram(int m, int num_sms, const float* A, long lda, bool allow_tc, const PanelOpts& opts) {
    if (!opts.gram_i8) return false;
    const bool aligned = lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    return allow_tc && aligned && m >= opts.gram_i8_min_rows && panel_gram_i8_fits(m, num_sms);
}

bool panel_uses_tc_apply(int m, const float* A, long lda, bool allow_tc, const PanelOpts& opts) {
    const int ov = opts.apply_tc;
    const bool tc_ok = lda % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    return tc_ok && (ov == 1 || (ov != 0 && allow_tc && m >= kTcApplyMinRows));
}

int panel_launch_count(int m, int num_sms, const float* A, long lda, bool allow_tc, const PanelOpts& opts) {
    const bool tc = panel_uses_tc_apply(m, A, lda, allow_tc, opts);
    const bool i8 = panel_uses_i8_gram(m, num_sms, A, lda, allow_tc, opts);
    return 4 + (tc ? 1 : 0) + (i8 ? 4 : 0) + (i8 && tc ? 1 : 0);
}

cudaError_t panel_qr128(cudaStream_t stream, int num_sms, int m, float* A, long lda, float* R,
                        long ldr, __half* Qh, long ldqh, void* scratch, bool allow_tc,
                        const PanelOpts& opts, int* info, int col0, bool colmax_ready, const PanelComm* comm) {
    const ScratchLayout L = scratch_layout(m, num_sms);
    uint8_t* base = static_cast<uint8_t*>(scratch);
    double* part = reinterpret_cast<double*>(base + L.part_off);
    double* G = reinterpret_cast<double*>(base + L.g_off);
    PanelFactors* fac = reinterpret_cast<PanelFactors*>(base + L.fac_off);
    const int ggrid = gram_grid(m, num_sms);
    const int rgrid = (GRAM_ELEMS + 31) / 32;
    const int* never = nullptr;
    // row-sharded factorisation: G becomes the Gram matrix of the global panel (sum over ranks)
    auto global_sum = [&](const int* only_if = nullptr) -> cudaError_t {
        return comm ? comm->allreduce_f64(comm->self, G, GRAM_ELEMS, stream, only_if) : cudaSuccess;
    };

    cudaError_t le;
    if (panel_uses_i8_gram(m, num_sms, A, lda, allow_tc, opts)) {
        // first attempt from the integer Gram matrix; the Cholesky kernel decides whether it is good
        // enough (info[INFO_REDO]), and the fp64 chain below runs for real only if it is not
        const int igrid = panel_gram_i8_grid(m, num_sms);
        if ((le = panel_gram_i8(stream, num_sms, m, A, lda, reinterpret_cast<float*>(base + L.colmax_off),
                                colmax_ready, part, info + INFO_FLAGS)) != cudaSuccess)
            return le;
        if ((le = launch_pdl(gram128_reduce_kernel, dim3(rgrid), dim3(1024), 0, stream, (const double*)part,
                             igrid, G, fac->flag, never)) != cudaSuccess) return le;
        if ((le = global_sum()) != cudaSuccess) return le;
        if ((le = launch_chol(stream, opts, G, R, ldr, fac, info, col0, 1, opts.i8_fallback_tau, never)) != cudaSuccess)
            return le;
        const int* redo = info + INFO_REDO;
        if ((le = launch_pdl(gram128_f64_kernel, dim3(ggrid), dim3(GRAM_THREADS), 0, stream, (const float*)A,
                             lda, m, part, redo)) != cudaSuccess) return le;
        if ((le = launch_pdl(gram128_reduce_kernel, dim3(rgrid), dim3(1024), 0, stream, (const double*)part,
                             ggrid, G, fac->flag, redo)) != cudaSuccess) return le;
        // (the launch sequence is static; the redo flag is the same on every rank, so the peer-memory
        // kernel returns at once when the panel was not redone - NCCL would sum a stale G that the skipped
        // Cholesky launch below never reads)
        if ((le = global_sum(redo)) != cudaSuccess) return le;
        if ((le = launch_chol(stream, opts, G, R, ldr, fac, info, col0, 0, 0.0, redo)) != cudaSuccess) return le;
    } else {
        if ((le = launch_pdl(gram128_f64_kernel, dim3(ggrid), dim3(GRAM_THREADS), 0, stream, (const float*)A,
                             lda, m, part, never)) != cudaSuccess) return le;
        if ((le = launch_pdl(gram128_reduce_kernel, dim3(rgrid), dim3(1024), 0, stream, (const double*)part,
                             ggrid, G, fac->flag, never)) != cudaSuccess) return le;
        if ((le = global_sum()) != cudaSuccess) return le;
        if ((le = launch_chol(stream, opts, G, R, ldr, fac, info, col0, 0, 0.0, never)) != cudaSuccess) return le;
    }
    const PanelFactors* cfac = fac;
    const dim3 agrid((m + APPLY_ROWS - 1) / APPLY_ROWS);
    if (panel_uses_tc_apply(m, A, lda, allow_tc, opts)) {
        // a panel that was redone in fp64 (integer-Gram panels only) skips the tensor-core apply and
        // is applied by forward substitution
        const int* redone = panel_uses_i8_gram(m, num_sms, A, lda, allow_tc, opts) ? info + INFO_REDO : nullptr;
        le = panel_apply_tc(stream, num_sms, m, A, lda, R, ldr, Qh, ldqh,
                            reinterpret_cast<TcApplyFactors*>(base + L.tc_off), redone);
        if (le != cudaSuccess || !redone) return le;
        le = launch_pdl(apply128_kernel, agrid, dim3(APPLY_THREADS), sizeof(ApplySmem), stream, A, lda, m, cfac,
                        Qh, ldqh, redone);
        return le != cudaSuccess ? le : cudaGetLastError();
    }
    // programmatic dependent launch: the apply grid may start while the Cholesky grid is running
    le = launch_pdl(apply128_kernel, agrid, dim3(APPLY_THREADS), sizeof(ApplySmem), stream, A, lda, m, cfac, Qh,
                    ldqh, never);
    return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace lb
