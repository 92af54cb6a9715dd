// Tensor-core apply for TALL panels: Q = A R^-1 as one split-precision tcgen05 product per
// 128-row tile, so that the step runs at HBM speed instead of fp32-FMA issue speed.
//
//   trinv128_kernel     T = R^-1 (128 x 128 upper triangular) in fp64 on one CTA: 8 x 8 diagonal blocks by
//                       back substitution in registers, the other blocks by block distance on the
//                       fp64 tensor path (mma.sync.m8n8k4.f64).  Emits what the product needs: the column
//                       scales s_k (powers of two bringing ||a_k|| into [2^13, 2^14), so that every
//                       entry of A s fits fp16), the matrix T~ = diag(1/s) T scaled by a power of two
//                       and split into three fp16 planes, and the scalar that undoes that scale.
//   apply128_tc_kernel  persistent, one CTA per SM.  Eight worker warps read a 128 x 128 fp32 tile of
//                       A straight from global memory, split x s_k = hi + lo (both fp16, exact to
//                       2^-22) and write the two planes into shared memory in the canonical
//                       MN-major SWIZZLE_128B operand layout (what TMA would have produced); one
//                       thread issues Ahi T1 + Ahi T2 + Alo T1 + Ahi T3 (32 tcgen05.mma, M128 N128 K16,
//                       fp32 accumulation in TMEM, K = 128 so the truncating accumulate is
//                       harmless); four drain warps read the accumulator with tcgen05.ld and store Q
//                       (fp32, in place) and its fp16 shadow directly - lanes are consecutive rows,
//                       so every store instruction writes one full line.  A buffers and
//                       accumulators are double-buffered: loading tile t+1 overlaps the MMAs of tile t
//                       and the stores of tile t-1, so reads and writes stream concurrently.
//
//   gram128_i8_kernel   the panel's Gram matrix G = A^T A on the INTEGER tensor path, exact in the
//                       sense of the Ozaki scheme: every column is brought to 31-bit fixed point with
//                       a power-of-two scale derived from its largest entry (colmax128_kernel, a
//                       read-only first pass that leaves the panel in L2 when it fits), each
//                       integer is cut into four balanced base-256 digits (one XOR-ADD pair: the
//                       digits are the bytes of (x + 0x00808080) ^ 0x00808080), and the digit planes
//                       D_s are multiplied with tcgen05.mma kind::i8 into four int32 TMEM accumulators,
//                       one per weight 256^(s+t), s + t = 3 .. 6 (the six pairs with s + t < 3 are below
//                       2^-26 of a product and dropped; of the other ten, three are transposes of
//                       pairs already computed and are recovered when the accumulators are drained:
//                       7 pairs = 28 MMAs per 128-row tile).  Integer accumulation is exact,
//                       so the result does not depend on summation order; the accumulators are read
//                       ONCE per CTA, recombined and unscaled in fp64, and handed to the same
//                       fixed-order reduce + fp64 Cholesky as the DMMA Gram kernel of panel.cu.  That
//                       kernel is bounded by the fp64 pipe (64 FMA/clk/SM: 0.87 ms for 2^20 rows);
//                       this one by HBM / the int8 tensor rate.
//
// The fp32 forward-substitution kernel (panel.cu) stays in use for short panels, where it hides
// behind the Cholesky kernel; this path adds the inverse (~15 us) to the dependency chain and only
// pays off when the apply itself is the long pole (m >= kTcApplyMinRows).
#include "panel.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "tc_gemm.cuh"

#include <cstdint>
#include <cstdio>

namespace lb {

using namespace ptx;

namespace {

constexpr int PW = kPanelWidth;            // 128
constexpr int TS_LD = 136;                 // doubles per row of T in shared memory (8 mod 16)
constexpr int TRINV_THREADS = 512;
constexpr int R_PACKED = PW * (PW + 1) / 2;                  // upper triangle, column-packed
constexpr size_t TRINV_SMEM = (size_t)PW * TS_LD * sizeof(double) + (size_t)R_PACKED * sizeof(double) +
                              3 * PW * sizeof(double) + 16 * 64 * sizeof(double) + 64;

__device__ __forceinline__ void dmma_884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// power of two s with x * s in [2^13, 2^14) (1 for x = 0 / non-finite)
__device__ __forceinline__ float pow2_scale(float x) {
    if (!(x > 0.f) || !isfinite(x)) return 1.f;
    int e;
    frexpf(x, &e);
    return ldexpf(1.f, 14 - e);
}

__global__ void __launch_bounds__(TRINV_THREADS, 1)
trinv128_kernel(const float* __restrict__ R, long ldr, TcApplyFactors* __restrict__ out,
                const int* __restrict__ skip) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double* Ts = reinterpret_cast<double*>(smem_raw);                    // T(k, j) at k * TS_LD + j
    double* Rp = Ts + PW * TS_LD;                                        // R(i, k), i <= k, at k (k+1)/2 + i
    float* sc = reinterpret_cast<float*>(Rp + R_PACKED);                 // column scales s_k
    float* red = sc + PW;                                                // block max reduction
    double* Sscr = reinterpret_cast<double*>(red + 32);                  // 16 warps x 64 doubles
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto Rat = [&](int i, int k) -> double { return Rp[(k * (k + 1) >> 1) + i]; };
    pdl_trigger();
    pdl_wait();          // R comes from the Cholesky kernel
    if (skip && *skip != 0) return;      // this panel is applied by forward substitution instead
#ifdef LB_TRINV_TRACE
    long long tr[8]; int ntr = 0;
#define TRINV_MARK() do { __syncthreads(); tr[ntr++] = clock64(); } while (0)
    TRINV_MARK();
#else
#define TRINV_MARK() do {} while (0)
#endif

    // (held in fp64 so that the inner loops below carry no conversions: F2F runs at 16 / clk / SM)
    {   // thread (i, k0): rows i of columns k0, k0 + 4, ...; all 32 loads in flight before the first use
        const int i = tid & (PW - 1), k0 = tid >> 7;
        float r[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = R[i + (long)(k0 + 4 * t) * ldr];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
            const int k = k0 + 4 * t;
            if (i <= k) Rp[(k * (k + 1) >> 1) + i] = (double)r[t];
        }
    }
    __syncthreads();
    TRINV_MARK();
    // column norms of the panel, ||a_k||^2 = sum_i R(i,k)^2, four threads per column
    {
        const int k = tid >> 2, part = tid & 3;
        double s = 0.0;
        for (int i = part; i <= k; i += 4) { const double v = Rat(i, k); s += v * v; }
        s += shfl_xor_f64(s, 1);
        s += shfl_xor_f64(s, 2);
        if (part == 0) sc[k] = pow2_scale((float)sqrt(s));
    }
    TRINV_MARK();
    // Blocked inverse with 8 x 8 blocks (16 per dimension), warp w owning block row w.
    // (1) diagonal blocks: lane j < 8 back-substitutes column j of its block entirely in registers
    if (lane < 8) {
        const int o = warp * 8, j = lane;
        double t[8];
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            double v = 0.0;
            if (i == j) {
                v = 1.0 / Rat(o + i, o + i);
            } else if (i < j) {
                double acc = 0.0;
#pragma unroll
                for (int k = i + 1; k < 8; ++k)
                    if (k <= j) acc += Rat(o + i, o + k) * t[k];
                v = -acc / Rat(o + i, o + i);
            }
            t[i] = v;
            Ts[(o + i) * TS_LD + o + j] = v;                  // (zeros below the diagonal included)
        }
    }
    __syncthreads();
    TRINV_MARK();
    // (2) block distance d = 1 .. 15: T(a,b) = -T(a,a) S,  S = sum_{c = a+1..b} R(a,c) T(c,b), b = a + d, on the
    // fp64 tensor path (mma.sync.m8n8k4.f64: one 8 x 8 tile per warp).  Fragments: A[r = lane / 4][k = lane % 4],
    // B[k = lane % 4][n = lane / 4], C[r = lane / 4][2 (lane % 4) + {0, 1}].
    {
        const int fr = lane >> 2, fk = lane & 3;
        const int a = warp;
        for (int d = 1; d < 16; ++d) {
            const int b = a + d;
            if (b < 16) {
                // (four independent accumulators: the chain of dependent DMMAs is the latency here)
                double c2[2] = {0.0, 0.0}, c3[2] = {0.0, 0.0}, c4[2] = {0.0, 0.0}, c5[2] = {0.0, 0.0};
                int k = 8 * (a + 1);
                for (; k + 16 <= 8 * (b + 1); k += 16) {
                    dmma_884(c2, Rat(8 * a + fr, k + fk), Ts[(k + fk) * TS_LD + 8 * b + fr]);
                    dmma_884(c3, Rat(8 * a + fr, k + 4 + fk), Ts[(k + 4 + fk) * TS_LD + 8 * b + fr]);
                    dmma_884(c4, Rat(8 * a + fr, k + 8 + fk), Ts[(k + 8 + fk) * TS_LD + 8 * b + fr]);
                    dmma_884(c5, Rat(8 * a + fr, k + 12 + fk), Ts[(k + 12 + fk) * TS_LD + 8 * b + fr]);
                }
                if (k < 8 * (b + 1)) {                          // one 8-wide block left (d odd)
                    dmma_884(c2, Rat(8 * a + fr, k + fk), Ts[(k + fk) * TS_LD + 8 * b + fr]);
                    dmma_884(c3, Rat(8 * a + fr, k + 4 + fk), Ts[(k + 4 + fk) * TS_LD + 8 * b + fr]);
                }
                c2[0] = (c2[0] + c3[0]) + (c4[0] + c5[0]);
                c2[1] = (c2[1] + c3[1]) + (c4[1] + c5[1]);
                double* Sw = Sscr + warp * 64;                // S, row-major 8 x 8, private to the warp
                *reinterpret_cast<double2*>(&Sw[fr * 8 + 2 * fk]) = make_double2(c2[0], c2[1]);
                __syncwarp();
                double e2[2] = {0.0, 0.0};
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    dmma_884(e2, Ts[(8 * a + fr) * TS_LD + 8 * a + 4 * h + fk], Sw[(4 * h + fk) * 8 + fr]);
                *reinterpret_cast<double2*>(&Ts[(8 * a + fr) * TS_LD + 8 * b + 2 * fk]) = make_double2(-e2[0], -e2[1]);
            }
            __syncthreads();
        }
    }
    TRINV_MARK();
    // T~ = diag(1/s) T, its power-of-two scale, and the three fp16 planes (column-major: k contiguous)
    // Pass 1 (lanes along j: conflict-free reads of T(k, :)): T~(k,j) = T(k,j) / s_k as fp32 into a
    // staging array laid out [j][k] (it reuses R's storage), and the largest magnitude.
    float* stage = reinterpret_cast<float*>(Rp);             // 128 x 129 floats = exactly R's 66 048 bytes
    float mx = 0.f;
    for (int idx = tid; idx < PW * PW; idx += TRINV_THREADS) {
        const int j = idx & (PW - 1), k = idx >> 7;
        const float v = k <= j ? (float)(Ts[k * TS_LD + j] * (double)(1.f / sc[k])) : 0.f;   // exact scaling
        stage[j * (PW + 1) + k] = v;
        mx = fmaxf(mx, fabsf(v));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = 0.f;
    for (int w = 0; w < TRINV_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
    const float st = pow2_scale(mx);
    // Pass 2 (lanes along k: coalesced stores): three fp16 planes of T~ * st, column-major
    for (int idx = tid; idx < PW * PW; idx += TRINV_THREADS) {
        const int k = idx & (PW - 1), j = idx >> 7;
        const float v = stage[j * (PW + 1) + k] * st;
        const __half t1 = __float2half_rn(v);
        const float r1 = v - __half2float(t1);
        const __half t2 = __float2half_rn(r1);
        out->T[0][idx] = t1;
        out->T[1][idx] = t2;
        out->T[2][idx] = __float2half_rn(r1 - __half2float(t2));
    }
    if (tid < PW) out->colscale[tid] = sc[tid];
    if (tid == 0) out->unscale = 1.f / st;
#ifdef LB_TRINV_TRACE
    TRINV_MARK();
    if (tid == 0)
        printf("trinv cycles: load %lld norms %lld diag %lld offdiag %lld output %lld\n", tr[1] - tr[0],
               tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4]);
#endif
}

// ---------------------------------------------------------------------------------------------
constexpr int TCA_WORKERS = 256;                    // 8 warps: load + split + stage
constexpr int TCA_DRAIN = 128;                      // 4 warps: TMEM -> Q (fp32) and its fp16 shadow
constexpr int TCA_THREADS = TCA_WORKERS + 32 + TCA_DRAIN;   // warp 8 = MMA issuer
constexpr int KB_BYTES = PW * 64 * 2;               // one 64-deep k-block of a 128-wide fp16 operand
constexpr int PLANE_BYTES = 2 * KB_BYTES;           // K = 128: 32 KiB per plane
constexpr int T_BYTES = 3 * PLANE_BYTES;            // t1 + t2 + t3
constexpr int A_BUF_BYTES = 2 * PLANE_BYTES;        // hi + lo of one tile
constexpr int TCA_SMEM = T_BYTES + 2 * A_BUF_BYTES + PW * 4 + 128 + 1024;

__global__ void __launch_bounds__(TCA_THREADS, 1)
apply128_tc_kernel(const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ CUtensorMap mapT2,
                   const __grid_constant__ CUtensorMap mapT3,
                   float* __restrict__ A, long lda, int m, const TcApplyFactors* __restrict__ fac,
                   __half* __restrict__ Qh, long ldqh, const int* __restrict__ skip) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t t_base = smem_base;
    const uint32_t a_base = smem_base + T_BYTES;
    float* colscale = reinterpret_cast<float*>(smem_gen + T_BYTES + 2 * A_BUF_BYTES);
    const uint32_t bar_base = smem_base + T_BYTES + 2 * A_BUF_BYTES + PW * 4;
    const uint32_t t_full = bar_base;
    auto a_full = [&](int b) { return bar_base + 8u * (1 + b); };
    auto a_empty = [&](int b) { return bar_base + 8u * (3 + b); };
    auto acc_full = [&](int b) { return bar_base + 8u * (5 + b); };
    auto acc_empty = [&](int b) { return bar_base + 8u * (7 + b); };
    const uint32_t tmem_slot = bar_base + 8u * 9;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int tiles = (m + PW - 1) / PW;

    if (warp == 8) {
        if (lane == 0) {
            prefetch_tensormap(&mapT1);
            prefetch_tensormap(&mapT2);
            prefetch_tensormap(&mapT3);
            mbar_init(t_full, 1);
            for (int b = 0; b < 2; ++b) {
                mbar_init(a_full(b), TCA_WORKERS);
                mbar_init(a_empty(b), 1);
                mbar_init(acc_full(b), 1);
                mbar_init(acc_empty(b), TCA_DRAIN);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();          // T planes, scales (trinv kernel) and A (whatever produced the panel)
    const bool skipped = skip && *skip != 0;     // the panel is applied by forward substitution instead
    if (skipped) tiles = 0;

    if (warp == 8) {
        // ------------------------------------------------------------------ T loader + MMA issuer
        if (lane == 0 && !skipped) {
            mbar_arrive_expect_tx(t_full, T_BYTES);
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(t_base + kb * KB_BYTES, &mapT1, t_full, kb * 64, 0);
                tma_load_2d(t_base + PLANE_BYTES + kb * KB_BYTES, &mapT2, t_full, kb * 64, 0);
                tma_load_2d(t_base + 2 * PLANE_BYTES + kb * KB_BYTES, &mapT3, t_full, kb * 64, 0);
            }
            mbar_wait(t_full, 0);
            constexpr uint32_t idesc = make_idesc(/*F16*/ 0, /*A MN-major*/ 1u, 0u, PW, PW);
            int n = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n) {
                const int b = n & 1;
                const uint32_t ph = (uint32_t)(n >> 1) & 1u;
                mbar_wait(a_full(b), ph);
                mbar_wait(acc_empty(b), ph ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + b * PW;
                const uint32_t a_buf = a_base + b * A_BUF_BYTES;
                // (A plane, T plane), smallest terms first: hi t3, lo t1, hi t2, hi t1.  The tensor core
                // truncates when it adds into the fp32 accumulator, a bias of up to one ulp of the
                // running sum per MMA: with the 2^-11 / 2^-22 terms accumulated first only the last
                // eight steps work on a full-size sum (measured |Q^T Q - I|_F of a panel: 8e-6 with the
                // main term first).  T is carried to 2^-33 (t3) because its error is the same for
                // every row and does not average out in Q^T Q; A's split (2^-22) is row-wise.
                const int pa[4] = {0, 1, 0, 0}, pt[4] = {2, 0, 1, 0};
                bool first = true;
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t a_desc = make_smem_desc_sw128(
                            a_buf + pa[pass] * PLANE_BYTES + kb * KB_BYTES, 64 * 64 * 2, 1024);
                        const uint64_t b_desc = make_smem_desc_sw128(
                            t_base + pt[pass] * PLANE_BYTES + kb * KB_BYTES, 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16(d_tmem, a_desc + k * (16 * 128 / 16), b_desc + k * (16 * 2 / 16), idesc,
                                     first ? 0u : 1u);
                            first = false;
                        }
                    }
                }
                umma_commit(a_empty(b));
                umma_commit(acc_full(b));
            }
        }
    } else if (warp > 8) {
        // ------------------------------------------------------------------ drain warps
        const float unscale = fac->unscale;
        const int quad = warp & 3;              // TMEM lane quarter this warp may read
        int n = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n) {
            const int b = n & 1;
            mbar_wait(acc_full(b), (uint32_t)(n >> 1) & 1u);
            tc_fence_after_sync();
            const int row = tile * PW + quad * 32 + lane;
            const bool row_ok = row < m;
            const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + b * PW;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t d[32];
                tmem_ld_32x32(t_addr + c * 32, d);
                tmem_ld_wait();
                if (row_ok) {
                    // lanes are consecutive rows: every store instruction writes one full line
                    float* dst = A + row + (long)(c * 32) * lda;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(d[j]) * unscale;
                        dst[(long)j * lda] = v;
                        if (Qh) Qh[row + (long)(c * 32 + j) * ldqh] = __float2half_rn(v);
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(acc_empty(b));
        }
    } else {
        // ------------------------------------------------------------------ convert warps
        if (tid < PW) colscale[tid] = fac->colscale[tid];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int g = tid & 15;                 // 8-row group of the tile
        const int cbase = tid >> 4;             // columns cbase + 16 i
        int n = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n) {
            const int b = n & 1;
            mbar_wait(a_empty(b), ((uint32_t)(n >> 1) & 1u) ^ 1u);
            // this thread's eight 8-row x 1-column chunks of the tile, all loads in flight at once
            const int row0 = tile * PW + g * 8;
            const bool ok = row0 < m;              // m is a multiple of 8: a chunk is in or out
            float4 v[8][2];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float* src = A + row0 + (long)(cbase + 16 * i) * lda;
                if (ok) {
                    v[i][0] = *reinterpret_cast<const float4*>(src);
                    v[i][1] = *(reinterpret_cast<const float4*>(src) + 1);
                } else {
                    v[i][0] = v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            uint8_t* buf = smem_gen + T_BYTES + b * A_BUF_BYTES;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = cbase + 16 * i;
                const float s = colscale[c];
                const float x[8] = {v[i][0].x * s, v[i][0].y * s, v[i][0].z * s, v[i][0].w * s,
                                    v[i][1].x * s, v[i][1].y * s, v[i][1].z * s, v[i][1].w * s};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __half2 h = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
                    hi[q] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[q] = *reinterpret_cast<const uint32_t*>(&l);
                }
                // MN-major SWIZZLE_128B: k-block, 64-row half, k row of 128 B, 16-byte chunk ^ (k & 7)
                const int kk = c & 63;
                const uint32_t off = (uint32_t)(c >> 6) * KB_BYTES + (uint32_t)(g >> 3) * (64 * 64 * 2) +
                                     (uint32_t)kk * 128 + (uint32_t)(((g & 7) ^ (kk & 7)) << 4);
                *reinterpret_cast<uint4*>(buf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(buf + PLANE_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core
            mbar_arrive(a_full(b));
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 256);
    }
}


// ---------------------------------------------------------------------------------------------
// tcgen05.mma kind::i8: signed 8-bit operands, int32 accumulate, K = 32 per instruction.
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor: D = S32 (2 at [4,6)), A and B signed 8 bit (1 at [7,10) and [10,13)),
// both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((PW >> 3) << 17) | ((PW >> 4) << 24);

constexpr int GI8_WORKERS = 256;
constexpr int GI8_THREADS = GI8_WORKERS + 32;
constexpr int DIGIT_PLANE_BYTES = PW * 128;                  // 128 columns x 128 rows of int8
constexpr int DIGIT_BUF_BYTES = 4 * DIGIT_PLANE_BYTES;       // four digit planes of one 128-row tile
constexpr int GI8_SMEM = 2 * DIGIT_BUF_BYTES + PW * 4 + PW * 8 + 128 + 1024;
constexpr int GI8_MAX_TILES_PER_CTA = 256;                   // 2^15 rows x (4 pairs x 2^14) < 2^31
constexpr int NTRI_BLOCKS = 10, GBLK = 32, G_ELEMS = NTRI_BLOCKS * GBLK * GBLK;

__host__ __device__ inline int tri_index4(int bi, int bj) { return bi * 4 - bi * (bi - 1) / 2 + (bj - bi); }

// max_i |a_ic| of every column, as kColmaxParts per-block partial maxima (no atomics, no zeroing).
__global__ void __launch_bounds__(256)
colmax128_kernel(const float* __restrict__ A, long lda, int m, float* __restrict__ colmax_part) {
    __shared__ float red[8];
    const int c = blockIdx.y;
    pdl_trigger();
    pdl_wait();
    const float* src = A + (long)c * lda;
    float mx = 0.f;
    for (int i = (blockIdx.x * 256 + threadIdx.x) * 4; i < m; i += gridDim.x * 256 * 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);      // m % 8 == 0, aligned
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
        colmax_part[c * kColmaxParts + blockIdx.x] = mx;
    }
    if (blockIdx.x == 0)       // slots no block of this grid owns
        for (int sidx = gridDim.x + threadIdx.x; sidx < kColmaxParts; sidx += 256) colmax_part[c * kColmaxParts + sidx] = 0.f;
}

__global__ void __launch_bounds__(GI8_THREADS, 1)
gram128_i8_kernel(const float* __restrict__ A, long lda, int m, const float* __restrict__ colmax_part,
                  double* __restrict__ part, int* __restrict__ info) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* scale = reinterpret_cast<float*>(smem_gen + 2 * DIGIT_BUF_BYTES);           // 2^E_c
    double* unscale = reinterpret_cast<double*>(smem_gen + 2 * DIGIT_BUF_BYTES + PW * 4);   // 2^-E_c
    const uint32_t bar_base = smem_base + 2 * DIGIT_BUF_BYTES + PW * 4 + PW * 8;
    auto a_full = [&](int b) { return bar_base + 8u * b; };
    auto a_empty = [&](int b) { return bar_base + 8u * (2 + b); };
    const uint32_t acc_done = bar_base + 8u * 4;
    const uint32_t tmem_slot = bar_base + 8u * 5;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles = (m + PW - 1) / PW;

    if (warp == 8) {
        if (lane == 0) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(a_full(b), GI8_WORKERS);
                mbar_init(a_empty(b), 1);
            }
            mbar_init(acc_done, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();          // the panel and the column-norm partials come from predecessors

    if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int n = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n) {
                const int b = n & 1;
                mbar_wait(a_full(b), (uint32_t)(n >> 1) & 1u);
                tc_fence_after_sync();
                const uint32_t buf = smem_base + b * DIGIT_BUF_BYTES;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (s + t < 3) continue;           // weight 256^(s+t): below 2^-26 of a product
                        // D_t^T D_s = (D_s^T D_t)^T: the groups made of such mirror pairs only (s + t = 3
                        // and 5) get one of each and are symmetrised when the accumulators are drained
                        if (s > t && ((s + t) & 1)) continue;
                        const uint64_t a_desc = make_smem_desc_sw128(buf + s * DIGIT_PLANE_BYTES, 16, 1024);
                        const uint64_t b_desc = make_smem_desc_sw128(buf + t * DIGIT_PLANE_BYTES, 16, 1024);
                        // the first pair of every group in issue order: (0,3), (1,3), (2,3), (3,3)
                        const bool group_start = (t == 3);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_i8(tmem_base + (s + t - 3) * PW, a_desc + 2 * k, b_desc + 2 * k, kIdescI8,
                                    (n == 0 && group_start && k == 0) ? 0u : 1u);
                    }
                }
                umma_commit(a_empty(b));
            }
            umma_commit(acc_done);
        }
    } else {
        // ------------------------------------------------------------------ workers
        if (tid < PW) {
            float bound = 0.f;                                           // max_i |a_ic|
            for (int b = 0; b < kColmaxParts; ++b) bound = fmaxf(bound, colmax_part[tid * kColmaxParts + b]);
            int e = 0;
            if (bound > 0.f && isfinite(bound)) frexpf(bound, &e);       // bound < 2^e
            int E = 30 - e;                                              // |a| 2^E < 2^30
            E = max(-96, min(96, E));
            scale[tid] = ldexpf(1.f, E);
            unscale[tid] = ldexp(1.0, -E);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int q = tid & 7;                  // 16-row chunk of the tile
        const int cb = tid >> 3;                // columns cb + 32 u
        bool overflow = false;
        // this thread's four chunks (16 rows x 1 column each) of a tile: 16 loads in flight at once
        auto issue = [&](int tile, float4 (&v)[4][4]) {
            const int row0 = tile * PW + q * 16;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* src = A + row0 + (long)(cb + 32 * u) * lda;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    v[u][w] = row0 + 4 * w < m ? *reinterpret_cast<const float4*>(src + 4 * w)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // One pipeline step: put the NEXT tile's loads in flight (second register buffer), then
        // convert the current one - with a single buffer every tile paid the full DRAM latency
        // (measured 2.9 us per tile where the 28 MMAs need 1.75 and the 64 KB of HBM traffic 1.45).
        auto step = [&](int tile, int n, float4 (&cur)[4][4], float4 (&nxt)[4][4]) {
            if (tile + (int)gridDim.x < tiles) issue(tile + gridDim.x, nxt);
            const int b = n & 1;
            mbar_wait(a_empty(b), ((uint32_t)(n >> 1) & 1u) ^ 1u);
            uint8_t* buf = smem_gen + b * DIGIT_BUF_BYTES;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + 32 * u;
                const float s = scale[c];
                uint32_t pl[4][4];           // [digit plane][word of 4 consecutive rows]
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float f[4] = {cur[u][w].x * s, cur[u][w].y * s, cur[u][w].z * s, cur[u][w].w * s};
                    uint32_t z[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        overflow |= !(fabsf(f[i]) <= 1073741824.f);     // (also true for NaN)
                        // balanced base-256 digits = bytes of (x + 0x00808080) ^ 0x00808080
                        z[i] = ((uint32_t)__float2int_rn(f[i]) + 0x00808080u) ^ 0x00808080u;
                    }
                    const uint32_t lo01 = __byte_perm(z[0], z[1], 0x5140), lo23 = __byte_perm(z[2], z[3], 0x5140);
                    const uint32_t hi01 = __byte_perm(z[0], z[1], 0x7362), hi23 = __byte_perm(z[2], z[3], 0x7362);
                    pl[0][w] = __byte_perm(lo01, lo23, 0x5410);
                    pl[1][w] = __byte_perm(lo01, lo23, 0x7632);
                    pl[2][w] = __byte_perm(hi01, hi23, 0x5410);
                    pl[3][w] = __byte_perm(hi01, hi23, 0x7632);
                }
                // K-major SWIZZLE_128B: one 128-byte row (128 consecutive matrix rows) per column
                const uint32_t off = (uint32_t)c * 128 + (uint32_t)((q ^ (c & 7)) << 4);
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    *reinterpret_cast<uint4*>(buf + d * DIGIT_PLANE_BYTES + off) =
                        make_uint4(pl[d][0], pl[d][1], pl[d][2], pl[d][3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(a_full(b));
        };
        {
            float4 v0[4][4], v1[4][4];
            int n = 0, tile = blockIdx.x;
            if (tile < tiles) issue(tile, v0);
            while (tile < tiles) {
                step(tile, n, v0, v1);
                tile += gridDim.x; ++n;
                if (tile >= tiles) break;
                step(tile, n, v1, v0);
                tile += gridDim.x; ++n;
            }
        }
        if (overflow) atomicOr(info, 1);        // INFO_FLAGS bit 0 (Inf / NaN in the panel)

        // ---- drain: G(j,k) 2^(E_j + E_k) = sum_g acc_g(j,k) 256^(g+3), upper 32 x 32 blocks only.
        // Groups 0 and 2 (s + t = 3, 5) hold one pair of each mirror couple: X = acc_0 256^3 + acc_2 256^5
        // (exact in fp64: 47 bits) is staged in shared memory (the digit buffers, now free; 128 x 128
        // doubles, XOR-swizzled so that row writes and column reads are both conflict-free) and
        // G gets X + X^T.
        mbar_wait(acc_done, 0);
        tc_fence_after_sync();
        const int quad = warp & 3, half = warp >> 2;
        const int j = quad * 32 + lane;
        double* X = reinterpret_cast<double*>(smem_gen);
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int k0 = (2 * half + cc) * 32;
            uint32_t d0[32], d2[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + 0 * PW + k0, d0);
            tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + 2 * PW + k0, d2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                X[j * PW + ((k0 + i) ^ (j & 31))] =
                    fma((double)(int)d2[i], 1099511627776.0 /* 256^5 */, (double)(int)d0[i] * 16777216.0 /* 256^3 */);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const double uj = unscale[j];
        double* dst_cta = part + (long)blockIdx.x * G_ELEMS;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            const int bj = 2 * half + cc;
            if (bj < quad) continue;                // strictly lower block: never read
            const int k0 = bj * 32;
            double sum[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
                sum[i] = X[j * PW + ((k0 + i) ^ (j & 31))] + X[(k0 + i) * PW + (j ^ ((k0 + i) & 31))];
#pragma unroll
            for (int g = 1; g < 4; g += 2) {        // the complete groups: s + t = 4 and 6
                uint32_t d[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + g * PW + k0, d);
                tmem_ld_wait();
                const double wgt = (double)(1ull << (8 * (g + 3)));
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[i] = fma((double)(int)d[i], wgt, sum[i]);
            }
            double* dst = dst_cta + tri_index4(quad, bj) * (GBLK * GBLK) + lane * GBLK;
#pragma unroll
            for (int i = 0; i < 32; i += 2)
                *reinterpret_cast<double2*>(dst + i) =
                    make_double2(sum[i] * uj * unscale[k0 + i], sum[i + 1] * uj * unscale[k0 + i + 1]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int panel_gram_i8_grid(int m, int num_sms) {
    const int tiles = (m + PW - 1) / PW;
    return tiles < num_sms ? tiles : num_sms;
}

bool panel_gram_i8_fits(int m, int num_sms) {
    const int tiles = (m + PW - 1) / PW, grid = panel_gram_i8_grid(m, num_sms);
    return (tiles + grid - 1) / grid <= GI8_MAX_TILES_PER_CTA;
}

cudaError_t panel_gram_i8(cudaStream_t stream, int num_sms, int m, const float* A, long lda,
                          float* colmax_part, bool colmax_ready, double* part, int* info) {
    cudaError_t e = cudaSuccess;
    if (!colmax_ready)
        e = launch_pdl(colmax128_kernel, dim3(64, PW), dim3(256), 0, stream, A, lda, m, colmax_part);
    if (e != cudaSuccess) return e;
    e = launch_pdl(gram128_i8_kernel, dim3(panel_gram_i8_grid(m, num_sms)), dim3(GI8_THREADS),
                   (size_t)GI8_SMEM, stream, A, lda, m, (const float*)colmax_part, part, info);
    return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t tc_apply_init() {
    cudaError_t e = cudaFuncSetAttribute(gram128_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GI8_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(trinv128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)TRINV_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(apply128_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCA_SMEM);
}

cudaError_t panel_apply_tc(cudaStream_t stream, int num_sms, int m, float* A, long lda, const float* R,
                           long ldr, __half* Qh, long ldqh, TcApplyFactors* fac, const int* skip) {
    cudaError_t e = launch_pdl(trinv128_kernel, dim3(1), dim3(TRINV_THREADS), TRINV_SMEM, stream, R, ldr, fac, skip);
    if (e != cudaSuccess) return e;
    CUtensorMap t1, t2, t3;
    HalfMatrix m1{fac->T[0], PW, PW, PW}, m2{fac->T[1], PW, PW, PW}, m3{fac->T[2], PW, PW, PW};
    if ((e = make_tensor_map_f16(&t1, m1, 64, PW)) != cudaSuccess) return e;
    if ((e = make_tensor_map_f16(&t2, m2, 64, PW)) != cudaSuccess) return e;
    if ((e = make_tensor_map_f16(&t3, m3, 64, PW)) != cudaSuccess) return e;
    const int tiles = (m + PW - 1) / PW;
    const int grid = tiles < num_sms ? tiles : num_sms;
    e = launch_pdl(apply128_tc_kernel, dim3(grid), dim3(TCA_THREADS), (size_t)TCA_SMEM, stream, t1, t2, t3, A,
                   lda, m, (const TcApplyFactors*)fac, Qh, ldqh, skip);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace lb
