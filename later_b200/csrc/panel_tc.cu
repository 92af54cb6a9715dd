// Tensor-core apply for TALL panels: Q = A R^-1 as one split-precision tcgen05 product per
// 128-row tile, so that the step runs at HBM speed instead of fp32-FMA issue speed.
//
//   trinv128_kernel     T = R^-1 (128 x 128 upper triangular) by back substitution in fp64, one CTA,
//                       four threads per column of T.  Emits what the product needs: the column
//                       scales s_k (powers of two bringing ||a_k|| into [2^13, 2^14), so that every
//                       entry of A s fits fp16), the matrix T~ = diag(1/s) T scaled by a power of two
//                       and split into three fp16 planes, and the scalar that undoes that scale.
//   apply128_tc_kernel  persistent, one CTA per SM.  Eight worker warps read a 128 x 128 fp32 tile of
//                       A straight from global memory, split x s_k = hi + lo (both fp16, exact to
//                       2^-22) and write the two planes into shared memory in the canonical
//                       MN-major SWIZZLE_128B operand layout (what TMA would have produced); one
//                       thread issues Ahi T1 + Ahi T2 + Alo T1 + Ahi T3 (32 tcgen05.mma, M128 N128 K16,
//                       fp32 accumulation in TMEM, K = 128 so the truncating accumulate is
//                       harmless); the same eight warps drain the previous tile's accumulator with
//                       tcgen05.ld and store Q (fp32, in place) and its fp16 shadow directly -
//                       lanes are consecutive rows, so every store instruction writes one full line.
//                       A buffers and accumulators are double-buffered: converting tile t+1 overlaps
//                       the MMAs of tile t and the drain of tile t-1.
//
// The fp32 forward-substitution kernel (panel.cu) stays in use for short panels, where it hides
// behind the Cholesky kernel; this path adds the inverse (~15 us) to the dependency chain and only
// pays off when the apply itself is the long pole (m >= kTcApplyMinRows).
#include "panel.cuh"
#include "launch.cuh"
#include "ptx.cuh"
#include "tc_gemm.cuh"

#include <cstdint>

namespace lb {

using namespace ptx;

namespace {

constexpr int PW = kPanelWidth;            // 128
constexpr int TS_LD = 136;                 // doubles per row of T in shared memory (8 mod 16)
constexpr int RS_LD = 129;                 // floats per column of R in shared memory
constexpr int TRINV_THREADS = 512;
constexpr size_t TRINV_SMEM = (size_t)PW * TS_LD * sizeof(double) + (size_t)PW * RS_LD * sizeof(float) +
                              3 * PW * sizeof(double) + 64;

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
trinv128_kernel(const float* __restrict__ R, long ldr, TcApplyFactors* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double* Ts = reinterpret_cast<double*>(smem_raw);                    // T[k][j] at k * TS_LD + j
    float* Rs = reinterpret_cast<float*>(Ts + PW * TS_LD);               // R(i, k) at k * RS_LD + i
    double* rinv = reinterpret_cast<double*>(Rs + PW * RS_LD);         // (66048 B: still 8-byte aligned)       // 1 / R(i, i)
    float* sc = reinterpret_cast<float*>(rinv + PW);                     // column scales s_k
    float* red = sc + PW;                                                // block max reduction
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    pdl_wait();          // R comes from the Cholesky kernel

    for (int idx = tid; idx < PW * PW; idx += TRINV_THREADS) {
        const int i = idx & (PW - 1), k = idx >> 7;
        Rs[k * RS_LD + i] = i <= k ? R[i + (long)k * ldr] : 0.f;
    }
    for (int idx = tid; idx < PW * TS_LD; idx += TRINV_THREADS) Ts[idx] = 0.0;   // (lower parts stay 0)
    __syncthreads();
    // column norms of the panel, ||a_k||^2 = sum_i R(i,k)^2, four threads per column
    {
        const int k = tid >> 2, part = tid & 3;
        double s = 0.0;
        for (int i = part; i <= k; i += 4) { const double v = Rs[k * RS_LD + i]; s += v * v; }
        s += shfl_xor_f64(s, 1);
        s += shfl_xor_f64(s, 2);
        if (part == 0) {
            sc[k] = pow2_scale((float)sqrt(s));
            rinv[k] = 1.0 / (double)Rs[k * RS_LD + k];
        }
    }
    __syncthreads();
    // (1) the four 32 x 32 diagonal blocks of T by back substitution, column j by the four lanes
    // {jj, 8 + jj, 16 + jj, 24 + jj} of warp j / 8:
    //     T(i,j) = -(sum_{k = i+1..j} R(i,k) T(k,j)) / R(i,i),   the sum split over k mod 4
    {
        const int jj = lane & 7, part = lane >> 3;
        const int j = warp * 8 + jj;
        const int i0 = (warp >> 2) * 32;                       // first row of this column's diagonal block
        if (part == 0) Ts[j * TS_LD + j] = rinv[j];
        __syncwarp();
        for (int i = warp * 8 + 6; i >= i0; --i) {
            double s0 = 0.0, s1 = 0.0;
            if (i < j) {
                int k = i + 1 + ((part - (i + 1)) & 3);      // first k > i with k = part (mod 4)
                for (; k + 4 <= j; k += 8) {
                    s0 += (double)Rs[k * RS_LD + i] * Ts[k * TS_LD + j];
                    s1 += (double)Rs[(k + 4) * RS_LD + i] * Ts[(k + 4) * TS_LD + j];
                }
                if (k <= j) s0 += (double)Rs[k * RS_LD + i] * Ts[k * TS_LD + j];
            }
            double s = s0 + s1;
            s += shfl_xor_f64(s, 8);
            s += shfl_xor_f64(s, 16);
            if (part == 0 && i < j) Ts[i * TS_LD + j] = -s * rinv[i];
            __syncwarp();
        }
    }
    __syncthreads();
    // (2) off-diagonal blocks by block distance d: T(a,b) = -T(a,a) S,  S = sum_{c = a+1..b} R(a,c) T(c,b).
    // Dense 32-wide products, every entry independent; S is parked in the unused mirror block (b,a).
    for (int d = 1; d < 4; ++d) {
        const int nblk = 4 - d;                                // blocks (a, a + d), a = 0 .. 3 - d
        for (int e = tid; e < nblk * 1024; e += TRINV_THREADS) {
            const int a = e >> 10, b = a + d, i = (e >> 5) & 31, j = e & 31;
            double s0 = 0.0, s1 = 0.0;
            const float* rrow = Rs + (32 * a + i);             // R(32a + i, k) at k * RS_LD
            const double* tcol = Ts + (32 * b + j);            // T(k, 32b + j) at k * TS_LD
            for (int k = 32 * (a + 1); k < 32 * (b + 1); k += 2) {
                s0 += (double)rrow[k * RS_LD] * tcol[k * TS_LD];
                s1 += (double)rrow[(k + 1) * RS_LD] * tcol[(k + 1) * TS_LD];
            }
            Ts[(32 * b + i) * TS_LD + 32 * a + j] = s0 + s1;    // S(i, j) in the mirror block
        }
        __syncthreads();
        for (int e = tid; e < nblk * 1024; e += TRINV_THREADS) {
            const int a = e >> 10, b = a + d, i = (e >> 5) & 31, j = e & 31;
            double s0 = 0.0;
            for (int k = i; k < 32; ++k)                       // T(a,a) is upper triangular
                s0 += Ts[(32 * a + i) * TS_LD + 32 * a + k] * Ts[(32 * b + k) * TS_LD + 32 * a + j];
            Ts[(32 * a + i) * TS_LD + 32 * b + j] = -s0;
        }
        __syncthreads();
    }
    // T~ = diag(1/s) T, its power-of-two scale, and the two fp16 planes (column-major: k contiguous)
    float mx = 0.f;
    for (int idx = tid; idx < PW * PW; idx += TRINV_THREADS) {
        const int k = idx & (PW - 1), j = idx >> 7;
        if (k <= j) mx = fmaxf(mx, fabsf((float)(Ts[k * TS_LD + j] * (double)(1.f / sc[k]))));
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = 0.f;
    for (int w = 0; w < TRINV_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
    const float st = pow2_scale(mx);
    for (int idx = tid; idx < PW * PW; idx += TRINV_THREADS) {
        const int k = idx & (PW - 1), j = idx >> 7;
        // (1 / s_k and st are powers of two: the two scalings are exact)
        const double vd = k <= j ? Ts[k * TS_LD + j] * (double)(1.f / sc[k]) * (double)st : 0.0;
        const float v = (float)vd;
        const __half t1 = __float2half_rn(v);
        const float r1 = v - __half2float(t1);
        const __half t2 = __float2half_rn(r1);
        out->T[0][idx] = t1;
        out->T[1][idx] = t2;
        out->T[2][idx] = __float2half_rn(r1 - __half2float(t2));
    }
    if (tid < PW) out->colscale[tid] = sc[tid];
    if (tid == 0) out->unscale = 1.f / st;
}

// ---------------------------------------------------------------------------------------------
constexpr int TCA_WORKERS = 256;                    // 8 warps: convert + drain
constexpr int TCA_THREADS = TCA_WORKERS + 32;       // + 1 MMA warp
constexpr int KB_BYTES = PW * 64 * 2;               // one 64-deep k-block of a 128-wide fp16 operand
constexpr int PLANE_BYTES = 2 * KB_BYTES;           // K = 128: 32 KiB per plane
constexpr int T_BYTES = 3 * PLANE_BYTES;            // t1 + t2 + t3
constexpr int A_BUF_BYTES = 2 * PLANE_BYTES;        // hi + lo of one tile
constexpr int TCA_SMEM = T_BYTES + 2 * A_BUF_BYTES + PW * 4 + 128 + 1024;

__global__ void __launch_bounds__(TCA_THREADS, 1)
apply128_tc_kernel(const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ CUtensorMap mapT2,
                   const __grid_constant__ CUtensorMap mapT3,
                   float* __restrict__ A, long lda, int m, const TcApplyFactors* __restrict__ fac,
                   __half* __restrict__ Qh, long ldqh) {
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
    const int tiles = (m + PW - 1) / PW;

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
                mbar_init(acc_empty(b), TCA_WORKERS);
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

    if (warp == 8) {
        // ------------------------------------------------------------------ T loader + MMA issuer
        if (lane == 0) {
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
    } else {
        // ------------------------------------------------------------------ workers
        if (tid < PW) colscale[tid] = fac->colscale[tid];
        const float unscale = fac->unscale;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int g = tid & 15;                 // 8-row group of the tile
        const int cbase = tid >> 4;             // columns cbase + 16 i
        const int quad = warp & 3, half = warp >> 2;

        auto drain = [&](int tile, int n) {
            const int b = n & 1;
            mbar_wait(acc_full(b), (uint32_t)(n >> 1) & 1u);
            tc_fence_after_sync();
            const int row = tile * PW + quad * 32 + lane;
            const bool row_ok = row < m;
            const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + b * PW + half * 64;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t d[32];
                tmem_ld_32x32(t_addr + c * 32, d);
                tmem_ld_wait();
                if (row_ok) {
                    const int col0 = half * 64 + c * 32;
                    float* dst = A + row + (long)col0 * lda;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(d[j]) * unscale;
                        dst[(long)j * lda] = v;
                        if (Qh) Qh[row + (long)(col0 + j) * ldqh] = __float2half_rn(v);
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(acc_empty(b));
        };

        int n = 0, prev_tile = -1;
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
            if (prev_tile >= 0) drain(prev_tile, n - 1);
            prev_tile = tile;
        }
        if (prev_tile >= 0) drain(prev_tile, n - 1);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

size_t tc_apply_scratch_bytes() { return sizeof(TcApplyFactors); }

cudaError_t tc_apply_init() {
    cudaError_t e = cudaFuncSetAttribute(trinv128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)TRINV_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(apply128_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCA_SMEM);
}

cudaError_t panel_apply_tc(cudaStream_t stream, int num_sms, int m, float* A, long lda, const float* R,
                           long ldr, __half* Qh, long ldqh, TcApplyFactors* fac) {
    cudaError_t e = launch_pdl(trinv128_kernel, dim3(1), dim3(TRINV_THREADS), TRINV_SMEM, stream, R, ldr, fac);
    if (e != cudaSuccess) return e;
    CUtensorMap t1, t2, t3;
    HalfMatrix m1{fac->T[0], PW, PW, PW}, m2{fac->T[1], PW, PW, PW}, m3{fac->T[2], PW, PW, PW};
    if ((e = make_tensor_map_f16(&t1, m1, 64, PW)) != cudaSuccess) return e;
    if ((e = make_tensor_map_f16(&t2, m2, 64, PW)) != cudaSuccess) return e;
    if ((e = make_tensor_map_f16(&t3, m3, 64, PW)) != cudaSuccess) return e;
    const int tiles = (m + PW - 1) / PW;
    const int grid = tiles < num_sms ? tiles : num_sms;
    e = launch_pdl(apply128_tc_kernel, dim3(grid), dim3(TCA_THREADS), (size_t)TCA_SMEM, stream, t1, t2, t3, A,
                   lda, m, (const TcApplyFactors*)fac, Qh, ldqh);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace lb
