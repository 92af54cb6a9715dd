// Split-K Gram product R12 = Q1^T A2 with the fp32 -> fp16 cast of A2 FUSED INTO THE LOAD.
//
// On the left spine of the recursion A2 is still the caller's fp32 input: no kernel has produced an
// fp16 shadow of it yet.  The generic path casts those columns first (cast_shadow_kernel: read 4,
// write 2 bytes per element) and the Gram kernel then reads the shadow (2 more).  For tall matrices,
// where this product is bandwidth-bound, this kernel reads the fp32 columns once instead (4 bytes
// per element): eight converter warps load A2 straight from global memory, round to fp16 (RN, the
// same rounding the cast kernel applies, so the product is bit-identical) and write the B tile into
// shared memory in the K-major SWIZZLE_128B layout TMA would have produced; the A operand (Q1, an
// fp16 shadow the panel kernel wrote) still arrives by TMA.  Everything downstream is the pipeline
// of tc_gemm.cu: tcgen05.mma into TMEM accumulators, split-K partials, fixed-order reduce.  The
// pre-update shadow of A2 is never needed by anyone else (the update kernel rewrites it), so the
// cast kernel disappears for these nodes.
#include "tc_gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>

namespace lb {

using namespace ptx;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int CONV_THREADS = 256;                       // warps 6 .. 13
constexpr int GC_THREADS = 192 + CONV_THREADS;

// One CTA owns ALL MT = Mc / 128 row tiles of the output for one 128-column tile and one K split, so
// that the expensive operand - the fp32 A2 tile it has to read and convert - is fetched once per
// k block and multiplied against every Q1 tile while it is in shared memory: MT accumulators of 128
// TMEM columns (MT <= 4 = all 512 columns).  Against one 128 x 256 tile per CTA this cuts the
// L2 -> SM traffic of the h = 512 node from 10.5 to 6.4 GB (1048576 rows).
constexpr int BN = 128;
template <int MT>
struct GcCfg {
    static constexpr int STAGES = MT == 4 ? 2 : (MT == 2 ? 4 : 6);
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = B_TILE_BYTES + MT * A_TILE_BYTES;
    static constexpr int TMEM_COLS = MT == 1 ? 128 : (MT == 2 ? 256 : 512);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

struct GcParams {
    int M, N;                 // output extent: Q1 columns x A2 columns
    int k_rows;               // matrix rows (K)
    int kb_total, kb_per_split, splits;
    int tiles_n;
    int a_c1;                 // first Q1 column in the shadow
    const float* B;           // A2: fp32, column-major, first column of the block
    long ldb;
    float* part;              // [splits][N][M]
};

template <int MT>
__global__ void __launch_bounds__(GC_THREADS, 1)
tc_gram_cast_kernel(const __grid_constant__ CUtensorMap mapA, const GcParams p) {
    using C = GcCfg<MT>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tempty_bar = bar_base + 8u * (2 * STAGES + 1);
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) prefetch_tensormap(&mapA);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), 1 + CONV_THREADS);   // the TMA thread + every converter thread
                mbar_init(empty_bar(s), 1);
            }
            mbar_init(tfull_bar, 1);
            mbar_init(tempty_bar, 4);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();

    const int items = p.tiles_n * p.splits;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (A = Q1)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int split = item / p.tiles_n;
                const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_arrive_expect_tx(full_bar(stage), MT * A_TILE_BYTES);
                    const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES + C::B_TILE_BYTES;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        tma_load_2d(a_dst + mt * A_TILE_BYTES, &mapA, full_bar(stage), kb * BK, p.a_c1 + mt * BM);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(0, 0u, 0u, BM, BN);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int split = item / p.tiles_n;
                const int kb0 = split * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                mbar_wait(tempty_bar, acc_phase ^ 1u);      // the previous item has been drained
                tc_fence_after_sync();
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after_sync();
                    const uint32_t b_src = smem_base + stage * C::STAGE_BYTES;
                    const uint64_t b_desc = make_smem_desc_sw128(b_src, 16, 1024);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t a_desc =
                            make_smem_desc_sw128(b_src + C::B_TILE_BYTES + mt * A_TILE_BYTES, 16, 1024);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_f16(tmem_base + mt * BN, a_desc + k * 2, b_desc + k * 2, idesc,
                                     (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar);
                acc_phase ^= 1u;
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ epilogue: split-K partials
        const int quad = warp & 3;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int split = item / p.tiles_n, n_blk = item - split * p.tiles_n;
            mbar_wait(tfull_bar, acc_phase);
            tc_fence_after_sync();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int row = mt * BM + quad * 32 + lane;
                const bool row_ok = row < p.M;
                const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + mt * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int col0 = n_blk * BN + c * 32;
                    uint32_t d[32];
                    tmem_ld_32x32(t_addr + c * 32, d);
                    tmem_ld_wait();
                    float* pp = p.part + (long)split * p.M * p.N + row + (long)col0 * p.M;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (row_ok && col0 + j < p.N) pp[(long)j * p.M] = __uint_as_float(d[j]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            acc_phase ^= 1u;
        }
    } else {
        // ------------------------------------------------------------------ converters (B = fp16(A2))
        // The global loads of NPF k blocks are kept in flight in registers, ahead of the shared-memory
        // ring: with one block in flight a converter sees the full DRAM latency per k block (measured:
        // 3 us per k block, 2.4 TB/s for the whole GPU).
        const int ct = threadIdx.x - 192;
        const int q = ct & 7;                       // 8-row chunk of the 64-row k block
        const int nb = ct >> 3;                     // columns nb + 32 u
        constexpr int CH = BN / 32;                 // chunks per thread and k block
        constexpr int NPF = 3;
        // cursor over this CTA's (item, k block) sequence
        struct Cursor { int item, kb, kb1, n_blk; bool valid; };
        auto start_item = [&](Cursor& c) {
            c.valid = c.item < items;
            if (!c.valid) return;
            const int split = c.item / p.tiles_n;
            c.n_blk = c.item - split * p.tiles_n;
            c.kb = split * p.kb_per_split;
            c.kb1 = min(p.kb_total, c.kb + p.kb_per_split);
        };
        auto advance = [&](Cursor& c) {
            if (++c.kb >= c.kb1) { c.item += gridDim.x; start_item(c); }
        };
        float4 v[NPF][CH][2];
        auto issue = [&](const Cursor& c, float4 (&buf)[CH][2]) {
            const int row = c.kb * BK + q * 8;
            const bool row_ok = row < p.k_rows;             // k_rows % 8 == 0: a chunk is in or out
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int col = c.n_blk * BN + nb + 32 * u;
                if (row_ok && col < p.N) {
                    const float4* src = reinterpret_cast<const float4*>(p.B + row + (long)col * p.ldb);
                    buf[u][0] = src[0];
                    buf[u][1] = src[1];
                } else {
                    buf[u][0] = buf[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        Cursor ld{(int)blockIdx.x, 0, 0, 0, false}, st{(int)blockIdx.x, 0, 0, 0, false};
        start_item(ld);
        start_item(st);
#pragma unroll
        for (int s = 0; s < NPF; ++s)
            if (ld.valid) { issue(ld, v[s]); advance(ld); }
        int stage = 0;
        uint32_t phase = 0;
        while (st.valid) {
#pragma unroll
            for (int s = 0; s < NPF; ++s) {
                if (!st.valid) break;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                uint8_t* b_dst = smem_gen + stage * C::STAGE_BYTES;
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const int n = nb + 32 * u;
                    const __half2 h0 = __floats2half2_rn(v[s][u][0].x, v[s][u][0].y);
                    const __half2 h1 = __floats2half2_rn(v[s][u][0].z, v[s][u][0].w);
                    const __half2 h2 = __floats2half2_rn(v[s][u][1].x, v[s][u][1].y);
                    const __half2 h3 = __floats2half2_rn(v[s][u][1].z, v[s][u][1].w);
                    uint4 o;
                    o.x = *reinterpret_cast<const uint32_t*>(&h0);
                    o.y = *reinterpret_cast<const uint32_t*>(&h1);
                    o.z = *reinterpret_cast<const uint32_t*>(&h2);
                    o.w = *reinterpret_cast<const uint32_t*>(&h3);
                    // K-major SWIZZLE_128B: 128-byte row per column n, 16-byte chunk q ^ (n & 7)
                    *reinterpret_cast<uint4*>(b_dst + n * 128 + ((q ^ (n & 7)) << 4)) = o;
                }
                fence_proxy_async_smem();
                mbar_arrive(full_bar(stage));
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                advance(st);
                if (ld.valid) { issue(ld, v[s]); advance(ld); }     // refill the slot just consumed
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <int MT>
cudaError_t launch_gc(cudaStream_t stream, int num_sms, const CUtensorMap& mapA, const GcParams& p) {
    const int items = p.tiles_n * p.splits;
    const int grid = std::max(1, std::min(items, num_sms));
    cudaError_t e = launch_pdl(tc_gram_cast_kernel<MT>, dim3(grid), dim3(GC_THREADS),
                               (size_t)GcCfg<MT>::SMEM_BYTES, stream, mapA, p);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

cudaError_t tc_gram_cast_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(tc_gram_cast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  GcCfg<1>::SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(tc_gram_cast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  GcCfg<2>::SMEM_BYTES)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_gram_cast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GcCfg<4>::SMEM_BYTES);
}

bool tc_gram_cast_supports(int Mc) { return Mc == 128 || Mc == 256 || Mc == 512; }

int tc_gram_cast_splits(int num_sms, int Mc, int k_rows) {
    const int strips = (Mc + BN - 1) / BN;                 // of the whole node: Nc = Mc
    const int kb_total = (k_rows + BK - 1) / BK;
    int s = std::max(2, std::min(num_sms / strips, kb_total / 8));      // one wave, >= 8 k blocks per split
    const int kb_per = (kb_total + s - 1) / s;
    return std::max(1, (kb_total + kb_per - 1) / kb_per);  // no empty split
}

cudaError_t tc_gram_cast(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_128, int k_rows,
                         int colA, int Mc, const float* B, long ldb, int Nc, float* C, long ldc, __half* Ch,
                         long ldch, float* part, int splits, float* Z) {
    if (!tc_gram_cast_supports(Mc) || splits < 2 || !part || k_rows % 8 != 0 || ldb % 4 != 0 ||
        (reinterpret_cast<uintptr_t>(B) & 15) != 0)
        return cudaErrorInvalidValue;
    GcParams p{};
    p.M = Mc; p.N = Nc; p.k_rows = k_rows;
    p.kb_total = (k_rows + BK - 1) / BK;
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;      // no empty split (as tc_gram)
    p.tiles_n = (Nc + BN - 1) / BN;
    p.a_c1 = colA;
    p.B = B; p.ldb = ldb; p.part = part;
    cudaError_t e = Mc == 512 ? launch_gc<4>(stream, num_sms, mapQ_128, p)
                  : Mc == 256 ? launch_gc<2>(stream, num_sms, mapQ_128, p)
                              : launch_gc<1>(stream, num_sms, mapQ_128, p);
    if (e != cudaSuccess) return e;
    return splitk_reduce(stream, part, p.splits, Mc, Nc, C, ldc, Ch, ldch, Z);
}

}  // namespace lb
