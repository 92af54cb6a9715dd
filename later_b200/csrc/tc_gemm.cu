// tcgen05 GEMM kernels for RGSQRF's trailing updates (see tc_gemm.cuh for the contract).
//
// One persistent, warp-specialised kernel template:
//   warp 0      TMA producer: fills a STAGES-deep ring of {A tile, B tile} in shared memory
//               (cp.async.bulk.tensor, SWIZZLE_128B), signalling "full" mbarriers by byte count.
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N=BN, K=16, fp16 -> fp32) from the
//               shared-memory descriptors into one of two TMEM accumulator stages; tcgen05.commit
//               releases ring slots ("empty") and publishes finished accumulators ("tmem full").
//   warps 2..5  epilogue: tcgen05.ld the accumulator (lane = output row, so a warp touches 32
//               consecutive fp32 of one column of the column-major C: one 128-byte line per
//               request), apply C = D | C -= D, emit the fp32 result and its fp16 shadow.
// The accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps
// the MMAs of tile i+1.
#include "tc_gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace lb {

using namespace ptx;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // halfs per k block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB

template <int BN>
struct Cfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + align slack
};

// Grouped rasterisation: consecutive work items sweep GN n-tiles, then the next m-tile, so the
// ~148 tiles in flight share a small set of operand panels in L2.
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& m_blk,
                                            int& n_blk) {
    constexpr int GN = 8;
    const int per_group = GN * tiles_m;
    const int g = t / per_group;
    const int r = t - g * per_group;
    const int gn = min(GN, tiles_n - g * GN);
    m_blk = r / gn;
    n_blk = g * GN + (r - m_blk * gn);
}

template <int BN, bool A_MN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const TcGemmParams p) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    // generic pointer to the tmem slot (same offset from smem_raw as the shared address)
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mapA);
        prefetch_tensormap(&mapB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full_bar(s), 1);
                mbar_init(empty_bar(s), 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(tfull_bar(a), 1);
                mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
            }
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
    pdl_trigger();   // the next kernel may start its prologue
    pdl_wait();      // operands are written by the predecessors

    const int tiles = p.tiles_m * p.tiles_n;
    const int items = tiles * p.splits;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int split = item / tiles;
                int m_blk, n_blk;
                tile_coords(item - split * tiles, p.tiles_m, p.tiles_n, m_blk, n_blk);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_TILE_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    if (A_MN) {
                        // two 64-row chunks of the M (contiguous) dimension, 64 k-columns each
                        tma_load_2d(a_dst, &mapA, full_bar(stage), p.a_c0 + m_blk * BM,
                                    p.a_c1 + kb * BK);
                        tma_load_2d(a_dst + 64 * BK * 2, &mapA, full_bar(stage),
                                    p.a_c0 + m_blk * BM + 64, p.a_c1 + kb * BK);
                    } else {
                        tma_load_2d(a_dst, &mapA, full_bar(stage), p.a_c0 + kb * BK,
                                    p.a_c1 + m_blk * BM);
                    }
                    tma_load_2d(b_dst, &mapB, full_bar(stage), p.b_c0 + kb * BK,
                                p.b_c1 + n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(/*F16*/ 0, A_MN ? 1u : 0u, 0u, BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int split = item / tiles;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after_sync();
                    const uint32_t a_src = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_src = a_src + A_TILE_BYTES;
                    const uint64_t a_desc = A_MN ? make_smem_desc_sw128(a_src, 64 * BK * 2, 1024)
                                                 : make_smem_desc_sw128(a_src, 16, 1024);
                    const uint64_t b_desc = make_smem_desc_sw128(b_src, 16, 1024);
                    // per UMMA_K step: K-major +32 B inside the swizzle row; MN-major +16 k-rows
                    constexpr uint64_t a_step = A_MN ? (UMMA_K * 128 / 16) : (UMMA_K * 2 / 16);
                    constexpr uint64_t b_step = UMMA_K * 2 / 16;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_f16(d_tmem, a_desc + k * a_step, b_desc + k * b_step, idesc,
                                 (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int split = item / tiles;
            int m_blk, n_blk;
            tile_coords(item - split * tiles, p.tiles_m, p.tiles_n, m_blk, n_blk);
            const int row = m_blk * BM + quad * 32 + lane;
            const bool row_ok = row < p.M;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after_sync();
            const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;
            const float dsc = p.dscale ? *p.dscale : 1.f;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n_blk * BN + c * 32;
                uint32_t d[32];
                tmem_ld_32x32(t_addr + c * 32, d);
                if (EPI == EPI_SUB || EPI == EPI_ADD) {
                    float cv[32];
                    float* cp = p.C + row + (long)col0 * p.ldc;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        cv[j] = (row_ok && col0 + j < p.N) ? cp[(long)j * p.ldc] : 0.f;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (row_ok && col0 + j < p.N) {
                            const float dv = __uint_as_float(d[j]) * dsc;
                            const float v = (EPI == EPI_SUB) ? cv[j] - dv : cv[j] + dv;
                            cp[(long)j * p.ldc] = v;
                            if (p.Ch) p.Ch[row + (long)(col0 + j) * p.ldch] = __float2half_rn(v);
                        }
                    }
                } else if (EPI == EPI_STORE) {
                    tmem_ld_wait();
                    float* cp = p.C + row + (long)col0 * p.ldc;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (row_ok && col0 + j < p.N) {
                            const float v = __uint_as_float(d[j]) * dsc;
                            cp[(long)j * p.ldc] = v;
                            if (p.Ch) p.Ch[row + (long)(col0 + j) * p.ldch] = __float2half_rn(v);
                            if (p.Z) p.Z[row + (long)(col0 + j) * p.ldc] = 0.f;
                        }
                    }
                } else {  // EPI_PARTIAL
                    tmem_ld_wait();
                    float* pp = p.part + (long)split * p.M * p.N + row + (long)col0 * p.M;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (row_ok && col0 + j < p.N) pp[(long)j * p.M] = __uint_as_float(d[j]);
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}


// ---------------------------------------------------------------------------------------------
// CTA-pair variant of the Gram product for the large (compute-bound) levels: a cluster of two
// CTAs computes a 256 x 256 tile with tcgen05.mma.cta_group::2 (M = 256).  Each CTA stages its own
// 128 rows of A and HALF of the B tile, so per SM the shared-memory fill traffic drops from 48 to
// 32 KiB per k-block (and the ring gets 6 stages instead of 4); the accumulator rows 0-127 live in
// the leader's TMEM, rows 128-255 in the peer's.  Protocol:
//   full[s]   leader's barrier; the leader arms it for the bytes of BOTH CTAs, the peer's TMA
//             (cp.async.bulk.tensor ... cta_group::2) credits it through its shared::cluster address
//   empty[s]  one per CTA; tcgen05.commit.cta_group::2 ... multicast arrives on both
//   tfull[a]  one per CTA (multicast commit); tempty[a] leader's, 4 local + 4 remote arrivals
template <int BN>
struct Cfg2 {
    static constexpr int STAGES = 6;
    static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_HALF_BYTES;      // per CTA
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gram2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const TcGemmParams p) {
    using C = Cfg2<BN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mapA);
        prefetch_tensormap(&mapB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tc_fence_before_sync();
    cluster_sync();                                    // barriers of both CTAs are initialised
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_trigger();
    pdl_wait();

    const int pairs_m = p.tiles_m >> 1;
    const int items = pairs_m * p.tiles_n;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = cluster_id; item < items; item += num_clusters) {
                int pm, n_blk;
                tile_coords(item, pairs_m, p.tiles_n, pm, n_blk);
                const int m_blk = 2 * pm + (int)rank;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_TILE_BYTES;
                    const uint32_t lead_full = mapa(full_bar(stage), 0);
                    if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
                    tma_load_2d_2cta(a_dst, &mapA, lead_full, p.a_c0 + kb * BK, p.a_c1 + m_blk * BM);
                    tma_load_2d_2cta(b_dst, &mapB, lead_full, p.b_c0 + kb * BK,
                                     p.b_c1 + n_blk * BN + (int)rank * (BN / 2));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader only)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc(0, 0u, 0u, 2 * BM, BN);   // M = 256
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int item = cluster_id; item < items; item += num_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after_sync();
                    const uint32_t a_src = smem_base + stage * C::STAGE_BYTES;
                    const uint64_t a_desc = make_smem_desc_sw128(a_src, 16, 1024);
                    const uint64_t b_desc = make_smem_desc_sw128(a_src + A_TILE_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_f16_2cta(d_tmem, a_desc + k * (UMMA_K * 2 / 16), b_desc + k * (UMMA_K * 2 / 16),
                                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2cta(empty_bar(stage), 3);   // frees the slot in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit_2cta(tfull_bar(acc), 3);         // accumulator complete, both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (both CTAs)
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = cluster_id; item < items; item += num_clusters) {
            int pm, n_blk;
            tile_coords(item, pairs_m, p.tiles_n, pm, n_blk);
            const int row = (2 * pm + (int)rank) * BM + quad * 32 + lane;
            const bool row_ok = row < p.M;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after_sync();
            const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n_blk * BN + c * 32;
                uint32_t d[32];
                tmem_ld_32x32(t_addr + c * 32, d);
                tmem_ld_wait();
                float* cp = p.C + row + (long)col0 * p.ldc;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (row_ok && col0 + j < p.N) {
                        const float v = __uint_as_float(d[j]);
                        cp[(long)j * p.ldc] = v;
                        if (p.Ch) p.Ch[row + (long)(col0 + j) * p.ldch] = __float2half_rn(v);
                        if (p.Z) p.Z[row + (long)(col0 + j) * p.ldc] = 0.f;
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(acc), 0));   // leader's barrier
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before_sync();
    cluster_sync();          // nobody tears down while the partner may still touch its smem / TMEM
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
    }
}

template <int BN>
cudaError_t launch_gram2(cudaStream_t stream, int num_sms, const CUtensorMap& mapA,
                         const CUtensorMap& mapB, const TcGemmParams& p) {
    const int items = (p.tiles_m / 2) * p.tiles_n;
    const int clusters = std::max(1, std::min(items, num_sms / 2));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg2<BN>::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gram2_kernel<BN>, mapA, mapB, p);
    return e != cudaSuccess ? e : cudaGetLastError();
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N,
                                     float* __restrict__ C, long ldc, __half* __restrict__ Ch,
                                     long ldch, float* __restrict__ Z) {
    const long total = (long)M * N;
    pdl_trigger();
    pdl_wait();
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[(long)k * total + idx];  // fixed order
        const int i = (int)(idx % M);
        const int j = (int)(idx / M);
        C[i + (long)j * ldc] = s;
        if (Ch) Ch[i + (long)j * ldch] = __float2half_rn(s);
        if (Z) Z[i + (long)j * ldc] = 0.f;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        }
    });
    return fn;
}

template <int BN, bool A_MN, int EPI>
cudaError_t set_smem_attr() {
    return cudaFuncSetAttribute(tc_gemm_kernel<BN, A_MN, EPI>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES);
}

template <int BN, bool A_MN, int EPI>
cudaError_t launch(cudaStream_t stream, int num_sms, const CUtensorMap& mapA,
                   const CUtensorMap& mapB, const TcGemmParams& p) {
    const int items = p.tiles_m * p.tiles_n * p.splits;
    const int grid = std::max(1, std::min(items, num_sms));
    cudaError_t e = launch_pdl(tc_gemm_kernel<BN, A_MN, EPI>, dim3(grid), dim3(kThreads),
                               (size_t)Cfg<BN>::SMEM_BYTES, stream, mapA, mapB, p);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

cudaError_t tc_gemm_init() {
    cudaError_t e;
#define LB_SET(BN, AMN, EPI) \
    if ((e = set_smem_attr<BN, AMN, EPI>()) != cudaSuccess) return e;
    LB_SET(128, false, EPI_STORE) LB_SET(256, false, EPI_STORE)
    LB_SET(128, false, EPI_PARTIAL) LB_SET(256, false, EPI_PARTIAL)
    LB_SET(128, true, EPI_SUB) LB_SET(256, true, EPI_SUB)
    LB_SET(128, true, EPI_STORE) LB_SET(256, true, EPI_STORE)
    LB_SET(128, false, EPI_ADD) LB_SET(256, false, EPI_ADD)
    LB_SET(128, true, EPI_ADD) LB_SET(256, true, EPI_ADD)
#undef LB_SET
    if ((e = cudaFuncSetAttribute(tc_gram2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg2<256>::SMEM_BYTES)) != cudaSuccess)
        return e;
    return cudaSuccess;
}

cudaError_t make_tensor_map_f16(CUtensorMap* out, const HalfMatrix& mat, int box_inner,
                                int box_outer) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return cudaErrorNotSupported;
    if (mat.ld % 8 != 0 || (reinterpret_cast<uintptr_t>(mat.ptr) & 15) != 0)
        return cudaErrorInvalidValue;
    cuuint64_t dims[2] = {(cuuint64_t)mat.rows, (cuuint64_t)mat.cols};
    cuuint64_t strides[1] = {(cuuint64_t)mat.ld * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(mat.ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t tc_gemm_launch(cudaStream_t stream, int num_sms, bool a_mn_major, int bn, int epi,
                           const CUtensorMap& mapA, const CUtensorMap& mapB, const TcGemmParams& p) {
#define LB_CASE(BN, AMN, EPI) \
    if (bn == BN && a_mn_major == AMN && epi == EPI) \
        return launch<BN, AMN, EPI>(stream, num_sms, mapA, mapB, p);
    LB_CASE(128, false, EPI_STORE) LB_CASE(256, false, EPI_STORE)
    LB_CASE(128, false, EPI_PARTIAL) LB_CASE(256, false, EPI_PARTIAL)
    LB_CASE(128, false, EPI_ADD) LB_CASE(256, false, EPI_ADD)
    LB_CASE(128, true, EPI_SUB) LB_CASE(256, true, EPI_SUB)
    LB_CASE(128, true, EPI_STORE) LB_CASE(256, true, EPI_STORE)
    LB_CASE(128, true, EPI_ADD) LB_CASE(256, true, EPI_ADD)
#undef LB_CASE
    return cudaErrorInvalidValue;
}

void tc_fill_gram(TcGemmParams& p, int bn, int row0, int k_rows, int colA, int Mc, int colB, int Nc,
                  float* C, long ldc, __half* Ch, long ldch) {
    p = TcGemmParams{};
    p.M = Mc; p.N = Nc;
    p.kb_total = (k_rows + BK - 1) / BK;
    p.splits = 1; p.kb_per_split = p.kb_total;
    p.tiles_m = (Mc + BM - 1) / BM;
    p.tiles_n = (Nc + bn - 1) / bn;
    p.a_c0 = row0; p.a_c1 = colA; p.b_c0 = row0; p.b_c1 = colB;
    p.C = C; p.ldc = ldc; p.Ch = Ch; p.ldch = ldch;
}

void tc_fill_update(TcGemmParams& p, int bn, int row0, int Mr, int colA, int K, int colB0, int Nc,
                    float* C, long ldc, __half* Ch, long ldch) {
    p = TcGemmParams{};
    p.M = Mr; p.N = Nc;
    p.kb_total = (K + BK - 1) / BK;
    p.splits = 1; p.kb_per_split = p.kb_total;
    p.tiles_m = (Mr + BM - 1) / BM;
    p.tiles_n = (Nc + bn - 1) / bn;
    p.a_c0 = row0; p.a_c1 = colA; p.b_c0 = 0; p.b_c1 = colB0;
    p.C = C; p.ldc = ldc; p.Ch = Ch; p.ldch = ldch;
}

// Split-K factor for the Gram-type product (tiny output, K = number of matrix rows).  Cost model in
// units of one BN=256 k-block (~512 MMA cycles): every wave pays its k-blocks plus a fixed
// prologue/epilogue overhead, and every extra split pays for writing and re-reading one partial.
int choose_gram_splits(int num_sms, int Mc, int Nc, int bn, int k_rows) {
    const int tiles = ((Mc + BM - 1) / BM) * ((Nc + bn - 1) / bn);
    const int kb_total = (k_rows + BK - 1) / BK;
    const double unit = bn == 256 ? 1.0 : 0.5;
    const double overhead = 8.0;
    const double per_split = (double)Mc * Nc * 8.0 / 3.0e12 / 0.27e-6;   // partial write + read
    int best = 1;
    double best_cost = 1e300;
    const int smax = std::min(kb_total, 2 * num_sms);
    for (int s = 1; s <= smax; ++s) {
        const int kb_per = (kb_total + s - 1) / s;
        if ((kb_total + kb_per - 1) / kb_per != s) continue;            // would leave an empty split
        const long items = (long)tiles * s;
        const long waves = (items + num_sms - 1) / num_sms;
        const double cost = waves * (kb_per * unit + overhead) + (s > 1 ? s * per_split : 0.0);
        if (cost < best_cost) { best_cost = cost; best = s; }
    }
    return best;
}

cudaError_t tc_gram(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_128,
                    const CUtensorMap& mapQ_bn, int bn, int row0, int k_rows, int colA, int Mc,
                    int colB, int Nc, float* C, long ldc, __half* Ch, long ldch, float* part,
                    int splits, float* Z) {
    TcGemmParams p{};
    p.M = Mc;
    p.N = Nc;
    p.kb_total = (k_rows + BK - 1) / BK;
    p.splits = std::max(1, splits);
    p.kb_per_split = (p.kb_total + p.splits - 1) / p.splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
    p.tiles_m = (Mc + BM - 1) / BM;
    p.tiles_n = (Nc + bn - 1) / bn;
    p.a_c0 = row0; p.a_c1 = colA;
    p.b_c0 = row0; p.b_c1 = colB;
    p.C = C; p.ldc = ldc; p.Ch = Ch; p.ldch = ldch; p.part = part;
    p.Z = p.splits == 1 ? Z : nullptr;      // with split-K the reduce kernel does the zeroing
    cudaError_t e;
    static const bool use_pair = [] { const char* v = getenv("LB_GRAM_2CTA"); return !v || atoi(v) != 0; }();
    const int ntiles = p.tiles_m * p.tiles_n;
    if (p.splits == 1 && bn == 256 && use_pair && p.tiles_m % 2 == 0 && ntiles >= 2 * num_sms &&
        ntiles <= 1024) {
        // mid-size level (measured on B200, m = 16384: h = 4096 412 us vs 438 us single-CTA; at
        // h = 8192 both variants sit at the power-capped tensor peak, 1.36-1.38 PFLOP/s):
        // CTA-pair kernel; its B map is the 128-row box (half of the 256-wide tile)
        return launch_gram2<256>(stream, num_sms, mapQ_128, mapQ_128, p);
    }
    if (p.splits == 1) {
        e = (bn == 256) ? launch<256, false, EPI_STORE>(stream, num_sms, mapQ_128, mapQ_bn, p)
                        : launch<128, false, EPI_STORE>(stream, num_sms, mapQ_128, mapQ_bn, p);
        return e;
    }
    if (!part) return cudaErrorInvalidValue;
    e = (bn == 256) ? launch<256, false, EPI_PARTIAL>(stream, num_sms, mapQ_128, mapQ_bn, p)
                    : launch<128, false, EPI_PARTIAL>(stream, num_sms, mapQ_128, mapQ_bn, p);
    if (e != cudaSuccess) return e;
    return splitk_reduce(stream, part, p.splits, Mc, Nc, C, ldc, Ch, ldch, Z);
}

cudaError_t tc_update(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_64,
                      const CUtensorMap& mapB_bn, int bn, int row0, int Mr, int colA, int K,
                      int colB0, int Nc, float* C, long ldc, __half* Ch, long ldch, bool sub) {
    TcGemmParams p{};
    p.M = Mr;
    p.N = Nc;
    p.kb_total = (K + BK - 1) / BK;
    p.splits = 1;
    p.kb_per_split = p.kb_total;
    p.tiles_m = (Mr + BM - 1) / BM;
    p.tiles_n = (Nc + bn - 1) / bn;
    p.a_c0 = row0; p.a_c1 = colA;
    p.b_c0 = 0; p.b_c1 = colB0;
    p.C = C; p.ldc = ldc; p.Ch = Ch; p.ldch = ldch; p.part = nullptr;
    if (sub) {
        return (bn == 256) ? launch<256, true, EPI_SUB>(stream, num_sms, mapQ_64, mapB_bn, p)
                           : launch<128, true, EPI_SUB>(stream, num_sms, mapQ_64, mapB_bn, p);
    }
    return (bn == 256) ? launch<256, true, EPI_STORE>(stream, num_sms, mapQ_64, mapB_bn, p)
                       : launch<128, true, EPI_STORE>(stream, num_sms, mapQ_64, mapB_bn, p);
}

cudaError_t splitk_reduce(cudaStream_t stream, const float* part, int splits, int M, int N, float* C,
                          long ldc, __half* Ch, long ldch, float* Z) {
    const long total = (long)M * N;
    const int threads = 256;
    const int blocks = (int)std::min<long>((total + threads - 1) / threads, 148L * 8);
    cudaError_t e = launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(threads), 0, stream, part, splits,
                               M, N, C, ldc, Ch, ldch, Z);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace lb
