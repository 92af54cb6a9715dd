// Trailing update  A2 -= Q1 * R12  (and the TSQR back-multiplication  Q = Qh * W) with the fp32
// C tile streamed through shared memory by TMA in both directions.
//
// Same MMA pipeline as tc_gemm.cu (TMA -> smem ring -> tcgen05.mma -> TMEM, double-buffered
// accumulator), but the epilogue never touches global memory with ld/st:
//   warp 2      C producer: cp.async.bulk.tensor loads of 128 x CCH fp32 chunks of C into a ring of
//               CSLOTS slots, running ahead of the epilogue (and across tiles), so several chunks
//               per SM are always in flight - this is what the HBM-bound levels (K = w/2 <= 1024)
//               need, a per-thread ld.global epilogue tops out near 2 TB/s.
//   warps 4..7  epilogue: tcgen05.ld the accumulator chunk, C - D in place in the smem slot, the
//               fp16 shadow next to it, fence.proxy.async, then one thread issues the two TMA stores
//               (fp32 C and fp16 shadow) and recycles the slot once the stores have read it.
#include "tc_gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>

namespace lb {

using namespace ptx;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;
constexpr int A_TILE_BYTES = BM * BK * 2;

template <int BN, int CCH, int CSLOTS>
struct UCfg {
    static constexpr int STAGES = 4;
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int C32_BYTES = BM * CCH * 4;
    static constexpr int C16_BYTES = BM * CCH * 2;
    static constexpr int CSLOT_BYTES = C32_BYTES + C16_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 1024;     // barriers (< 256 B) + 128 column maxima
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + CSLOTS * CSLOT_BYTES + BAR_BYTES + 1024;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct UpdParams {
    int M, N;
    int kb_total;
    int tiles_m, tiles_n;
    int a_c0, a_c1;   // A operand origin (row, col) in the fp16 shadow
    int b_c1;         // first column of B (R12h / Wh) to use
    int c_r0, c_c0;   // C origin (row, col) in the fp32 matrix (and in the fp16 shadow)
    // Optional: largest |value| written to each of the first 128 columns of the C block, as one
    // partial per CTA, colmax_part[col * colmax_parts + blockIdx.x] (slots >= gridDim.x zeroed).
    // These columns are the next panel; its integer Gram kernel scales by them (panel_tc.cu).
    float* colmax_part;
    int colmax_parts;
    // The fp16 shadow of the first shadow_from columns of the block is not written: those columns are the
    // next panel, whose apply kernel rewrites their shadow before anybody reads it (a multiple of CCH).
    int shadow_from;
};

__device__ __forceinline__ void tile_coords_u(int t, int tiles_m, int tiles_n, int& m_blk,
                                              int& n_blk) {
    constexpr int GN = 8;
    const int per_group = GN * tiles_m;
    const int g = t / per_group;
    const int r = t - g * per_group;
    const int gn = min(GN, tiles_n - g * GN);
    m_blk = r / gn;
    n_blk = g * GN + (r - m_blk * gn);
}

template <int CCH>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&d)[CCH]) {
    if constexpr (CCH == 32) tmem_ld_32x32(taddr, d);
    else tmem_ld_32x16(taddr, d);
}

template <int BN, int CCH, int CSLOTS, bool SUB, bool SHADOW>
__global__ void __launch_bounds__(kThreads, 1)
tc_update_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapH,
                 const UpdParams p) {
    using C = UCfg<BN, CCH, CSLOTS>;
    constexpr int STAGES = C::STAGES;
    constexpr int NCHUNK = BN / CCH;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t cring_base = smem_base + STAGES * C::STAGE_BYTES;
    const uint32_t bar_base = cring_base + CSLOTS * C::CSLOT_BYTES;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));   // generic view of smem_base
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    auto cfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
    auto cempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + CSLOTS + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4 + 2 * CSLOTS);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mapA);
        prefetch_tensormap(&mapB);
        prefetch_tensormap(&mapC);
        if (SHADOW) prefetch_tensormap(&mapH);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
            for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
            for (int s = 0; s < CSLOTS; ++s) { mbar_init(cfull_bar(s), 1); mbar_init(cempty_bar(s), 1); }
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
    const int tiles = p.tiles_m * p.tiles_n;
    pdl_trigger();   // the next kernel may start its prologue
    pdl_wait();      // operands and C are written by the predecessors

    if (warp == 0) {
        // ------------------------------------------------------------------ A/B producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int m_blk, n_blk;
                tile_coords_u(t, p.tiles_m, p.tiles_n, m_blk, n_blk);
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_TILE_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    tma_load_2d(a_dst, &mapA, full_bar(stage), p.a_c0 + m_blk * BM, p.a_c1 + kb * BK);
                    tma_load_2d(a_dst + 64 * BK * 2, &mapA, full_bar(stage), p.a_c0 + m_blk * BM + 64,
                                p.a_c1 + kb * BK);
                    tma_load_2d(b_dst, &mapB, full_bar(stage), kb * BK, p.b_c1 + n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(0, 1u, 0u, BM, BN);   // A MN-major, B K-major
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after_sync();
                    const uint32_t a_src = smem_base + stage * C::STAGE_BYTES;
                    const uint64_t a_desc = make_smem_desc_sw128(a_src, 64 * BK * 2, 1024);
                    const uint64_t b_desc = make_smem_desc_sw128(a_src + A_TILE_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_f16(d_tmem, a_desc + k * (UMMA_K * 128 / 16), b_desc + k * (UMMA_K * 2 / 16),
                                 idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ C producer
        if (SUB && lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int m_blk, n_blk;
                tile_coords_u(t, p.tiles_m, p.tiles_n, m_blk, n_blk);
                for (int c = 0; c < NCHUNK; ++c) {
                    mbar_wait(cempty_bar(slot), phase ^ 1u);
                    mbar_arrive_expect_tx(cfull_bar(slot), C::C32_BYTES);
                    tma_load_2d(cring_base + slot * C::CSLOT_BYTES, &mapC, cfull_bar(slot),
                                p.c_r0 + m_blk * BM, p.c_c0 + n_blk * BN + c * CCH);
                    if (++slot == CSLOTS) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int quad = warp & 3;
        const int r = quad * 32 + lane;          // row inside the tile = TMEM lane
        const bool leader = (warp == 4 && lane == 0);
        unsigned* cmax = reinterpret_cast<unsigned*>(smem_gen + (bar_base - smem_base) + 256);
        if (p.colmax_part) {
            cmax[r] = 0u;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        int acc = 0;
        uint32_t acc_phase = 0;
        int slot = 0;
        uint32_t cphase = 0;
        int pending_slot = -1;                   // slot whose stores were committed last
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            int m_blk, n_blk;
            tile_coords_u(t, p.tiles_m, p.tiles_n, m_blk, n_blk);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after_sync();
            const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                uint32_t d[CCH];
                tmem_ld_chunk<CCH>(t_addr + c * CCH, d);
                float* sc = reinterpret_cast<float*>(smem_gen + STAGES * C::STAGE_BYTES +
                                                     slot * C::CSLOT_BYTES);
                __half* sh = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(sc) + C::C32_BYTES);
                if (SUB) {
                    mbar_wait(cfull_bar(slot), cphase);          // C chunk landed
                } else {
                    // no C producer: the slot is free once the stores issued CSLOTS chunks ago
                    // have finished reading it
                    if (leader) tma_store_wait_read<CSLOTS - 1>();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                tmem_ld_wait();
                const bool track = p.colmax_part && n_blk == 0 && c * CCH < 128;   // warp-uniform
                const bool shadow = SHADOW && n_blk * BN + c * CCH >= p.shadow_from;
#pragma unroll
                for (int j = 0; j < CCH; ++j) {
                    const float dv = __uint_as_float(d[j]);
                    const float v = SUB ? sc[j * BM + r] - dv : dv;
                    sc[j * BM + r] = v;
                    if (shadow) sh[j * BM + r] = __float2half_rn(v);
                    if (track) d[j] = __float_as_uint(fabsf(v));   // (ordered like unsigned integers)
                }
                if (track) {
                    // Column maxima over this warp's 32 rows for all CCH columns at once: recursive
                    // halving, lane L ends up with column L (mod CCH) - CCH - 1 shuffles, not 5 CCH.
#pragma unroll
                    for (int half = CCH / 2; half >= 1; half >>= 1) {
                        const bool upper = (lane & half) != 0;
#pragma unroll
                        for (int i = 0; i < half; ++i) {
                            const uint32_t send = upper ? d[i] : d[i + half];
                            const uint32_t keep = upper ? d[i + half] : d[i];
                            d[i] = max(keep, __shfl_xor_sync(0xffffffffu, send, half));
                        }
                    }
                    if (CCH == 16) d[0] = max(d[0], __shfl_xor_sync(0xffffffffu, d[0], 16));
                    if (lane < CCH) atomicMax(&cmax[c * CCH + lane], d[0]);
                }
                fence_proxy_async_smem();
                asm volatile("bar.sync 1, 128;" ::: "memory");   // the 4 epilogue warps
                if (leader) {
                    const int row0 = p.c_r0 + m_blk * BM;
                    const int col0 = p.c_c0 + n_blk * BN + c * CCH;
                    tma_store_2d(&mapC, cring_base + slot * C::CSLOT_BYTES, row0, col0);
                    if (shadow)
                        tma_store_2d(&mapH, cring_base + slot * C::CSLOT_BYTES + C::C32_BYTES, row0, col0);
                    tma_store_commit();
                    if (SUB && pending_slot >= 0) {
                        tma_store_wait_read<1>();                 // previous chunk's stores read smem
                        mbar_arrive(cempty_bar(pending_slot));    // hand the slot back to the producer
                    }
                    pending_slot = slot;
                }
                if (++slot == CSLOTS) { slot = 0; cphase ^= 1u; }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (leader) {
            tma_store_wait<0>();
            // (remaining c_empty arrivals are irrelevant: the producer has finished)
        }
        if (p.colmax_part) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float* dst = p.colmax_part + (long)r * p.colmax_parts;
            dst[blockIdx.x] = __uint_as_float(cmax[r]);
            if (blockIdx.x == 0)
                for (int sidx = gridDim.x; sidx < p.colmax_parts; ++sidx) dst[sidx] = 0.f;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(sym);
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

// Un-swizzled {box_rows x box_cols} map over a column-major matrix of 2- or 4-byte elements.
cudaError_t make_plain_map(CUtensorMap* out, const void* ptr, int elem_bytes, long rows, long cols,
                           long ld, int box_rows, int box_cols) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return cudaErrorNotSupported;
    if ((ld * elem_bytes) % 16 != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
        return cudaErrorInvalidValue;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                     2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// A whole small node of the recursion (half-width H = 128 or 256) in ONE launch, for matrices of at most
// 128 x #SMs rows:
//     R12 = Q1^T A2 (H x H),   A2 -= Q1 R12        (Q1, A2: m x H)
// Separately this is a split-K Gram kernel, its reduce kernel and an update kernel - three launches with
// a few microseconds of work each (16384 rows, H = 128: 12 + 5 + 22 us under ncu for 3 us of memory time),
// 64 + 32 times per 16384 x 16384 factorisation.  Here CTA t owns rows 128 t .. 128 t + 127 from start to end:
//   1. TMA: its Q1 tile and the fp16 shadow of its A2 tile (K-major, 2 k blocks), and - running ahead - the
//      fp32 A2 tile (first 128 columns); tcgen05.mma -> partial Gram block in TMEM -> global partial t (fp32)
//   2. grid barrier; the CTAs share the fixed-order sum over the partials (t = 0, 1, ...: deterministic), CTA c
//      taking outputs [c per, (c + 1) per): R12 to R (fp32), its zero mirror block, and the fp16 operand
//   3. grid barrier; per 128 columns of A2: every CTA stages the fp16 R12 columns as the K-major B operand
//      and multiplies the Q1 tile STILL IN SHARED MEMORY by it (the K-major tile of the Gram product is, byte
//      for byte, the MN-major A operand of the update: one 128-byte row per matrix column); epilogue A2 - D in
//      the staged fp32 tile, TMA store.  The fp16 shadow of the new A2 is written from column 128 on only: the
//      first 128 columns are the next panel (whose apply kernel rewrites their shadow before anybody reads it).
// The grid barriers are counters in global memory (self-resetting: the last CTA to leave clears them); all
// CTAs are co-resident because the grid is at most one CTA per SM and nothing before it in the stream waits
// for it.  Spins are bounded (trap), like every other wait in this library.
constexpr int NODE_THREADS = 192;
template <int H>
struct NodeCfg {
    static constexpr int NQ = H / 128;                     // 128-column blocks of Q1 (M tiles of step 1) and of A2
    static constexpr int KTILE_BYTES = H * BK * 2;         // one k block of an m x H fp16 tile: [H columns][64 rows]
    static constexpr int C_BYTES = BM * 128 * 4;           // fp32 A2 tile, 128 columns at a time
    static constexpr int TMEM_COLS = H == 128 ? 256 : 512; // step 1: NQ accumulators of H columns; step 3: 128
    static constexpr int OUTPUTS = H * H;                  // floats per partial
    static constexpr int SMEM_BYTES = 4 * KTILE_BYTES + C_BYTES + 256 + 1024;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct NodeParams {
    int tiles;             // row tiles = gridDim.x
    int q_c1, b_c1;        // first column of Q1 / of A2 (shadow and matrix coordinates)
    float* part;           // [tiles][H n][H m]
    float* R12; long ldr;  // R12 block inside R
    float* Z;              // mirror block to clear (same ld) or null
    __half* R12h;          // fp16 R12, ld H
    int* sync;             // three zero-initialised counters
};

__device__ __forceinline__ void node_grid_barrier(int* counter, int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // this CTA's global writes before the arrival
        atomicAdd(counter, 1);
        unsigned long long spins = 0;
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spins > (unsigned long long)(LB_SPIN_LIMIT)) __trap();
        } while (v < target);
    }
    __syncthreads();
}

template <int H>
__global__ void __launch_bounds__(NODE_THREADS, 1)
tc_node_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapC,
               const __grid_constant__ CUtensorMap mapH, const NodeParams p) {
    using C = NodeCfg<H>;
    constexpr int NQ = C::NQ;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t sQ = smem_base;                         // Q1 tile: 2 k blocks x [H columns][64 rows] fp16
    const uint32_t sB = sQ + 2 * C::KTILE_BYTES;           // A2 shadow tile; later R12h columns, new shadow
    const uint32_t sC = sB + 2 * C::KTILE_BYTES;           // fp32 A2 tile: 4 chunks x [32 columns][128 rows]
    const uint32_t bar_base = sC + C::C_BYTES;
    const uint32_t bar_in = bar_base, bar_c = bar_base + 8, bar_acc1 = bar_base + 16, bar_acc2 = bar_base + 24;
    const uint32_t tmem_slot = bar_base + 32;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&mapQ);
        prefetch_tensormap(&mapC);
        if (NQ > 1) prefetch_tensormap(&mapH);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(bar_in, 1); mbar_init(bar_c, 1); mbar_init(bar_acc1, 1); mbar_init(bar_acc2, 1);
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

    // ---------------------------------------------------------------- 1. partial Gram block of this row tile
    if (warp == 0 && lane == 0) {
        mbar_arrive_expect_tx(bar_in, 4 * C::KTILE_BYTES);
        for (int kb = 0; kb < 2; ++kb)
            for (int blk = 0; blk < NQ; ++blk) {
                tma_load_2d(sQ + kb * C::KTILE_BYTES + blk * A_TILE_BYTES, &mapQ, bar_in, tile * BM + kb * BK,
                            p.q_c1 + blk * 128);
                tma_load_2d(sB + kb * C::KTILE_BYTES + blk * A_TILE_BYTES, &mapQ, bar_in, tile * BM + kb * BK,
                            p.b_c1 + blk * 128);
            }
        mbar_arrive_expect_tx(bar_c, C::C_BYTES);          // (needed in step 3 only)
        for (int c = 0; c < 4; ++c)
            tma_load_2d(sC + c * (BM * 32 * 4), &mapC, bar_c, tile * BM, p.b_c1 + 32 * c);
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc(0, 0u, 0u, BM, H);        // both operands K-major, N = H
        mbar_wait(bar_in, 0);
        tc_fence_after_sync();
        for (int mt = 0; mt < NQ; ++mt)
            for (int kb = 0; kb < 2; ++kb) {
                const uint64_t a_desc = make_smem_desc_sw128(sQ + kb * C::KTILE_BYTES + mt * A_TILE_BYTES, 16, 1024);
                const uint64_t b_desc = make_smem_desc_sw128(sB + kb * C::KTILE_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_f16(tmem_base + mt * H, a_desc + k * (UMMA_K * 2 / 16), b_desc + k * (UMMA_K * 2 / 16),
                             idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
        umma_commit(bar_acc1);
    } else if (warp >= 2) {
        const int quad = warp & 3;
        mbar_wait(bar_acc1, 0);
        tc_fence_after_sync();
#pragma unroll 1
        for (int mt = 0; mt < NQ; ++mt) {
            // Q1 column mt 128 + quad 32 + lane = TMEM lane of accumulator mt
            float* pp = p.part + (long)tile * C::OUTPUTS + mt * 128 + quad * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < H / 32; ++c) {
                uint32_t d[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + mt * H + c * 32, d);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) pp[(c * 32 + j) * H] = __uint_as_float(d[j]);
            }
        }
        tc_fence_before_sync();
    }
    node_grid_barrier(p.sync + 0, p.tiles);

    // ---------------------------------------------------------------- 2. fixed-order sum, this CTA's share
    {
        const int per = (((C::OUTPUTS + p.tiles - 1) / p.tiles) + 3) & ~3;
        const int lo = min(C::OUTPUTS, tile * per), hi = min(C::OUTPUTS, lo + per);
        if constexpr (H == 128) {
            for (int e = lo + (int)threadIdx.x; e < hi; e += NODE_THREADS) {
                const float* src = p.part + e;
                float acc = 0.f;
                int t = 0;
                for (; t + 32 <= p.tiles; t += 32) {
                    float v[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) v[u] = __ldcg(src + (long)(t + u) * C::OUTPUTS);
#pragma unroll
                    for (int u = 0; u < 32; ++u) acc += v[u];
                }
                for (; t < p.tiles; ++t) acc += __ldcg(src + (long)t * C::OUTPUTS);
                const int i = e % H, j = e / H;
                p.R12[i + (long)j * p.ldr] = acc;
                if (p.Z) p.Z[i + (long)j * p.ldr] = 0.f;
                p.R12h[e] = __float2half_rn(acc);
            }
        } else {
            // four consecutive outputs (same column of R12) per thread: 16-byte loads, 32 in flight
            for (int e = lo + 4 * (int)threadIdx.x; e < hi; e += 4 * NODE_THREADS) {
                const float4* src = reinterpret_cast<const float4*>(p.part + e);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int t = 0;
                for (; t + 32 <= p.tiles; t += 32) {
                    float4 v[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) v[u] = __ldcg(src + (long)(t + u) * (C::OUTPUTS / 4));
#pragma unroll
                    for (int u = 0; u < 32; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
                }
                for (; t < p.tiles; ++t) {
                    const float4 v = __ldcg(src + (long)t * (C::OUTPUTS / 4));
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                const int i = e % H, j = e / H;
                float* r = p.R12 + i + (long)j * p.ldr;
                r[0] = acc.x; r[1] = acc.y; r[2] = acc.z; r[3] = acc.w;
                if (p.Z) { float* z = p.Z + i + (long)j * p.ldr; z[0] = 0.f; z[1] = 0.f; z[2] = 0.f; z[3] = 0.f; }
                const __half2 h01 = __floats2half2_rn(acc.x, acc.y), h23 = __floats2half2_rn(acc.z, acc.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(p.R12h + e) = pk;
            }
        }
    }
    node_grid_barrier(p.sync + 1, p.tiles);

    // ---------------------------------------------------------------- 3. A2 tile -= Q1 tile * R12, 128 columns at a time
#pragma unroll 1
    for (int np = 0; np < NQ; ++np) {
        // fp16 R12(:, 128 np ..) -> K-major SWIZZLE_128B B operand: column n is one 128-byte row per k block,
        // 16-byte chunk kc at position kc ^ (n & 7).  (What lived here - the A2 shadow tile, the previous
        // pass's operand - has been consumed: barrier 1, the __syncthreads at the end of the previous pass.)
        for (int q = threadIdx.x; q < 128 * (H / 8); q += NODE_THREADS) {
            const int n = q / (H / 8), kc = q % (H / 8);
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p.R12h + (long)(np * 128 + n) * H) + kc);
            *reinterpret_cast<uint4*>(smem_gen + (sB - smem_base) + (kc >> 3) * A_TILE_BYTES + n * 128 +
                                      (((kc & 7) ^ (n & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        __syncthreads();
        const uint32_t acc_col = H == 128 ? 128u : uint32_t(np) * 128u;   // (step 1's accumulators are drained)
        if (warp == 1 && lane == 0) {
            constexpr uint32_t idesc = make_idesc(0, 1u, 0u, BM, 128);  // A MN-major (rows), B K-major
            tc_fence_after_sync();
            for (int kb = 0; kb < H / BK; ++kb) {          // K = Q1 columns 64 kb .. 64 kb + 63
                // rows 0..63 of those columns sit in the first k-block tile of step 1, rows 64..127 in the second
                const uint64_t a_desc = make_smem_desc_sw128(sQ + kb * (64 * 128), C::KTILE_BYTES, 1024);
                const uint64_t b_desc = make_smem_desc_sw128(sB + kb * A_TILE_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_f16(tmem_base + acc_col, a_desc + k * (UMMA_K * 128 / 16), b_desc + k * (UMMA_K * 2 / 16),
                             idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_acc2);
        } else if (warp >= 2) {
            const int quad = warp & 3;
            const int r = quad * 32 + lane;                // row inside the tile = TMEM lane
            const bool shadow = np > 0;                    // columns 128.. of A2 are somebody's B operand later
            mbar_wait(bar_c, np & 1);
            mbar_wait(bar_acc2, np & 1);                   // (also: the MMAs have finished reading sB)
            tc_fence_after_sync();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t d[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc_col + c * 32, d);
                tmem_ld_wait();
                float* sc = reinterpret_cast<float*>(smem_gen + (sC - smem_base)) + c * (BM * 32);
                __half* sh = reinterpret_cast<__half*>(smem_gen + (sB - smem_base)) + c * (BM * 32);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float v = sc[j * BM + r] - __uint_as_float(d[j]);
                    sc[j * BM + r] = v;
                    if (shadow) sh[j * BM + r] = __float2half_rn(v);
                }
            }
            tc_fence_before_sync();
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory"); // the 4 epilogue warps
            if (warp == 2 && lane == 0) {
                const int col0 = p.b_c1 + np * 128;
                for (int c = 0; c < 4; ++c) {
                    tma_store_2d(&mapC, sC + c * (BM * 32 * 4), tile * BM, col0 + 32 * c);
                    if (shadow) tma_store_2d(&mapH, sB + c * (BM * 32 * 2), tile * BM, col0 + 32 * c);
                }
                tma_store_commit();
                if (np + 1 < NQ) {
                    tma_store_wait_read<0>();              // the stores have read sC: fetch the next 128 columns
                    mbar_arrive_expect_tx(bar_c, C::C_BYTES);
                    for (int c = 0; c < 4; ++c)
                        tma_load_2d(sC + c * (BM * 32 * 4), &mapC, bar_c, tile * BM, col0 + 128 + 32 * c);
                } else {
                    tma_store_wait<0>();
                }
            }
        }
        __syncthreads();   // the pass's MMAs are complete (the epilogue waited for them): sB may be restaged
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (threadIdx.x == 0) {
        // everybody who increments this counter has passed both barriers: the last one clears them
        if (atomicAdd(p.sync + 2, 1) == p.tiles - 1) { p.sync[0] = 0; p.sync[1] = 0; p.sync[2] = 0; }
    }
}

template <int BN, int CCH, int CSLOTS, bool SUB, bool SHADOW>
cudaError_t launch_u(cudaStream_t stream, int num_sms, const CUtensorMap& a, const CUtensorMap& b,
                     const CUtensorMap& c, const CUtensorMap& h, const UpdParams& p) {
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = std::max(1, std::min(tiles, num_sms));
    cudaError_t e = launch_pdl(tc_update_kernel<BN, CCH, CSLOTS, SUB, SHADOW>, dim3(grid), dim3(kThreads),
                               (size_t)UCfg<BN, CCH, CSLOTS>::SMEM_BYTES, stream, a, b, c, h, p);
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int BN, int CCH, int CSLOTS, bool SUB, bool SHADOW>
cudaError_t set_attr_u() {
    return cudaFuncSetAttribute(tc_update_kernel<BN, CCH, CSLOTS, SUB, SHADOW>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                UCfg<BN, CCH, CSLOTS>::SMEM_BYTES);
}

}  // namespace

cudaError_t tc_update_init() {
    cudaError_t e;
    if ((e = set_attr_u<128, 32, 4, true, true>()) != cudaSuccess) return e;
    if ((e = set_attr_u<256, 16, 2, true, true>()) != cudaSuccess) return e;
    if ((e = set_attr_u<128, 32, 4, false, false>()) != cudaSuccess) return e;
    if ((e = set_attr_u<256, 16, 2, false, false>()) != cudaSuccess) return e;
    return cudaSuccess;
}

// C[Mr x Nc] at (c_r0.., c_c0..) of the fp32 matrix Cmat (-)= Qh(row0:row0+Mr, colA:colA+K) * Bh[K x Nc]
// with the fp16 shadow of the new C written to the same coordinates of Hmat (sub only).
cudaError_t tc_update_tma(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_64,
                          const CUtensorMap& mapB_bn, int bn, int row0, int Mr, int colA, int K,
                          int colB0, int Nc, float* Cmat, long c_rows, long c_cols, long ldc, int c_c0,
                          __half* Hmat, long ldh, bool sub, float* colmax_part, int colmax_parts,
                          int shadow_from) {
    UpdParams p{};
    p.shadow_from = sub ? shadow_from : 0;
    p.colmax_part = sub ? colmax_part : nullptr;
    p.colmax_parts = colmax_parts;
    p.M = Mr; p.N = Nc;
    p.kb_total = (K + BK - 1) / BK;
    p.tiles_m = (Mr + BM - 1) / BM;
    p.tiles_n = (Nc + bn - 1) / bn;
    p.a_c0 = row0; p.a_c1 = colA; p.b_c1 = colB0;
    p.c_r0 = row0; p.c_c0 = c_c0;
    const int cch = bn == 256 ? 16 : 32;
    CUtensorMap mapC, mapH;
    cudaError_t e = make_plain_map(&mapC, Cmat, 4, c_rows, c_cols, ldc, BM, cch);
    if (e != cudaSuccess) return e;
    if (sub) {
        if ((e = make_plain_map(&mapH, Hmat, 2, c_rows, c_cols, ldh, BM, cch)) != cudaSuccess) return e;
        return bn == 256 ? launch_u<256, 16, 2, true, true>(stream, num_sms, mapQ_64, mapB_bn, mapC, mapH, p)
                         : launch_u<128, 32, 4, true, true>(stream, num_sms, mapQ_64, mapB_bn, mapC, mapH, p);
    }
    mapH = mapC;
    return bn == 256 ? launch_u<256, 16, 2, false, false>(stream, num_sms, mapQ_64, mapB_bn, mapC, mapH, p)
                     : launch_u<128, 32, 4, false, false>(stream, num_sms, mapQ_64, mapB_bn, mapC, mapH, p);
}

bool tc_node_supports(int num_sms, int m, int h) {
    return (h == 128 || h == 256) && m >= 1 && (m + BM - 1) / BM <= num_sms;
}

size_t tc_node_part_floats(int m, int h) { return (size_t)((m + BM - 1) / BM) * h * h; }

cudaError_t tc_node_init() {
    cudaError_t e = cudaFuncSetAttribute(tc_node_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         NodeCfg<128>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_node_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                NodeCfg<256>::SMEM_BYTES);
}

namespace {
// Programmatic dependent launch plus, if asked for, the cooperative attribute: the driver then guarantees that
// all CTAs of the grid are resident together (and keeps two such grids on one device from starving each other).
template <int H>
cudaError_t launch_node(int grid, cudaStream_t stream, bool cooperative, const CUtensorMap& q, const CUtensorMap& c,
                        const CUtensorMap& h, const NodeParams& p) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NODE_THREADS);
    cfg.dynamicSmemBytes = (size_t)NodeCfg<H>::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cooperative ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, tc_node_kernel<H>, q, c, h, p);
}
std::atomic<int> g_node_coop{0};   // 0 = untried, 1 = the driver takes cooperative + PDL launches, -1 = it does not
}  // namespace

int tc_node_cooperative_state() { return g_node_coop.load(); }

cudaError_t tc_node(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_128, int m, int h, int colQ, int colB,
                    float* Amat, long a_cols, long lda, __half* Hmat, long ldh, float* R12, long ldr, float* Z,
                    __half* R12h, float* part, int* sync, bool cooperative) {
    if (!tc_node_supports(num_sms, m, h) || !part || !sync || !R12h) return cudaErrorInvalidValue;
    NodeParams p{};
    p.tiles = (m + BM - 1) / BM;
    p.q_c1 = colQ; p.b_c1 = colB;
    p.part = part; p.R12 = R12; p.ldr = ldr; p.Z = Z; p.R12h = R12h; p.sync = sync;
    CUtensorMap mapC, mapH;
    cudaError_t e = make_plain_map(&mapC, Amat, 4, m, a_cols, lda, BM, 32);
    if (e != cudaSuccess) return e;
    if (h == 256) {
        if ((e = make_plain_map(&mapH, Hmat, 2, m, a_cols, ldh, BM, 32)) != cudaSuccess) return e;
    } else {
        mapH = mapC;   // (not used)
    }
    auto launch = [&](bool coop) {
        return h == 256 ? launch_node<256>(p.tiles, stream, coop, mapQ_128, mapC, mapH, p)
                        : launch_node<128>(p.tiles, stream, coop, mapQ_128, mapC, mapH, p);
    };
    // The first cooperative launch of the process is never made inside a stream capture (a rejected launch
    // would invalidate the capture): until the driver has accepted one, captures use the plain launch.
    bool coop = cooperative && g_node_coop.load() >= 0;
    if (coop && g_node_coop.load() == 0) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) coop = false;
    }
    e = launch(coop);
    if (coop) {
        if (e == cudaSuccess) {
            g_node_coop.store(1);
        } else {               // not launched: remember, clear the error, launch the plain way
            (void)cudaGetLastError();
            if (g_node_coop.exchange(-1) != -1)
                fprintf(stderr, "later_b200: cooperative launch of the fused node kernel rejected (%s); using plain launches\n",
                        cudaGetErrorString(e));
            e = launch(false);
        }
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace lb
