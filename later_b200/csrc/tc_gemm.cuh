// Host-side interface of the tcgen05 GEMM kernels that carry RGSQRF's trailing updates.
//
// They replace the two cublasGemmEx calls plus the three separate s2h cast kernels of the
// reference recursion (reference QR/later_rgsqrf.cu:41-56):
//   gram   : R12 = Q1^T * A2      (both operands K-major: K runs down the rows of the column-major
//                                   fp16 shadow, so it is the contiguous dimension)
//   update : A2 -= Q1 * R12       (A operand MN-major, B operand K-major, fp32 read-modify-write)
//   assign : C   = Qh * W         (same operand layout as update, no C read; TSQR back-multiply)
// Operands are fp16 ("shadow" copies written by the producing kernels' epilogues), accumulation is
// fp32 in TMEM, exactly the arithmetic cublasGemmEx(CUDA_R_16F in, CUDA_R_32F compute) performs.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace lb {

// A 2-D fp16 matrix in device memory, column-major, as the TMA unit sees it.
struct HalfMatrix {
    const __half* ptr;  // element (0,0)
    int rows;           // extent of the contiguous dimension
    int cols;
    long ld;            // leading dimension in elements (multiple of 8 -> 16-byte strides)
};

enum EpiMode : int {
    EPI_STORE = 0,    // C = D            (+ optional fp16 copy)
    EPI_SUB = 1,      // C = C - D        (+ optional fp16 copy of the new C)
    EPI_PARTIAL = 2,  // part[split] = D  (split-K; reduced by splitk_reduce)
    EPI_ADD = 3,      // C = C + D
};

struct TcGemmParams {
    int M, N;               // output extent
    int kb_total;           // 64-wide k blocks over the whole K
    int kb_per_split;
    int splits;
    int tiles_m, tiles_n;
    int a_c0, a_c1;         // TMA origin of the A operand (inner, outer)
    int b_c0, b_c1;         // TMA origin of the B operand (inner = K, outer = N)
    float* C;
    long ldc;
    __half* Ch;             // optional fp16 mirror of C
    long ldch;
    float* part;            // split-K workspace: [splits][N][M]
    const float* dscale;    // optional device scalar multiplied into D before the epilogue op
    float* Z;               // optional: a block of the same shape and ld as C that is set to zero
                            // (R21 of the recursion node: never produced, must read as 0)
};

// Encodes a SWIZZLE_128B tiled tensor map over a column-major fp16 matrix with box
// {box_inner (<=64), box_outer (<=256)}.  Returns cudaSuccess or the failure.
cudaError_t make_tensor_map_f16(CUtensorMap* out, const HalfMatrix& mat, int box_inner,
                                int box_outer);

struct TcGemmPlan {
    int bn;        // 128 or 256
    int splits;
    int kb_per_split;
    int grid;
};

// R12[Mc x Nc] (fp32, ld ldc) = A1^T * A2 with A1 = Q(:, colA : colA+Mc), A2 = Q(:, colB : colB+Nc),
// K = rows [row0, row0 + k_rows) of the fp16 shadow Q.  Optional fp16 copy of R12 into Ch.
// `part` must hold splits*Mc*Nc floats when the plan uses split-K.
cudaError_t tc_gram(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_128,
                    const CUtensorMap& mapQ_bn, int bn, int row0, int k_rows, int colA, int Mc,
                    int colB, int Nc, float* C, long ldc, __half* Ch, long ldch, float* part,
                    int splits, float* Z = nullptr);

// C[Mr x Nc] (fp32, ld ldc) (-)= Qh(row0:row0+Mr, colA:colA+K) * Bh[K x Nc]; Bh is addressed through
// its own tensor map with origin (0, colB0).  sub=true: C -= D; sub=false: C = D.  Optional fp16
// mirror of the new C into Ch.
cudaError_t tc_update(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_64,
                      const CUtensorMap& mapB_bn, int bn, int row0, int Mr, int colA, int K,
                      int colB0, int Nc, float* C, long ldc, __half* Ch, long ldch, bool sub);

// Generic launch: a_mn_major=false -> both operands K-major ("gram" layout, modes STORE / PARTIAL /
// ADD); a_mn_major=true -> A MN-major ("update" layout, modes SUB / STORE).
cudaError_t tc_gemm_launch(cudaStream_t stream, int num_sms, bool a_mn_major, int bn, int epi,
                           const CUtensorMap& mapA, const CUtensorMap& mapB, const TcGemmParams& p);
void tc_fill_gram(TcGemmParams& p, int bn, int row0, int k_rows, int colA, int Mc, int colB, int Nc,
                  float* C, long ldc, __half* Ch, long ldch);
void tc_fill_update(TcGemmParams& p, int bn, int row0, int Mr, int colA, int K, int colB0, int Nc,
                    float* C, long ldc, __half* Ch, long ldch);

// Sums split-K partials in a fixed order (deterministic) and writes C (+ fp16 copy).
cudaError_t splitk_reduce(cudaStream_t stream, const float* part, int splits, int M, int N, float* C,
                          long ldc, __half* Ch, long ldch, float* Z = nullptr);

// Picks split-K so that tiles*splits roughly fills the machine.
int choose_gram_splits(int num_sms, int Mc, int Nc, int bn, int k_rows);

// Split-K Gram product with the B operand taken from an fp32 matrix and rounded to fp16 on the way
// into shared memory (tc_gram_cast.cu): C[Mc x Nc] = Qh(:, colA:colA+Mc)^T * fp16(B[k_rows x Nc]).
// Same splits, same accumulation order and therefore the same bits as cast + tc_gram.  splits >= 2.
// Mc must be 128, 256 or 512 (tc_gram_cast_supports): one CTA holds all Mc / 128 row tiles.
cudaError_t tc_gram_cast(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_128, int k_rows,
                         int colA, int Mc, const float* B, long ldb, int Nc, float* C, long ldc, __half* Ch,
                         long ldch, float* part, int splits, float* Z);
bool tc_gram_cast_supports(int Mc);
// split-K factor of the cast-fused product for a node of half-width Mc (its CTAs own Mc x 128 output
// strips, so the generic choose_gram_splits, which counts 128 x bn tiles, would fill half the SMs)
int tc_gram_cast_splits(int num_sms, int Mc, int k_rows);
cudaError_t tc_gram_cast_init();

// Sets the dynamic shared-memory limits of all kernel instantiations (once per device).
cudaError_t tc_gemm_init();
cudaError_t tc_update_init();

// A whole small node of the recursion (half-width h = 128 or 256: R12 = Q1^T A2, A2 -= Q1 R12; Q1 = columns
// colQ.., A2 = columns colB.. of the m-row matrix Amat / of its fp16 shadow Hmat, which mapQ_128 also covers) in
// one launch, one CTA per 128-row tile with two grid barriers (tc_update.cu) - for m <= 128 * num_sms.  Writes
// R12 (fp32, ld ldr), clears the mirror block Z (optional), leaves fp16 R12 in R12h (ld h); the shadow of the
// new A2 is written from its column 128 on only (the caller factors the first 128 columns next).
// part: tc_node_part_floats(m, h) floats; sync: three ints, zero before the first launch (the kernel leaves
// them zero).
bool tc_node_supports(int num_sms, int m, int h);
size_t tc_node_part_floats(int m, int h);
cudaError_t tc_node(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_128, int m, int h, int colQ, int colB,
                    float* Amat, long a_cols, long lda, __half* Hmat, long ldh, float* R12, long ldr, float* Z,
                    __half* R12h, float* part, int* sync, bool cooperative);
cudaError_t tc_node_init();
// cooperative = launch with the cooperative attribute on top of programmatic dependent launch (co-residency of
// the grid guaranteed by the driver); falls back to the plain launch, once and for all, if the driver rejects the
// combination.  tc_node_cooperative_state(): 0 untried, 1 accepted, -1 rejected.
int tc_node_cooperative_state();

// Trailing update with the C tile streamed through shared memory by TMA (tc_update.cu):
// C block = Cmat(row0:row0+Mr, c_c0:c_c0+Nc) (-)= Qh(row0:row0+Mr, colA:colA+K) * Bh(:, colB0:colB0+Nc).
// Cmat is the whole column-major fp32 matrix (c_rows x c_cols, ld ldc); with sub=true the fp16
// shadow of the new C block goes to the same coordinates of Hmat (ld ldh).  Nc must be a multiple
// of bn unless the block ends at the matrix edge.  colmax_part (sub only, optional): per-CTA partial
// maxima of |new C| over the block's first 128 columns, [col][colmax_parts] floats, every slot written.
// shadow_from (sub only, a multiple of 32): the shadow of the block's first shadow_from columns is not
// written - for columns whose shadow the next panel's apply kernel rewrites before anybody reads it.
cudaError_t tc_update_tma(cudaStream_t stream, int num_sms, const CUtensorMap& mapQ_64,
                          const CUtensorMap& mapB_bn, int bn, int row0, int Mr, int colA, int K,
                          int colB0, int Nc, float* Cmat, long c_rows, long c_cols, long ldc, int c_c0,
                          __half* Hmat, long ldh, bool sub, float* colmax_part = nullptr,
                          int colmax_parts = 0, int shadow_from = 0);

}  // namespace lb
