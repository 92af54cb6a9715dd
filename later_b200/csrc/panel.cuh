// Tall-skinny 128-column panel factorisation (the recursion's base case).
//
// Replaces the reference's mgs_caqr_panel_256x128 / mgs_caqr_panel_256x32 / mgs_kernel2 chain
// (reference QR/panel.cu:10-134, :246-325): ~26 dependent launches per 128 columns there, four here.
//
// Algorithm: Gram-Schmidt in its Gram-matrix form.  All column norms and inter-column dot products
// of the panel (G = A^T A, 128x128) are accumulated in ONE pass over the panel with exact fp32xfp32
// products and fp64 sums; R = chol(G) is computed in fp64; Q = A R^-1 is applied in fp32 row by
// row with the same block structure the reference uses (32-column blocks: project out the earlier
// blocks, then normalise against the diagonal block).  With the Gram matrix held in fp64 the loss of
// orthogonality is O(kappa * eps_fp32) like MGS, as long as kappa(panel)^2 * eps_fp64 << 1.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstddef>

namespace lb {

constexpr int kPanelWidth = 128;

// ---- tensor-core kernels for tall panels (panel_tc.cu)
// Rows from which panel_qr128 switches from the fp32 forward-substitution apply (hidden behind the
// Cholesky kernel on short panels) to the split-precision tcgen05 apply (HBM-bound).
constexpr int kTcApplyMinRows = 65536;
// Rows from which the Gram matrix is formed on the integer tensor path instead of DMMA.
constexpr int kI8GramMinRows = 65536;

// Knobs of the panel path; read from the environment ONCE, when a context is created (tests and
// profiling only - the defaults are the product).
struct PanelOpts {
    int apply_tc = -1;                       // LB_APPLY_TC: -1 auto, 0 forward substitution, 1 tensor-core apply
    bool gram_i8 = true;                     // LB_GRAM_I8 = 0 keeps the fp64 (DMMA) Gram kernel everywhere
    int gram_i8_min_rows = kI8GramMinRows;   // LB_GRAM_I8_MIN_ROWS
    // LB_I8_FALLBACK_TAU: an integer-Gram panel whose smallest Cholesky pivot ratio piv_k / G_kk falls
    // below this is factored again from the fp64 Gram matrix (see panel_qr128)
    double i8_fallback_tau = 0.0078125;
};

// Status words of a factorisation (device int[kInfoWords], cleared by the caller before the first
// panel, read back after the last one; include/later_b200.h: later_b200_last_info).
constexpr int kInfoWords = 8;
enum PanelInfo : int {
    INFO_BAD_COLUMN = 0,   // 1 + first (global) column whose Cholesky pivot was not positive, 0 = none
    INFO_FLAGS = 1,        // bit 0: non-finite / out-of-range entry met by the integer Gram kernel
                           // bit 1: at least one panel fell back from the integer to the fp64 Gram matrix
    INFO_FALLBACKS = 2,    // number of such panels
    INFO_COND_LOG2 = 3,    // max over panels of ceil(-log2(min_k piv_k / G_kk)): ~ 2 log2(cond(panel))
    INFO_REDO = 4,         // scratch: "factor this panel again" flag between the two Cholesky launches
};

// Collective hook of the row-sharded (multi-GPU) factorisation: sums `count` doubles in place over all
// ranks, on `stream`.  With it the panel's Gram matrix is the Gram matrix of the GLOBAL panel, and every
// rank derives the same R from it.
struct PanelComm {
    void* self;
    // only_if (optional, device): a flag that holds the same value on every rank; when it reads 0 the
    // collective may be skipped (the peer-memory kernel does; NCCL cannot and sums stale data nobody reads)
    cudaError_t (*allreduce_f64)(void* self, double* buf, size_t count, cudaStream_t stream, const int* only_if);
};

// Scratch (bytes) needed by panel_qr128 for an m-row panel on a device with num_sms SMs.
size_t panel_scratch_bytes(int m, int num_sms);

// A[m x 128] (fp32, ld lda) -> Q in place; R[128 x 128] (fp32, ld ldr) upper triangular with the
// strictly lower part zeroed; Qh (optional) receives the fp16 copy of Q (ld ldqh).
// allow_tc: tall panels (m >= kTcApplyMinRows) may form Q with the split-precision tensor-core
// apply (panel_tc.cu), whose Q is accurate to ~1e-6 instead of ~1e-7: the recursion, which consumes
// Q rounded to fp16 anyway, says yes; the stand-alone panel entry point, which replaces the
// reference's all-fp32 panel, says no.
// colmax_ready: panel_colmax_scratch() already holds this panel's column maxima.
// info: the factorisation's status words (see PanelInfo); col0: global index of the panel's first
// column (for INFO_BAD_COLUMN).
//
// Breakdown handling.  CholeskyQR needs G accurate relative to its smallest eigenvalue.  The fp64
// Gram matrix (exact products, fp64 sums) is good for cond(panel) up to ~1e6-1e7; the integer Gram
// matrix drops digit pairs below 2^-26 of a product, a systematic error of ~1e-8 |G|, so
// its panels are checked: if the smallest pivot ratio piv_k / G_kk (~ 1 / cond^2) is below
// opts.i8_fallback_tau, or a pivot is not positive, the panel is factored AGAIN from the fp64 Gram
// matrix, and Q = A R^-1 is then formed by the fp32 forward substitution (row-wise backward stable)
// instead of the tensor-core product with the explicit inverse, whose backward error grows with
// cond(R) (measured at 131072 x 128, cond 1e6: 6e-7 against the reference's 1.5e-7).  The fallback
// kernels are always enqueued (the launch sequence must not depend on data: it is replayed from a
// CUDA graph) and exit at once unless the flag is set; the tensor-core apply exits at once if it is.
// A non-positive pivot on the fp64 path is clamped and reported in info[INFO_BAD_COLUMN].
cudaError_t panel_qr128(cudaStream_t stream, int num_sms, int m, float* A, long lda, float* R,
                        long ldr, __half* Qh, long ldqh, void* scratch, bool allow_tc,
                        const PanelOpts& opts, int* info, int col0, bool colmax_ready = false,
                        const PanelComm* comm = nullptr);

cudaError_t panel_init();
// Which kernels panel_qr128 will use (4 launches per panel with forward substitution, 5 with the
// tensor-core apply: + the triangular inverse; + 5 or 6 with the integer Gram: column maxima unless
// ready, the Gram kernel itself, the three conditional fallback kernels and the conditional
// forward-substitution apply).
bool panel_uses_tc_apply(int m, const float* A, long lda, bool allow_tc, const PanelOpts& opts);
bool panel_uses_i8_gram(int m, int num_sms, const float* A, long lda, bool allow_tc, const PanelOpts& opts);
int panel_launch_count(int m, int num_sms, const float* A, long lda, bool allow_tc, const PanelOpts& opts);

struct TcApplyFactors {
    // three fp16 planes t1 + t2 + t3 = fp32(diag(1/s) R^-1) * 2^e exactly, column-major
    __half T[3][kPanelWidth * kPanelWidth];
    float colscale[kPanelWidth];             // s_k: power of two with ||a_k|| s_k in [2^13, 2^14)
    float unscale;                           // 2^-e
    float pad[127];
};

cudaError_t tc_apply_init();
// Integer-tensor-core Gram matrix of an m x 128 panel into per-CTA partials laid out like the DMMA
// kernel's ([cta][upper 32 x 32 block][r][c] doubles); *_grid = number of partials it writes.
int panel_gram_i8_grid(int m, int num_sms);
bool panel_gram_i8_fits(int m, int num_sms);
// colmax_part: kColmaxParts floats of scratch per column (written by the kernel's own first pass).
constexpr int kColmaxParts = 160;   // >= the grid of any kernel that fills them (colmax128: 64, update: <= SMs)
// colmax_ready: the partials were already written by the update kernel that produced the panel.
cudaError_t panel_gram_i8(cudaStream_t stream, int num_sms, int m, const float* A, long lda,
                          float* colmax_part, bool colmax_ready, double* part, int* info);
// Where in a panel scratch area (panel_scratch_bytes(m, num_sms)) the column-maxima partials live.
float* panel_colmax_scratch(void* scratch, int m, int num_sms);
// Q = A R^-1 for an m x 128 panel whose R (fp32, upper triangular) is already in place; needs
// lda % 4 == 0 and a 16-byte aligned A.
// skip (optional, device): when *skip != 0 both kernels return at once (the panel was factored again
// from the fp64 Gram matrix and is applied by forward substitution).
cudaError_t panel_apply_tc(cudaStream_t stream, int num_sms, int m, float* A, long lda, const float* R,
                           long ldr, __half* Qh, long ldqh, TcApplyFactors* fac, const int* skip = nullptr);

}  // namespace lb
