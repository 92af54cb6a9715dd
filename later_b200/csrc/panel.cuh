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

// Scratch (bytes) needed by panel_qr128 for an m-row panel on a device with num_sms SMs.
size_t panel_scratch_bytes(int m, int num_sms);

// A[m x 128] (fp32, ld lda) -> Q in place; R[128 x 128] (fp32, ld ldr) upper triangular with the
// strictly lower part zeroed; Qh (optional) receives the fp16 copy of Q (ld ldqh).
// allow_tc: tall panels (m >= kTcApplyMinRows) may form Q with the split-precision tensor-core
// apply (panel_tc.cu), whose Q is accurate to ~1e-6 instead of ~1e-7: the recursion, which consumes
// Q rounded to fp16 anyway, says yes; the stand-alone panel entry point, which replaces the
// reference's all-fp32 panel, says no.
// colmax_ready: panel_colmax_scratch() already holds this panel's column maxima.
cudaError_t panel_qr128(cudaStream_t stream, int num_sms, int m, float* A, long lda, float* R,
                        long ldr, __half* Qh, long ldqh, void* scratch, bool allow_tc,
                        bool colmax_ready = false);

cudaError_t panel_init();
// Which apply panel_qr128 will use (4 launches per panel with forward substitution, 5 with the
// tensor-core apply: + the triangular inverse).
bool panel_uses_tc_apply(int m, const float* A, long lda, bool allow_tc);
bool panel_uses_i8_gram(int m, int num_sms, const float* A, long lda, bool allow_tc);
int panel_launch_count(int m, int num_sms, const float* A, long lda, bool allow_tc);

// ---- tensor-core apply for tall panels (panel_tc.cu)
// Rows from which panel_qr128 switches from the fp32 forward-substitution apply (hidden behind the
// Cholesky kernel on short panels) to the split-precision tcgen05 apply (HBM-bound).
constexpr int kTcApplyMinRows = 65536;
// Rows from which the Gram matrix is formed on the integer tensor path instead of DMMA.
constexpr int kI8GramMinRows = 65536;

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
cudaError_t panel_apply_tc(cudaStream_t stream, int num_sms, int m, float* A, long lda, const float* R,
                           long ldr, __half* Qh, long ldqh, TcApplyFactors* fac);

}  // namespace lb
