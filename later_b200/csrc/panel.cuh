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
cudaError_t panel_qr128(cudaStream_t stream, int num_sms, int m, float* A, long lda, float* R,
                        long ldr, __half* Qh, long ldqh, void* scratch);

cudaError_t panel_init();

}  // namespace lb
