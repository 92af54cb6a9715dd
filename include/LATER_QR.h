// LATER_QR.h - panel entry points, signature-compatible with the reference's QR/include/LATER_QR.h
// (reference QR/include/LATER_QR.h:20-34), so that the reference's panel drivers
// (test/test_mgs_panel.cu, test/test_caqr_panel.cu) compile and link against this library.
#pragma once
#include "LATER.h"

// QR of a tall-skinny m x 128 panel: A <- Q, R (128 x 128 block, leading dimension ldr) <- upper
// triangular factor.  `work` is accepted for source compatibility and ignored.
// (reference QR/panel.cu:10-63)
void mgs_caqr_panel_256x128(cudaCtxt ctxt, int m, int n, float* A, int lda, float* R, int ldr,
                            float* work);

// QR of an m x 32 strip, n must be 32 (reference QR/panel.cu:65-134).  `work` is ignored.
void mgs_caqr_panel_256x32(cudaCtxt ctxt, int m, int n, float* A, int lda, float* R, int ldr,
                           float* work);

// QR of every 256-row block of an m x n (n <= 32) strip on its own; block b's R goes to rows
// 32 b .. 32 b + 31 of RR (leading dimension ldr), zeros below the diagonal.  Launch as the
// reference does: mgs_kernel2<<<blocks, dim3(32, 32)>>>, mgs_kernel<<<blocks, 256>>>
// (reference QR/panel.cu:136-235, :246-325).
__global__ void mgs_kernel(int m, int n, float* AA, int lda, float* RR, int ldr);
__global__ void mgs_kernel2(int m, int n, float* AA, int lda, float* RR, int ldr);

// Householder CAQR strip (reference QR/panel.cu:341-378): A (m x 32) <- explicit Q, R <- its factor.
// Only the <256, 32> instance exists, as in the reference (QR/panel.cu:811).
template <int M, int N>
void hou_caqr_panel(cudaCtxt ctxt, int m, int n, float* A, int lda, float* R, int ldr, float* work);
