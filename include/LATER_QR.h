// LATER_QR.h - panel entry points, signature-compatible with the reference's QR/include/LATER_QR.h
// (reference QR/include/LATER_QR.h:20-25) for the Gram-Schmidt path only.
#pragma once
#include "LATER.h"

// QR of a tall-skinny m x 128 panel: A <- Q, R (128 x 128 block, leading dimension ldr) <- upper
// triangular factor.  `work` is accepted for source compatibility and ignored.
void mgs_caqr_panel_256x128(cudaCtxt ctxt, int m, int n, float* A, int lda, float* R, int ldr,
                            float* work);
