// TEST INFRASTRUCTURE.  extern "C" shim around the UNMODIFIED reference entry points so that
// Python (ctypes) can run the reference's own code on a device buffer for parity and for the
// `bench.py --impl reference` arm.  Compiled together with the reference sources by oracle/Makefile;
// contains no algorithm of its own.
#include "LATER.h"
#include "LATER_QR.h"

static cudaCtxt g_ctxt;
static bool g_init = false;
static void ensure() {
    if (!g_init) {
        cublasCreate(&g_ctxt.cublas_handle);
        cusolverDnCreate(&g_ctxt.cusolver_handle);
        g_init = true;
    }
}

extern "C" {
// reference later_rgsqrf (QR/later_rgsqrf.cu:62) with caller-provided work buffers sized as in
// test/test_qr.cu:73-78: work = m/256*32*n floats, hwork = m*n halves.
int ref_later_rgsqrf(int m, int n, float* A, int lda, float* R, int ldr, float* work, int lwork,
                     void* hwork, int lhwork) {
    ensure();
    later_rgsqrf(g_ctxt, m, n, A, lda, R, ldr, work, lwork, (__half*)hwork, lhwork);
    return (int)cudaGetLastError();
}
int ref_mgs_caqr_panel_256x128(int m, int n, float* A, int lda, float* R, int ldr, float* work) {
    ensure();
    mgs_caqr_panel_256x128(g_ctxt, m, n, A, lda, R, ldr, work);
    return (int)cudaGetLastError();
}
int ref_later_ormqr(int m, int n, float* W, int ldw, float* Y, int ldy, float* work) {
    later_ormqr(m, n, W, ldw, Y, ldy, work);
    return (int)cudaGetLastError();
}
int ref_later_ormqr2(int m, int n, float* W, int ldw, float* Y, int ldy, float* work) {
    later_ormqr2(m, n, W, ldw, Y, ldy, work);
    return (int)cudaGetLastError();
}
int ref_mgs_caqr_panel_256x32(int m, int n, float* A, int lda, float* R, int ldr, float* work) {
    ensure();
    mgs_caqr_panel_256x32(g_ctxt, m, n, A, lda, R, ldr, work);
    return (int)cudaGetLastError();
}
// the reference's 256 x 32 block kernel with the launch configuration of its callers
// (QR/panel.cu:79,92,112): one (32, 32) block per 256 rows, R_b at rows 32 b of RR
int ref_mgs_kernel2(int m, int n, float* A, int lda, float* RR, int ldr) {
    mgs_kernel2<<<(m + 255) / 256, dim3(32, 32)>>>(m, n, A, lda, RR, ldr);
    return (int)cudaGetLastError();
}
// reference later_qdwh_polar (EVD/later_qdwh_polar.cu:24): tmpA (n x n) in, top block of A (2n x n) out
int ref_later_qdwh_polar(int n, float* A, int lda, float* tmpA, float* work, void* hwork) {
    ensure();
    later_qdwh_polar(g_ctxt, n, A, lda, nullptr, n, tmpA, work, (__half*)hwork);
    return (int)cudaGetLastError();
}
// reference later_rhouqr / later_bhouqr (QR/later_rhouqr.cu:26, QR/later_bhouqr.cu): A <- Y, W, R
int ref_later_rhouqr(int m, int n, float* A, int lda, float* W, int ldw, float* R, int ldr, float* work, int lwork,
                     void* hwork, int lhwork, float* U) {
    ensure();
    later_rhouqr(g_ctxt, m, n, A, lda, W, ldw, R, ldr, work, lwork, (__half*)hwork, lhwork, U);
    return (int)cudaGetLastError();
}
int ref_later_bhouqr(int m, int n, float* A, int lda, float* W, int ldw, float* R, int ldr, float* work, int lwork,
                     void* hwork, int lhwork, float* U) {
    later_bhouqr(m, n, A, lda, W, ldw, R, ldr, work, lwork, (__half*)hwork, lhwork, U);
    return (int)cudaGetLastError();
}
void ref_generate_uniform(float* dA, int m, int n) { generateUniformMatrix(dA, m, n); }
}
