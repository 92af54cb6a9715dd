// LATER.h - C++-linkage entry points of the RGSQRF path, signature-compatible with the reference
// header of the same name (reference include/LATER.h:19-22, :39-47, :104-110, :140-154, :213-222)
// so that the reference's own test/test_qr.cu compiles and links against this library unchanged.
// Only the symbols that driver needs are declared; everything is a thin wrapper over the C ABI in
// later_b200.h.  Matrices are column-major fp32 on the device.
#pragma once

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cublas_v2.h>   // only for the handle types inside cudaCtxt; no cuBLAS call is made
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <curand.h>
#include <cusolverDn.h>

#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

// Passed BY VALUE by callers (reference include/LATER.h:19-22).  The handles are not used by this
// implementation; work runs on the legacy default stream, which is the stream the handles of the
// reference driver are bound to, so the caller's event timers bracket it the same way.
struct cudaCtxt {
    cublasHandle_t cublas_handle;
    cusolverDnHandle_t cusolver_handle;
};

// A = Q R, A (m x n, lda == m) overwritten by explicit Q, R (n x n) upper triangular.
// work / hwork are accepted for source compatibility and ignored (internal stream-ordered arena).
void later_rgsqrf(cudaCtxt ctxt, int m, int n, float* A, int lda, float* R, int ldr, float* work,
                  int lwork, __half* hwork, int lhwork);

// Explicit Q = I - W Y^T from a Householder WY pair (reference QR/later_ormqr.cu:18-85).
void later_ormqr(int m, int n, float* W, int ldw, float* Y, int ldy, float* work);
void later_ormqr2(int m, int n, float* W, int ldw, float* Y, int ldy, float* work);

// Recursive / blocked Householder QR in WY form (reference include/LATER.h:41,45; QR/later_rhouqr.cu,
// QR/later_bhouqr.cu): A <- Y, W, R out; later_ormqr (after later_rhouqr) or later_ormqr2 (after
// later_bhouqr) forms the explicit Q = I - W Y^T.  work, hwork and U are accepted and ignored.
void later_rhouqr(cudaCtxt ctxt, int m, int n, float* A, int lda, float* W, int ldw, float* R,
                  int ldr, float* work, int lwork, __half* hwork, int lhwork, float* U);
void later_bhouqr(int m, int n, float* A, int lda, float* W, int ldw, float* R, int ldr,
                  float* work, int lwork, __half* hwork, int lhwork, float* U);

// QDWH polar iteration, the reference's caller of later_rgsqrf (reference include/LATER.h:70,
// EVD/later_qdwh_polar.cu:24-110): tmpA (n x n, ld n) holds the matrix and is normalised in place,
// the top n x n block of A (2n x n, lda) receives the orthogonal polar factor.  H, work and hwork are
// accepted and ignored (the reference never writes H).
void later_qdwh_polar(cudaCtxt ctxt, int n, float* A, int lda, float* H, int ldh, float* tmpA, float* work,
                      __half* hwork);

// Utilities the reference driver calls (reference util/util.cu).
void startTimer();
float stopTimer();                                   // milliseconds since startTimer()
void generateUniformMatrix(float* dA, int m, int n); // cuRAND XORWOW, seed 3000, U(0,1]
void generateNormalMatrix(float* dA, int m, int n);
float snorm(int m, int n, float* dA);                // Frobenius norm of m*n contiguous floats
void print_env();

// Writes the m x n block of a device matrix as CSV (one row per line), as the reference's panel
// drivers expect (reference include/LATER.h:168-196).
template <typename T>
void printMatrixDeviceBlock(const char* filename, int m, int n, T* dA, int lda) {
    FILE* f = fopen(filename, "w");
    if (!f) { printf("cannot open %s\n", filename); return; }
    T* h = (T*)malloc(sizeof(T) * (size_t)lda * n);
    cudaMemcpy(h, dA, sizeof(T) * ((size_t)lda * (n - 1) + m), cudaMemcpyDeviceToHost);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) fprintf(f, j + 1 == n ? "%lf\n" : "%lf,", (double)h[i + (size_t)j * lda]);
    free(h);
    fclose(f);
}

#define gpuErrchk(ans) { gpuAssert((ans), __FILE__, __LINE__); }
inline void gpuAssert(cudaError_t code, const char* file, int line, bool abort = true) {
    if (code != cudaSuccess) {
        fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
        if (abort) exit(code);
    }
}
#ifdef DEBUG_CUDA_KERNEL_LAUNCH
#define CHECK_KERNEL() do { gpuErrchk(cudaDeviceSynchronize()); gpuErrchk(cudaPeekAtLastError()); } while (0)
#else
#define CHECK_KERNEL(x) do {} while (0)
#endif

__global__ void s2h(int m, int n, float* as, int ldas, __half* ah, int ldah);
__global__ void h2s(int m, int n, __half* ah, int ldah, float* as, int ldas);
__global__ void setEye(int m, int n, float* a, int lda);
__global__ void clearTri(char uplo, int m, int n, float* a, int lda);
__global__ void deviceCopy(int m, int n, float* da, int lda, float* db, int ldb);          // db <- da
__global__ void sSubstractAndSquare(int m, int n, float* dA, int lda, float* dB, int ldb);  // dB <- (dA - dB)^2
