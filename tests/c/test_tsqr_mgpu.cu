// C++ caller of the multi-GPU entry point (include/later_b200.h: later_b200_tsqr_mgpu): what a user of the
// reference's LATER.h would write to factor a tall-skinny matrix that is row-sharded over the GPUs of a
// node.  Checks, on the host in fp64: the global backward error ||A - Q R|| / ||A||, the global
// orthogonality ||I - Q^T Q|| / n, identical R on every device, R upper triangular with r_ii > 0.
//   nvcc -std=c++17 -I include tests/c/test_tsqr_mgpu.cu later_b200/liblater_b200.so -o t && ./t P m_local n [tsqr]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#include "later_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 2;
    const int m_local = argc > 2 ? atoi(argv[2]) : 16384;
    const int n = argc > 3 ? atoi(argv[3]) : 256;
    const bool use_tsqr = argc > 4 && strcmp(argv[4], "tsqr") == 0;      // default: the row-sharded recursion
    setvbuf(stdout, nullptr, _IONBF, 0);
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (count < P) { printf("SKIP: %d device(s), need %d\n", count, P); return 77; }
    std::vector<int> devices(P);
    for (int p = 0; p < P; ++p) devices[p] = p;
    later_b200_mgpu* g = nullptr;
    int rc = later_b200_mgpu_create(&g, P, devices.data());
    if (rc != 0) { printf("FAIL: later_b200_mgpu_create rc=%d\n", rc); return 1; }

    const long m = (long)P * m_local;
    std::vector<float> A0((size_t)m * n);              // global matrix, stored per shard: [p][col][row]
    std::mt19937 gen(1234);
    std::normal_distribution<float> dist(0.f, 1.f);
    for (auto& v : A0) v = dist(gen);
    std::vector<float*> dA(P), dR(P);
    for (int p = 0; p < P; ++p) {
        CK(cudaSetDevice(p));
        CK(cudaMalloc(&dA[p], sizeof(float) * (size_t)m_local * n));
        CK(cudaMalloc(&dR[p], sizeof(float) * (size_t)n * n));
        CK(cudaMemcpy(dA[p], A0.data() + (size_t)p * m_local * n, sizeof(float) * (size_t)m_local * n, cudaMemcpyHostToDevice));
        CK(cudaMemset(dR[p], 0xff, sizeof(float) * (size_t)n * n));       // every entry of R must be written
    }
    for (int rep = 0; rep < 3; ++rep) {                // direct launch, graph capture, graph replay
        if (rep > 0)
            for (int p = 0; p < P; ++p) {
                CK(cudaSetDevice(p));
                CK(cudaMemcpy(dA[p], A0.data() + (size_t)p * m_local * n, sizeof(float) * (size_t)m_local * n, cudaMemcpyHostToDevice));
            }
        rc = use_tsqr ? later_b200_tsqr_mgpu(g, m_local, n, dA.data(), m_local, dR.data(), n)
                      : later_b200_rgsqrf_mgpu(g, m_local, n, dA.data(), m_local, dR.data(), n);
        if (rc == 0) rc = later_b200_mgpu_sync(g);
        if (rc != 0) { printf("FAIL: later_b200_tsqr_mgpu rc=%d: %s\n", rc, later_b200_mgpu_last_error(g)); return 1; }
    }
    std::vector<float> Q((size_t)m * n), R((size_t)n * n), Rp((size_t)n * n);
    bool same = true;
    for (int p = 0; p < P; ++p) {
        CK(cudaSetDevice(p));
        CK(cudaMemcpy(Q.data() + (size_t)p * m_local * n, dA[p], sizeof(float) * (size_t)m_local * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(p == 0 ? R.data() : Rp.data(), dR[p], sizeof(float) * (size_t)n * n, cudaMemcpyDeviceToHost));
        if (p > 0) same = same && memcmp(R.data(), Rp.data(), sizeof(float) * (size_t)n * n) == 0;
    }
    bool structure = true;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            const float v = R[i + (size_t)j * n];
            if (i > j && v != 0.f) structure = false;
            if (i == j && !(v > 0.f)) structure = false;
        }
    // ||A - Q R||_F / ||A||_F and ||I - Q^T Q||_F / n in fp64 (sampled rows for the residual keep this quick)
    double res2 = 0, nrm2 = 0;
    std::vector<double> G((size_t)n * n, 0.0);
    for (int p = 0; p < P; ++p) {
        const float* q = Q.data() + (size_t)p * m_local * n;
        const float* a = A0.data() + (size_t)p * m_local * n;
        for (int i = 0; i < m_local; i += 37) {
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int k = 0; k <= j; ++k) s += (double)q[i + (size_t)k * m_local] * R[k + (size_t)j * n];
                const double d = a[i + (size_t)j * m_local] - s;
                res2 += d * d;
                nrm2 += (double)a[i + (size_t)j * m_local] * a[i + (size_t)j * m_local];
            }
        }
        for (int j = 0; j < n; ++j)
            for (int k = 0; k <= j; ++k) {
                double s = 0;
                for (int i = 0; i < m_local; ++i) s += (double)q[i + (size_t)k * m_local] * q[i + (size_t)j * m_local];
                G[k + (size_t)j * n] += s;
            }
    }
    double orth2 = 0;
    for (int j = 0; j < n; ++j)
        for (int k = 0; k <= j; ++k) {
            const double d = G[k + (size_t)j * n] - (k == j ? 1.0 : 0.0);
            orth2 += (k == j ? 1.0 : 2.0) * d * d;
        }
    const double back = std::sqrt(res2 / nrm2), orth = std::sqrt(orth2) / n;
    // (the recursion keeps single-GPU accuracy; the TSQR variant rounds Q and W to fp16 once more)
    const bool ok = same && structure && back < (use_tsqr ? 5e-4 : 5e-5) && orth < 5e-5;
    printf("%s P=%d %ldx%d: same_R=%d structure=%d backward=%.3e orth/n=%.3e %s\n", use_tsqr ? "tsqr_mgpu" : "rgsqrf_mgpu", P, m, n, (int)same,
           (int)structure, back, orth, ok ? "OK" : "FAIL");
    for (int p = 0; p < P; ++p) { cudaSetDevice(p); cudaFree(dA[p]); cudaFree(dR[p]); }
    later_b200_mgpu_destroy(g);
    return ok ? 0 : 1;
}
