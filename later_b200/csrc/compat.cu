// C++-linkage LATER.h entry points (include/LATER.h) as thin wrappers over the C ABI, plus the
// small utilities the reference's test/test_qr.cu calls (reference util/util.cu).  A process-global
// context on device 0 / legacy default stream stands in for the reference's implicit global state.
#include "../../include/LATER.h"
#include "../../include/LATER_QR.h"
#include "../../include/later_b200.h"

#include <cmath>
#include <mutex>

namespace {

later_b200_ctx* default_ctx() {
    static later_b200_ctx* ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        cudaGetDevice(&dev);
        int rc = later_b200_create(&ctx, dev, nullptr);
        if (rc != 0) {
            fprintf(stderr, "later_b200: cannot create context on device %d (rc=%d): "
                            "an sm_100 GPU is required, there is no fallback path\n", dev, rc);
            exit(3);
        }
    });
    return ctx;
}

void die_on(int rc, const char* what) {
    if (rc == LATER_B200_ERANK) {   // the factorisation ran; the reference would carry on silently
        fprintf(stderr, "later_b200: %s: %s\n", what, later_b200_last_error(default_ctx()));
        return;
    }
    if (rc != 0) {
        fprintf(stderr, "later_b200: %s failed (rc=%d): %s\n", what, rc,
                later_b200_last_error(default_ctx()));
        exit(4);
    }
}

cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;

// Sum of squares with fp64 accumulation, two stages, fixed order.
__global__ void sumsq_partial(const float* __restrict__ x, long n, double* __restrict__ part) {
    __shared__ double sh[256];
    double s = 0.0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double v = x[i];
        s += v * v;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

}  // namespace

void later_rgsqrf(cudaCtxt, int m, int n, float* A, int lda, float* R, int ldr, float*, int,
                  __half*, int) {
    // The reference ignores lda and uses m (QR/later_rgsqrf.cu:70); lda == m is its precondition.
    die_on(later_b200_rgsqrf(default_ctx(), m, n, A, lda, R, ldr), "later_rgsqrf");
}

void mgs_caqr_panel_256x128(cudaCtxt, int m, int n, float* A, int lda, float* R, int ldr, float*) {
    die_on(later_b200_panel_qr(default_ctx(), m, n, A, lda, R, ldr), "mgs_caqr_panel_256x128");
}

void mgs_caqr_panel_256x32(cudaCtxt, int m, int n, float* A, int lda, float* R, int ldr, float*) {
    if (n != 32) {   // the reference's message and behaviour (QR/panel.cu:67-71)
        printf("[Error]: CAQR_32 does not support n!=32\n");
        return;
    }
    die_on(later_b200_panel32_qr(default_ctx(), m, n, A, lda, R, ldr), "mgs_caqr_panel_256x32");
}

// Householder CAQR strip (reference QR/panel.cu:341-378): A <- explicit Q, R <- 32 x 32 factor.  Served by
// the Gram/Cholesky strip factorisation (r_ii > 0; the reference's reflectors give r_ii < 0).
template <int M, int N>
void hou_caqr_panel(cudaCtxt, int m, int n, float* A, int lda, float* R, int ldr, float*) {
    static_assert(N == 32, "only the 32-column strip exists");
    die_on(later_b200_panel32_qr(default_ctx(), m, n, A, lda, R, ldr), "hou_caqr_panel");
}
template void hou_caqr_panel<256, 32>(cudaCtxt, int, int, float*, int, float*, int, float*);

void later_ormqr(int m, int n, float* W, int ldw, float* Y, int ldy, float*) {
    die_on(later_b200_ormqr(default_ctx(), m, n, W, ldw, Y, ldy), "later_ormqr");
}

void later_ormqr2(int m, int n, float* W, int ldw, float* Y, int ldy, float*) {
    die_on(later_b200_ormqr2(default_ctx(), m, n, W, ldw, Y, ldy), "later_ormqr2");
}

void later_qdwh_polar(cudaCtxt, int n, float* A, int lda, float*, int, float* tmpA, float*, __half*) {
    int iters = 0;
    die_on(later_b200_qdwh_polar(default_ctx(), n, tmpA, n, A, lda, 0.f, 0, &iters), "later_qdwh_polar");
    printf("later_qdwh_polar: %d iterations\n", iters);
}

void later_rhouqr(cudaCtxt, int m, int n, float* A, int lda, float* W, int ldw, float* R, int ldr, float*, int, __half*,
                  int, float*) {
    die_on(later_b200_rhouqr(default_ctx(), m, n, A, lda, W, ldw, R, ldr, 0), "later_rhouqr");
}

void later_bhouqr(int m, int n, float* A, int lda, float* W, int ldw, float* R, int ldr, float*, int, __half*, int,
                  float*) {
    die_on(later_b200_rhouqr(default_ctx(), m, n, A, lda, W, ldw, R, ldr, 1), "later_bhouqr");
}

void startTimer() {
    if (!g_t0) { cudaEventCreate(&g_t0); cudaEventCreate(&g_t1); }
    cudaEventRecord(g_t0, 0);
}

float stopTimer() {
    float ms = 0.f;
    cudaEventRecord(g_t1, 0);
    cudaEventSynchronize(g_t1);
    cudaEventElapsedTime(&ms, g_t0, g_t1);
    return ms;
}

void generateUniformMatrix(float* dA, int m, int n) {
    // identical stream to the reference: default pseudo generator (XORWOW), seed 3000
    curandGenerator_t gen;
    curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT);
    curandSetPseudoRandomGeneratorSeed(gen, 3000ULL);
    curandGenerateUniform(gen, dA, (size_t)m * n);
    curandDestroyGenerator(gen);
}

void generateNormalMatrix(float* dA, int m, int n) {
    curandGenerator_t gen;
    curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT);
    curandSetPseudoRandomGeneratorSeed(gen, (unsigned long long)(rand() % 3000));
    curandGenerateNormal(gen, dA, (size_t)m * n, 0.f, 1.f);
    curandDestroyGenerator(gen);
}

float snorm(int m, int n, float* dA) {
    const long total = (long)m * n;
    const int blocks = 592;
    double* part = nullptr;
    cudaMalloc(&part, blocks * sizeof(double));
    sumsq_partial<<<blocks, 256>>>(dA, total, part);
    double h[592];
    cudaMemcpy(h, part, blocks * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(part);
    double s = 0.0;
    for (int i = 0; i < blocks; ++i) s += h[i];
    return (float)std::sqrt(s);
}

void print_env() {
    int dev = 0, rt = 0, drv = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, dev);
    cudaRuntimeGetVersion(&rt);
    cudaDriverGetVersion(&drv);
    printf("=== later_b200 device ===\n%s, sm_%d%d, %d SMs, %.1f GiB, L2 %d MiB, CUDA rt %d drv %d\n\n",
           prop.name, prop.major, prop.minor, prop.multiProcessorCount,
           prop.totalGlobalMem / 1073741824.0, prop.l2CacheSize >> 20, rt, drv);
}

__global__ void s2h(int m, int n, float* as, int ldas, __half* ah, int ldah) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < m && j < n) ah[i + (long)j * ldah] = __float2half_rn(as[i + (long)j * ldas]);
}

__global__ void h2s(int m, int n, __half* ah, int ldah, float* as, int ldas) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < m && j < n) as[i + (long)j * ldas] = __half2float(ah[i + (long)j * ldah]);
}

__global__ void setEye(int m, int n, float* a, int lda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < m && j < n) a[i + (long)j * lda] = (i == j) ? 1.f : 0.f;
}

__global__ void clearTri(char uplo, int m, int n, float* a, int lda) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= m || j >= n) return;
    const bool kill = (uplo == 'l') ? (i > j) : (i < j);
    if (kill) a[i + (long)j * lda] = 0.f;
}

__global__ void deviceCopy(int m, int n, float* da, int lda, float* db, int ldb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < m && j < n) db[i + (long)j * ldb] = da[i + (long)j * lda];
}

__global__ void sSubstractAndSquare(int m, int n, float* dA, int lda, float* dB, int ldb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= m || j >= n) return;
    const float d = dA[i + (long)j * lda] - dB[i + (long)j * ldb];
    dB[i + (long)j * ldb] = d * d;
}
