// Row-sharded tall-skinny QR over the GPUs of one node, behind the C ABI (SURVEY.md par.8 b / e: the
// reference has no multi-GPU code; the structure is the one its CAQR panel uses inside one GPU across
// 256-row blocks, reference QR/panel.cu:87-104, lifted to devices):
//
//   1. every device factors its own row block        A_p = Q_p R_p       (later_b200_rgsqrf, local)
//   2. the P small R_p (n x n) are exchanged                              (ONE ncclAllGather over NVLink)
//   3. every device factors the same stack [R_0; ...; R_{P-1}] = W R      (redundantly: same inputs, same
//      code, same bits, so R and W need no broadcast)
//   4. Q_p <- Q_p W_p                                                      (tcgen05 GEMM, local)
//
// One host thread drives all devices (one stream each); only step 2 communicates.  NCCL is loaded at
// run time (dlopen of libnccl.so.2), so the library itself has no link-time dependency on it.
// The multi-process variant of the same algorithm (one rank per GPU, torch.distributed) is
// later_b200/tsqr.py; both call the same device entry points.
#include "../../include/later_b200.h"

#include <string>
#include <thread>
#include <vector>

#include "context.h"
#include "nccl_dl.h"

using lb::Nccl;
using lb::ncclComm_t;
using lb::ncclResult_t;
using lb::kNcclFloat;

struct later_b200_mgpu {
    int P = 0;
    std::vector<int> devices;
    std::vector<cudaStream_t> streams;
    std::vector<later_b200_ctx*> main_ctx, stack_ctx;
    std::vector<ncclComm_t> comms;
    // per device: R_p (n x n), the gathered [q][n x n] buffer and the stack (P n x n, ld = P n)
    std::vector<float*> Rloc, gathered, stack;
    int n_alloc = 0;
    Nccl* nccl = nullptr;
    std::string error;
};

namespace {

int mfail(later_b200_mgpu* g, int code, const std::string& msg) {
    if (g) g->error = msg;
    return code;
}

int ensure_buffers(later_b200_mgpu* g, int n) {
    if (n <= g->n_alloc) return 0;
    const size_t nn = (size_t)n * n * sizeof(float);
    for (int p = 0; p < g->P; ++p) {
        lb::DeviceGuard guard(g->devices[p]);
        cudaStreamSynchronize(g->streams[p]);
        for (float** buf : {&g->Rloc[p], &g->gathered[p], &g->stack[p]}) {
            if (*buf) cudaFree(*buf);
            *buf = nullptr;
        }
        cudaError_t e = cudaMalloc(&g->Rloc[p], nn);
        if (e == cudaSuccess) e = cudaMalloc(&g->gathered[p], nn * g->P);
        if (e == cudaSuccess) e = cudaMalloc(&g->stack[p], nn * g->P);
        if (e != cudaSuccess) return mfail(g, LATER_B200_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    g->n_alloc = n;
    return 0;
}

}  // namespace

extern "C" {

int later_b200_mgpu_create(later_b200_mgpu** out, int P, const int* devices) {
    if (!out || P < 1 || !devices) return LATER_B200_EINVAL;
    *out = nullptr;
    later_b200_mgpu* g = new (std::nothrow) later_b200_mgpu();
    if (!g) return LATER_B200_ENOMEM;
    g->P = P;
    g->devices.assign(devices, devices + P);
    g->streams.assign(P, nullptr);
    g->main_ctx.assign(P, nullptr);
    g->stack_ctx.assign(P, nullptr);
    g->Rloc.assign(P, nullptr);
    g->gathered.assign(P, nullptr);
    g->stack.assign(P, nullptr);
    int rc = 0;
    for (int p = 0; p < P && rc == 0; ++p) {
        lb::DeviceGuard guard(devices[p]);
        if (guard.error() != cudaSuccess || cudaStreamCreateWithFlags(&g->streams[p], cudaStreamNonBlocking) != cudaSuccess)
            rc = LATER_B200_ENODEV;
        if (rc == 0) rc = later_b200_create(&g->main_ctx[p], devices[p], g->streams[p]);
        if (rc == 0) rc = later_b200_create(&g->stack_ctx[p], devices[p], g->streams[p]);
    }
    // communicators: one set for the row-sharded recursion (inside the main contexts), one for the
    // all-gather of the TSQR variant
    if (rc == 0 && P > 1 && (rc = later_b200_comm_init_all(g->main_ctx.data(), P)) != 0)
        g->error = later_b200_last_error(g->main_ctx[0]);
    if (rc == 0 && P > 1 && (rc = later_b200_peer_init_all(g->main_ctx.data(), P, (size_t)4 << 20)) != 0)
        g->error = later_b200_last_error(g->main_ctx[0]);
    if (rc == 0 && P > 1) {
        g->nccl = Nccl::get(&g->error);
        if (!g->nccl) rc = LATER_B200_ESTATE;
        if (rc == 0) {
            g->comms.assign(P, nullptr);
            ncclResult_t r = g->nccl->CommInitAll(g->comms.data(), P, devices);
            if (r != 0) {
                g->comms.clear();
                rc = mfail(g, LATER_B200_ESTATE, std::string("ncclCommInitAll: ") + g->nccl->GetErrorString(r));
            }
        }
    }
    if (rc != 0) {
        fprintf(stderr, "later_b200_mgpu_create: %s (rc=%d)\n", g->error.c_str(), rc);
        later_b200_mgpu_destroy(g);
        return rc;
    }
    *out = g;
    return 0;
}

int later_b200_mgpu_destroy(later_b200_mgpu* g) {
    if (!g) return LATER_B200_EINVAL;
    for (int p = 0; p < g->P; ++p) {
        lb::DeviceGuard guard(g->devices[p]);
        if (g->streams[p]) cudaStreamSynchronize(g->streams[p]);
        if (p < (int)g->comms.size() && g->comms[p]) g->nccl->CommDestroy(g->comms[p]);
        if (g->main_ctx[p]) later_b200_destroy(g->main_ctx[p]);
        if (g->stack_ctx[p]) later_b200_destroy(g->stack_ctx[p]);
        for (float* buf : {g->Rloc[p], g->gathered[p], g->stack[p]})
            if (buf) cudaFree(buf);
        if (g->streams[p]) cudaStreamDestroy(g->streams[p]);
    }
    delete g;
    return 0;
}

const char* later_b200_mgpu_last_error(const later_b200_mgpu* g) { return g ? g->error.c_str() : "null handle"; }

int later_b200_mgpu_sync(later_b200_mgpu* g) {
    if (!g) return LATER_B200_EINVAL;
    for (int p = 0; p < g->P; ++p) {
        lb::DeviceGuard guard(g->devices[p]);
        cudaError_t e = cudaStreamSynchronize(g->streams[p]);
        if (e != cudaSuccess) return mfail(g, (int)e, std::string("sync: ") + cudaGetErrorString(e));
    }
    return 0;
}

int later_b200_tsqr_mgpu(later_b200_mgpu* g, int m_local, int n, float* const* A, int lda, float* const* R, int ldr) {
    if (!g || !A || !R) return LATER_B200_EINVAL;
    const int P = g->P;
    auto ctx_fail = [&](later_b200_ctx* c, int rc) { return mfail(g, rc, later_b200_last_error(c)); };
    if (P == 1) {
        int rc = later_b200_rgsqrf(g->main_ctx[0], m_local, n, A[0], lda, R[0], ldr);
        return rc ? ctx_fail(g->main_ctx[0], rc) : 0;
    }
    if ((long)P * n > 2147483647L / 4) return mfail(g, LATER_B200_EINVAL, "stack too tall");
    int rc = ensure_buffers(g, n);
    if (rc) return rc;
    const size_t nn = (size_t)n * n;
    // 1. local factorisations (asynchronous: every device gets its work before anyone waits)
    for (int p = 0; p < P; ++p)
        if ((rc = later_b200_rgsqrf(g->main_ctx[p], m_local, n, A[p], lda, g->Rloc[p], n)) != 0)
            return ctx_fail(g->main_ctx[p], rc);
    // 2. the exchange: one all-gather of the column-major storage of every R_p
    ncclResult_t r = g->nccl->GroupStart();
    for (int p = 0; p < P && r == 0; ++p)
        r = g->nccl->AllGather(g->Rloc[p], g->gathered[p], nn, kNcclFloat, g->comms[p], g->streams[p]);
    if (r == 0) r = g->nccl->GroupEnd(); else g->nccl->GroupEnd();
    if (r != 0) return mfail(g, LATER_B200_ESTATE, std::string("ncclAllGather: ") + g->nccl->GetErrorString(r));
    for (int p = 0; p < P; ++p) {
        lb::DeviceGuard guard(g->devices[p]);
        // block q of the gathered buffer -> rows [q n, (q + 1) n) of the stack (canonical order: device 0 on top)
        for (int q = 0; q < P; ++q) {
            cudaError_t e = cudaMemcpy2DAsync(g->stack[p] + (size_t)q * n, (size_t)P * n * sizeof(float),
                                              g->gathered[p] + q * nn, (size_t)n * sizeof(float),
                                              (size_t)n * sizeof(float), n, cudaMemcpyDeviceToDevice, g->streams[p]);
            if (e != cudaSuccess) return mfail(g, (int)e, std::string("stack copy: ") + cudaGetErrorString(e));
        }
        // 3. redundant QR of the stack; 4. back-multiplication with this device's n x n block of its Q
        if ((rc = later_b200_rgsqrf(g->stack_ctx[p], P * n, n, g->stack[p], P * n, R[p], ldr)) != 0)
            return ctx_fail(g->stack_ctx[p], rc);
        if ((rc = later_b200_tsqr_apply(g->main_ctx[p], m_local, n, A[p], lda, g->stack[p] + (size_t)p * n, P * n)) != 0)
            return ctx_fail(g->main_ctx[p], rc);
    }
    return 0;
}

int later_b200_rgsqrf_mgpu(later_b200_mgpu* g, int m_local, int n, float* const* A, int lda, float* const* R, int ldr) {
    if (!g || !A || !R) return LATER_B200_EINVAL;
    const int P = g->P;
    if (P == 1) {
        int rc = later_b200_rgsqrf(g->main_ctx[0], m_local, n, A[0], lda, R[0], ldr);
        return rc ? mfail(g, rc, later_b200_last_error(g->main_ctx[0])) : 0;
    }
    // one host thread per device, each enqueuing its own launch sequence; at every all-reduce the threads
    // meet and ONE of them issues the P requests inside one NCCL group call (lb::CommGroup)
    std::vector<int> rcs(P, 0);
    std::vector<std::thread> threads;
    if (g->main_ctx[0]->comm_group) g->main_ctx[0]->comm_group->reset();
    for (int p = 0; p < P; ++p)
        threads.emplace_back([&, p] { rcs[p] = later_b200_rgsqrf_dist(g->main_ctx[p], m_local, n, A[p], lda, R[p], ldr); });
    for (auto& t : threads) t.join();
    for (int p = 0; p < P; ++p)
        if (rcs[p] != 0) return mfail(g, rcs[p], later_b200_last_error(g->main_ctx[p]));
    return 0;
}

}  // extern "C"
