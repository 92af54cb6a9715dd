// Per-device context behind the C ABI (include/later_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "arena.h"
#include "nccl_dl.h"
#include "panel.cuh"

namespace lb {
// Diagnostic knobs, read from the environment once per context (later_b200_create).
struct Options {
    PanelOpts panel;
    bool gram_cast = true;        // LB_GRAM_CAST = 0: always cast first, never in the Gram load path
    int update_variant = 0;       // LB_UPDATE_VARIANT (later_b200_gemm_update only)
    int ormqr_kchunk = 2048;      // LB_ORMQR_KCHUNK
    bool gram_2cta = true;        // LB_GRAM_2CTA = 0: never use the CTA-pair Gram kernel
    bool peer_allreduce = true;   // LB_PEER_ALLREDUCE = 0: NCCL for every all-reduce of the row-sharded path
    bool node_coop = true;        // LB_NODE_COOP = 0: launch the fused node kernel without the cooperative attribute
    int node_fuse = 128;          // LB_NODE_FUSE: largest half-width handled by the fused node kernel (0, 128, 256;
                                  // 256 works but measures 0.8 % slower on 16384^2)
};
constexpr int kSyncWords = 8;      // grid-barrier counters of the fused node kernel, behind the status words
constexpr int kGraphSlots = 4;    // cached executable graphs per entry point (LRU)

// NVLink peer-memory all-reduce of the small blocks the row-sharded factorisation exchanges
// (peer_comm.cu): every rank owns a slab that all its peers have mapped.
struct PeerComm {
    static constexpr int kMaxRanks = 16;
    void* slab = nullptr;                 // this rank's slab (cudaMalloc)
    void* peers[kMaxRanks] = {};          // every rank's slab in this rank's address space
    bool ipc[kMaxRanks] = {};             // opened with cudaIpcOpenMemHandle (to be closed)
    int nranks = 1, rank = 0;
    size_t half_bytes = 0;                // largest message
    bool ready = false;
    size_t slab_bytes() const;
    cudaError_t allocate(size_t max_message_bytes);
    void release();
    bool fits(size_t bytes) const;
    // in-place sum over the ranks of `count` doubles / floats, same bits on every rank
    // only_if (optional, device): same value on every rank; 0 = skip (consistently everywhere)
    cudaError_t allreduce(void* buf, size_t count, bool f64, cudaStream_t stream, const int* only_if = nullptr);
};
}  // namespace lb

struct later_b200_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool use_graph = true;
    std::string error;
    lb::Arena arena;
    lb::Options opts;
    long launches = 0;          // kernels launched by the most recent call
    // status words of the most recent factorisation (lb::PanelInfo): device copy the kernels write,
    // page-locked host copy filled by a D2H copy at the end of every factorisation
    int* d_info = nullptr;
    int* h_info = nullptr;

    // Layout of the most recent factorisation's workspace (for tsqr_apply and graph reuse).
    struct Plan {
        int m = 0, n = 0;
        float* A = nullptr; int lda = 0;
        float* R = nullptr; int ldr = 0;
        __half* Qh = nullptr; long ldh = 0;       // fp16 shadow of A/Q
        __half* R12h = nullptr;                   // fp16 R12 of the current node
        __half* Wh = nullptr;                     // fp16 W for tsqr_apply (n x n)
        float* part = nullptr; size_t part_floats = 0;
        void* panel_scratch = nullptr;
        unsigned long arena_gen = 0;
        bool dist = false;                        // row-sharded factorisation (all-reduces inside)
        float* stage = nullptr;                   // its contiguous R12 staging block (n/2 x n/2 fp32)
        // host buffers of later_b200_rgsqrf_host (null for the device entry points): the copies are
        // part of that call's graph, so they are part of its identity
        float* hA = nullptr; long hlda = 0;
        float* hR = nullptr; long hldr = 0;
        bool valid = false;
    } plan;

    // Cached executable graphs: [0] factorisation of a device matrix, [1] the host entry point
    // (same launches with the PCIe copies woven in as memcpy nodes on forked branches).  Each entry
    // point keeps the lb::kGraphSlots most recently used plans (shape + every pointer), so a caller
    // that alternates between a few buffers or shapes (double-buffered out-of-core drivers,
    // reference QR/later_oc_qr.cu:21; the QDWH iteration, EVD/later_qdwh_polar.cu:79) keeps replaying.
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        Plan plan;
        long launches = 0;
        bool seen = false;      // plan was launched directly once; capture on the next identical call
        unsigned long tick = 0; // last use (LRU)
    } graphs[3][lb::kGraphSlots];      // [2]: the row-sharded factorisation (NCCL all-reduces captured with it)
    unsigned long graph_tick = 0;
    long graph_replays = 0, graph_captures = 0;   // counters (tests)

    // pipelined host path: copy streams and a pool of fork/join events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> events;

    // device staging buffers for the *_host entry point
    float* dA = nullptr; size_t dA_bytes = 0;
    float* dR = nullptr; size_t dR_bytes = 0;
    // communicator of the row-sharded factorisation (later_b200_comm_init / _comm_init_all)
    lb::Nccl* nccl = nullptr;
    lb::ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    std::shared_ptr<lb::CommGroup> comm_group;   // set when the ranks are contexts of this process
    lb::PeerComm peer;                            // NVLink peer-memory path for the small all-reduces
    // scratch of the callers built on top of the factorisation (re-orthogonalisation, QDWH): lives
    // outside the arena, which every factorisation carves anew
    void* aux = nullptr; size_t aux_bytes = 0;
};

namespace lb {
int fail(later_b200_ctx* ctx, int code, const std::string& msg);
int cuda_fail(later_b200_ctx* ctx, cudaError_t e, const char* where);

// C = A * B, all fp32 column-major (A: M x K, B: K x N), fp32-faithful split-precision tcgen05
// products (ormqr.cu).  M and K multiples of 8; scratch: split_gemm_scratch_bytes(M, N, K).
size_t split_gemm_scratch_bytes(int M, int N, int K);
int split_gemm_nn(later_b200_ctx* ctx, int M, int N, int K, const float* A, long lda, const float* B, long ldb,
                  float* C, long ldc, void* scratch, long* launches);

// C2 -= X (Z^T C2), fp32 column-major, `rows` rows (Z, X: rows x h; C2: rows x nb), fp32-faithful (ormqr.cu).
size_t split_project_scratch_bytes(int rows, int h, int nb);
int split_project(later_b200_ctx* ctx, int rows, int h, int nb, const float* Z, long ldz, const float* X, long ldx,
                  float* C2, long ldc, void* scratch, long* launches);

// Makes the context's device current for the duration of a C-ABI call and restores the caller's.
class DeviceGuard {
public:
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
        // (always: in a thread that has made no runtime call yet, cudaGetDevice reports device 0 without
        // binding its context, and the driver-API tensor-map encoder then fails with "invalid argument")
        err_ = cudaSetDevice(device);
        if (prev_ == device) prev_ = -1;          // nothing to restore
    }
    ~DeviceGuard() { if (prev_ >= 0) cudaSetDevice(prev_); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
    cudaError_t error() const { return err_; }
private:
    int prev_ = -1;
    cudaError_t err_ = cudaSuccess;
};
}  // namespace lb
