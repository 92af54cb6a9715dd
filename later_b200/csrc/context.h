// Per-device context behind the C ABI (include/later_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "arena.h"

struct later_b200_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool use_graph = true;
    std::string error;
    lb::Arena arena;
    long launches = 0;          // kernels launched by the most recent call

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
        // host buffers of later_b200_rgsqrf_host (null for the device entry points): the copies are
        // part of that call's graph, so they are part of its identity
        float* hA = nullptr; long hlda = 0;
        float* hR = nullptr; long hldr = 0;
        bool valid = false;
    } plan;

    // Cached executable graphs: [0] factorisation of a device matrix, [1] the host entry point
    // (same launches with the PCIe copies woven in as memcpy nodes on forked branches).
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        Plan plan;
        long launches = 0;
        bool seen = false;      // plan was launched directly once; capture on the next identical call
    } graphs[2];

    // pipelined host path: copy streams and a pool of fork/join events
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> events;

    // device staging buffers for the *_host entry point
    float* dA = nullptr; size_t dA_bytes = 0;
    float* dR = nullptr; size_t dR_bytes = 0;
};

namespace lb {
int fail(later_b200_ctx* ctx, int code, const std::string& msg);
int cuda_fail(later_b200_ctx* ctx, cudaError_t e, const char* where);
}  // namespace lb
