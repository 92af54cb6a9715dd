// All-reduce of the small blocks the row-sharded factorisation exchanges (a panel's 80 KB Gram matrix,
// an R12 block of at most a few MB), written for NVLink peer memory instead of calling NCCL: at 8 GPUs an
// NCCL all-reduce of 80 KB costs ~33 us of latency, and the factorisation of 1048576 x 1024 needs 23 of
// them behind one another (0.75 ms of a 2.76 ms step).
//
// Every rank owns a "slab" that its peers have mapped (cudaIpcOpenMemHandle across processes, peer
// access inside one process).  One launch, no host involvement, no grid-wide barrier: CTA c owns slice c
// of the message and
//   1. copies its slice of the local data into the local slab (double-buffered by the parity of a
//      device-side sequence number, so a slow peer may still be reading the previous message),
//   2. fences, then raises flag[c][my rank] = sequence IN EVERY PEER'S slab (remote stores),
//   3. waits until its own flags flag[c][p] have reached the sequence for every p (local polling),
//   4. reads slice c from every rank's slab (remote loads) and adds them up IN RANK ORDER - every rank
//      performs the same additions in the same order, so the result is bit-identical everywhere.
// Reuse of a slab half two messages later is safe: nobody passes step 3 of message s + 1 before every
// rank has finished step 4 of message s (stream order on each rank).
#include "../../include/later_b200.h"

#include <algorithm>
#include <cstring>
#include <vector>

#include "context.h"
#include "launch.cuh"

namespace lb {

namespace {

constexpr int kPeerCtas = 32;                 // slices per message
constexpr int kPeerThreads = 512;

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// remote data: never from a stale L1 line (the slab halves are reused every second message)
template <typename T>
__device__ __forceinline__ T ld_relaxed_sys(const T* p);
template <>
__device__ __forceinline__ double ld_relaxed_sys<double>(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
template <>
__device__ __forceinline__ float ld_relaxed_sys<float>(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

struct PeerTable {
    uint8_t* slab[PeerComm::kMaxRanks];       // every rank's slab as mapped into this rank's address space
    int nranks, rank;
    size_t half_bytes;                        // size of one data half
};

// slab layout: [data half 0][data half 1][flags: kPeerCtas x kMaxRanks ints][seq: kPeerCtas ints]
__host__ __device__ inline size_t flags_offset(size_t half_bytes) { return 2 * half_bytes; }
__host__ __device__ inline size_t seq_offset(size_t half_bytes) {
    return 2 * half_bytes + sizeof(int) * kPeerCtas * PeerComm::kMaxRanks;
}

template <typename T>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(T* __restrict__ data, size_t count, PeerTable t, const int* __restrict__ only_if) {
    __shared__ int s_seq;
    const int c = blockIdx.x;
    uint8_t* mine = t.slab[t.rank];
    int* my_flags = reinterpret_cast<int*>(mine + flags_offset(t.half_bytes)) + c * PeerComm::kMaxRanks;
    int* my_seq = reinterpret_cast<int*>(mine + seq_offset(t.half_bytes)) + c;
    pdl_trigger();
    pdl_wait();                               // `data` comes from the preceding kernel
    if (only_if && *only_if == 0) return;     // (same decision on every rank: the counters stay in step)
    if (threadIdx.x == 0) s_seq = *my_seq + 1;   // per-CTA message counter, identical on every rank
    __syncthreads();
    const int seq = s_seq;
    const size_t per = (count + gridDim.x - 1) / gridDim.x;
    const size_t lo = min(count, (size_t)c * per), hi = min(count, lo + per);
    T* half = reinterpret_cast<T*>(mine + (size_t)(seq & 1) * t.half_bytes);
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) half[i] = data[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < t.nranks) {             // raise my flag for this slice in every rank's slab
        int* f = reinterpret_cast<int*>(t.slab[threadIdx.x] + flags_offset(t.half_bytes)) + c * PeerComm::kMaxRanks + t.rank;
        st_release_sys(f, seq);
    }
    if (threadIdx.x < t.nranks) {             // ... and wait for everybody else's
        unsigned long long spins = 0;
        while (ld_acquire_sys(my_flags + threadIdx.x) < seq) {
            if (++spins > (1ull << 31)) __trap();     // a rank never arrived: fail, do not hang the GPU for ever
        }
    }
    __syncthreads();
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        T s = 0;
        for (int p = 0; p < t.nranks; ++p)
            s += ld_relaxed_sys(reinterpret_cast<const T*>(t.slab[p] + (size_t)(seq & 1) * t.half_bytes) + i);
        data[i] = s;
    }
    if (threadIdx.x == 0) *my_seq = seq;
}

}  // namespace

size_t PeerComm::slab_bytes() const { return seq_offset(half_bytes) + sizeof(int) * kPeerCtas; }

cudaError_t PeerComm::allocate(size_t max_message_bytes) {
    half_bytes = (max_message_bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&slab, slab_bytes());
    if (e != cudaSuccess) return e;
    return cudaMemset(slab, 0, slab_bytes());
}

void PeerComm::release() {
    for (int p = 0; p < nranks; ++p)
        if (ipc[p] && peers[p] && p != rank) cudaIpcCloseMemHandle(peers[p]);
    if (slab) cudaFree(slab);
    slab = nullptr;
    ready = false;
}

bool PeerComm::fits(size_t bytes) const { return ready && bytes <= half_bytes; }

cudaError_t PeerComm::allreduce(void* buf, size_t count, bool f64, cudaStream_t stream, const int* only_if) {
    PeerTable t{};
    for (int p = 0; p < nranks; ++p) t.slab[p] = static_cast<uint8_t*>(peers[p]);
    t.nranks = nranks; t.rank = rank; t.half_bytes = half_bytes;
    cudaError_t e = f64 ? launch_pdl(peer_allreduce_kernel<double>, dim3(kPeerCtas), dim3(kPeerThreads), 0, stream,
                                     static_cast<double*>(buf), count, t, only_if)
                        : launch_pdl(peer_allreduce_kernel<float>, dim3(kPeerCtas), dim3(kPeerThreads), 0, stream,
                                     static_cast<float*>(buf), count, t, only_if);
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace lb

using namespace lb;

extern "C" {

int later_b200_peer_export(later_b200_ctx* ctx, size_t max_message_bytes, void* handle64) {
    if (!ctx || !handle64) return LATER_B200_EINVAL;
    DeviceGuard guard(ctx->device);
    if (ctx->peer.slab) return fail(ctx, LATER_B200_ESTATE, "peer slab already allocated");
    cudaError_t e = ctx->peer.allocate(max_message_bytes);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "peer slab");
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, ctx->peer.slab)) != cudaSuccess) return cuda_fail(ctx, e, "cudaIpcGetMemHandle");
    static_assert(sizeof(h) == LATER_B200_PEER_HANDLE_BYTES, "IPC handle size");
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

int later_b200_peer_import(later_b200_ctx* ctx, int nranks, int rank, const void* handles) {
    if (!ctx || !handles || nranks < 1 || nranks > PeerComm::kMaxRanks || rank < 0 || rank >= nranks) return LATER_B200_EINVAL;
    if (!ctx->peer.slab) return fail(ctx, LATER_B200_ESTATE, "later_b200_peer_export first");
    DeviceGuard guard(ctx->device);
    ctx->peer.nranks = nranks;
    ctx->peer.rank = rank;
    for (int p = 0; p < nranks; ++p) {
        if (p == rank) { ctx->peer.peers[p] = ctx->peer.slab; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)p * sizeof(h), sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&ctx->peer.peers[p], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaIpcOpenMemHandle");
        ctx->peer.ipc[p] = true;
    }
    ctx->peer.ready = true;
    return 0;
}

// The ranks are contexts of this process: plain peer access.
int later_b200_peer_init_all(later_b200_ctx* const* ctxs, int nranks, size_t max_message_bytes) {
    if (!ctxs || nranks < 1 || nranks > PeerComm::kMaxRanks) return LATER_B200_EINVAL;
    for (int p = 0; p < nranks; ++p) {
        if (!ctxs[p] || ctxs[p]->peer.slab) return LATER_B200_EINVAL;
        DeviceGuard guard(ctxs[p]->device);
        for (int q = 0; q < nranks; ++q)
            if (q != p) {
                cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(ctxs[p], e, "peer access");
                (void)cudaGetLastError();
            }
        cudaError_t e = ctxs[p]->peer.allocate(max_message_bytes);
        if (e != cudaSuccess) return cuda_fail(ctxs[p], e, "peer slab");
    }
    for (int p = 0; p < nranks; ++p) {
        ctxs[p]->peer.nranks = nranks;
        ctxs[p]->peer.rank = p;
        for (int q = 0; q < nranks; ++q) ctxs[p]->peer.peers[q] = ctxs[q]->peer.slab;
        ctxs[p]->peer.ready = true;
    }
    return 0;
}

}  // extern "C"
