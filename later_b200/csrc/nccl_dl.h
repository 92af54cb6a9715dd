// The handful of NCCL entry points this library uses, bound at run time (dlopen of libnccl.so.2), so
// that liblater_b200.so has no link-time dependency on NCCL and nccl.h is not needed to build.  In a
// process that has already loaded NCCL (PyTorch) the loader returns that same copy.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <condition_variable>
#include <cstddef>
#include <memory>
#include <mutex>
#include <string>

namespace lb {

typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7, kNcclDouble = 8;     // ncclFloat32, ncclFloat64
constexpr int kNcclSum = 0;
constexpr int kNcclUniqueIdBytes = 128;
struct NcclUniqueId { char internal[kNcclUniqueIdBytes]; };

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    // Process-wide instance, loaded on first use; nullptr (and *err filled) when NCCL cannot be loaded.
    static Nccl* get(std::string* err) {
        static Nccl inst;
        static std::once_flag once;
        std::call_once(once, [] { inst.load(); });
        if (!inst.handle || !inst.error.empty()) {
            if (err) *err = inst.error;
            return nullptr;
        }
        return &inst;
    }

private:
    void load() {
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!handle) { error = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto sym = [&](const char* name) {
            void* p = dlsym(handle, name);
            if (!p && error.empty()) error = std::string("libnccl.so.2 lacks ") + name;
            return p;
        };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    }
};

// The P contexts of ONE process that share a communicator set (later_b200_comm_init_all): NCCL wants the
// collectives of several devices driven from one process to be issued together, inside one group call
// by one thread.  Every context's host thread brings its request here; the last one to arrive issues
// all P of them and releases the others.
struct CommGroup {
    static constexpr int kMax = 16;
    std::mutex mu;
    std::condition_variable cv;
    int P = 0, arrived = 0;
    long generation = 0;
    bool failed = false;
    struct Req { void* buf; size_t count; int dtype; cudaStream_t stream; ncclComm_t comm; } req[kMax];

    // returns 0, an NCCL error code, or -1 when another rank gave up
    int allreduce(Nccl* nccl, int rank, void* buf, size_t count, int dtype, ncclComm_t comm, cudaStream_t stream) {
        std::unique_lock<std::mutex> lock(mu);
        if (failed) return -1;
        req[rank] = Req{buf, count, dtype, stream, comm};
        if (++arrived == P) {
            ncclResult_t r = nccl->GroupStart();
            for (int p = 0; p < P && r == 0; ++p)
                r = nccl->AllReduce(req[p].buf, req[p].buf, req[p].count, req[p].dtype, kNcclSum, req[p].comm, req[p].stream);
            const ncclResult_t r2 = nccl->GroupEnd();
            if (r == 0) r = r2;
            arrived = 0;
            ++generation;
            if (r != 0) failed = true;
            cv.notify_all();
            return r;
        }
        const long gen = generation;
        cv.wait(lock, [&] { return generation != gen || failed; });
        return failed ? -1 : 0;
    }
    void abort() {
        std::lock_guard<std::mutex> lock(mu);
        failed = true;
        cv.notify_all();
    }
    void reset() {
        std::lock_guard<std::mutex> lock(mu);
        failed = false;
        arrived = 0;
    }
};

}  // namespace lb
