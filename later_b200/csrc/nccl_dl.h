// The handful of NCCL entry points this library uses, bound at run time (dlopen of libnccl.so.2), so
// that liblater_b200.so has no link-time dependency on NCCL and nccl.h is not needed to build.  In a
// process that has already loaded NCCL (PyTorch) the loader returns that same copy.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>
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

}  // namespace lb
