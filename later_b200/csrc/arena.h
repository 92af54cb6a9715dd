// Stream-ordered workspace: one slab obtained with cudaMallocAsync on the context's stream,
// bump sub-allocation inside it, no per-call cudaMalloc.  (The reference has no workspace on the
// RGSQRF path - the caller passes work/hwork, reference test/test_qr.cu:73-78 - and its
// util/mem_pool.cu serves only the out-of-core code; this is the new facility SURVEY.md par.7
// step 1 asks for.)
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace lb {

class Arena {
public:
    Arena() = default;
    ~Arena() { release(); }
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;

    void bind(cudaStream_t stream) { stream_ = stream; }

    // Ensures capacity >= bytes.  Growing frees the old slab in stream order (work already
    // enqueued keeps using it safely) and bumps generation() so cached graphs are rebuilt.
    cudaError_t reserve(size_t bytes) {
        if (bytes <= capacity_) return cudaSuccess;
        if (base_) {
            cudaError_t e = cudaFreeAsync(base_, stream_);
            if (e != cudaSuccess) return e;
            base_ = nullptr;
            capacity_ = 0;
        }
        const size_t want = (bytes + ((size_t)1 << 21) - 1) & ~(((size_t)1 << 21) - 1);  // 2 MiB
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, want, stream_);
        if (e != cudaSuccess) return e;
        base_ = static_cast<uint8_t*>(p);
        capacity_ = want;
        ++generation_;
        return cudaSuccess;
    }

    void reset() { offset_ = 0; }

    // 256-byte aligned sub-allocation; nullptr when the slab is exhausted.
    void* alloc(size_t bytes) {
        const size_t start = (offset_ + 255) & ~(size_t)255;
        if (start + bytes > capacity_) return nullptr;
        offset_ = start + bytes;
        return base_ + start;
    }
    template <typename T>
    T* alloc_n(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }

    size_t capacity() const { return capacity_; }
    unsigned long generation() const { return generation_; }

    void release() {
        if (base_) {
            cudaFreeAsync(base_, stream_);
            base_ = nullptr;
            capacity_ = 0;
            offset_ = 0;
        }
    }

private:
    cudaStream_t stream_ = nullptr;
    uint8_t* base_ = nullptr;
    size_t capacity_ = 0;
    size_t offset_ = 0;
    unsigned long generation_ = 0;
};

}  // namespace lb
