// Kernel launch with programmatic dependent launch (PDL) enabled: the grid may start while its
// predecessor in the stream is still running; every kernel launched this way executes
// griddepcontrol.wait (pdl_wait) before it touches memory written by predecessors, and
// griddepcontrol.launch_dependents (pdl_trigger) as early as possible, so launch latency and
// prologues (barrier init, TMEM allocation, tensor-map prefetch) overlap the previous kernel's tail.
#pragma once
#include <cuda_runtime.h>

namespace lb {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                       cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace lb
