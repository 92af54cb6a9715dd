// prints per-kernel resource usage as the driver sees it (run on the GPU box)
#include <cstdio>
#include <cuda_runtime.h>
int main() { cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("regsPerBlock %d regsPerSM %d smemPerBlockOptin %zu maxThreadsPerSM %d\n", p.regsPerBlock, p.regsPerMultiprocessor, p.sharedMemPerBlockOptin, p.maxThreadsPerMultiProcessor); return 0; }
