// Kernel launches with programmatic dependent launch (PDL).
//
// A training step is ~180 dependent launches on one stream, most of them replayed from a CUDA graph; between two fully
// serialised kernels the GPU idles for the completion flush of the first plus the launch latency of the second.  Every
// kernel of this library therefore (1) is launched with cudaLaunchAttributeProgrammaticStreamSerialization, which lets
// its CTAs become resident as soon as every CTA of the stream predecessor has exited (no kernel here triggers
// earlier: a waiting GEMM CTA would take registers away from a bandwidth-bound predecessor), and (2) executes
// `griddepcontrol.wait` — completion and visibility of the predecessor — before its first global-memory access; what a
// kernel does before that (barrier init, TMEM allocation, tensor-map prefetch) overlaps the predecessor's tail.
// EVERY thread of EVERY kernel launched through launch_k must execute pdl_wait(), also CTAs that have no work: a grid
// that finishes without waiting would let ITS successor run ahead of the predecessor (ordering is transitive only
// through the waits).  Under stream capture the attribute becomes a programmatic edge of the graph — measured 1 % slower
// to replay than plain edges, so graph.GraphedTrainStep switches it off while it records (tools/ab_pdl.py).
// b200_set_pdl(0) restores plain launches (the attribute is dropped; the wait is then a no-op).
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace b200 {

extern int g_pdl;   // api.cu

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
    cudaLaunchConfig_t cfg;
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace b200
