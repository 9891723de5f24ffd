// Micro-probe (dev tool, not on the hot path): cycles per tcgen05.mma (kind::f16, M=128, SS mode) as a function of N,
// operands resident in shared memory, no TMA traffic.  Answers how much of the shared-memory operand fetch of a
// narrow-N MMA is exposed.  Entry point: b200_probe_mma(n, iters, reuse_a, out_cycles[gridDim]).
#include <cuda_runtime.h>
#include <stdint.h>
#include "ptx.cuh"

namespace b200 {

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int iters, int stages, long long* out, int m,
                                                          int mn_major) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // A stages: 16 KB each; B stages: n * 128 B each; barrier + tmem ptr at the end
    const uint32_t smem_a = base, smem_b = base + stages * 16384, bar = smem_b + stages * n * 128, tptr = bar + 16;
    for (uint32_t i = threadIdx.x; i < (bar - base) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 0) {
        tmem_alloc(tptr, 512);
        tmem_relinquish();
    }
    if (threadIdx.x == 32) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tptr - base));
    if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(m, n, mn_major, mn_major);
        const uint64_t a0 = make_smem_desc_sw128(smem_a, mn_major ? 8192 : 0, 1024);
        const uint64_t b0 = make_smem_desc_sw128(smem_b, mn_major ? 8192 : 0, 1024);
        const uint32_t kinc = mn_major ? 128u : 2u;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const int st = it % stages;
                const uint64_t a = a0 + st * (16384 >> 4), b = b0 + st * ((n * 128) >> 4);
                umma_f16(tmem, a, b, idesc, it > 0);
                umma_f16(tmem, a + kinc, b + kinc, idesc, 1u);
                umma_f16(tmem, a + 2 * kinc, b + 2 * kinc, idesc, 1u);
                umma_f16(tmem, a + 3 * kinc, b + 3 * kinc, idesc, 1u);
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        if (elect_one()) {
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace b200

extern "C" int b200_probe_mma2(int m, int n, int mn_major, int iters, int stages, long long* out_cycles, int blocks,
                               void* stream);
extern "C" int b200_probe_mma(int n, int iters, int stages, long long* out_cycles, int blocks, void* stream) {
    return b200_probe_mma2(128, n, 0, iters, stages, out_cycles, blocks, stream);
}
extern "C" int b200_probe_mma2(int m, int n, int mn_major, int iters, int stages, long long* out_cycles, int blocks,
                               void* stream) {
    using namespace b200;
    const size_t smem = 1024 + (size_t)stages * (16384 + n * 128) + 64;
    if (smem > 227 * 1024 || n % 16 || n < 16 || n > 256) return 1;
    if (cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 3;
    mma_probe_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(n, iters, stages, out_cycles, m, mn_major);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
