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

// ---------------------------------------------------------------------------------------------- CTA-pair probe
// D[256][n] = A[256][k] * B[n][k]^T with one tcgen05.mma.cta_group::2 chain per k-block, operands fetched by each CTA
// of the pair with the pair form of the TMA load (bytes counted on the leader's barrier).  Functional check of the
// pair primitives (allocation, operand split, D placement, multicast commit) and cycles per M = 256 MMA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
    pair_probe_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap b_map, int n,
                      int kblocks, int iters, float* d_out, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t rank = cluster_ctarank();
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t bhalf = (uint32_t)(n / 2) * 128;
    const uint32_t smem_a = base, smem_b = base + kblocks * 16384, bar_full = smem_b + kblocks * bhalf,
                   bar_done = bar_full + 8, tptr = bar_full + 16;
    if (warp == 1 && lane == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tptr, 256);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tptr - base));
    if (warp == 0) {
        if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(bar_full, 2u * kblocks * (16384u + bhalf));
            for (int kb = 0; kb < kblocks; ++kb) {
                tma_load_3d_pair(smem_a + kb * 16384, &a_map, bar_full, kb * 64, (int)rank * 128, 0);
                tma_load_3d_pair(smem_b + kb * bhalf, &b_map, bar_full, kb * 64, (int)rank * (n / 2), 0);
            }
        }
        __syncwarp();
    } else if (warp == 1 && rank == 0) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(256, n, 0, 0);
        const uint64_t a0 = make_smem_desc_sw128(smem_a, 0, 1024), b0 = make_smem_desc_sw128(smem_b, 0, 1024);
        if (elect_one()) {
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it)
                for (int kb = 0; kb < kblocks; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_pair(tmem, a0 + kb * (16384 >> 4) + 2 * k, b0 + kb * (bhalf >> 4) + 2 * k, idesc,
                                      (it | kb | k) != 0);
            umma_commit_pair(bar_done, 3u);
            while (!mbar_try_wait(bar_done, 0)) {}
            cycles[blockIdx.x >> 1] = clock64() - t0;
        }
        __syncwarp();
    } else if (warp >= 4) {
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const int q = warp - 4, row = q * 32 + lane;
        float* dst = d_out + ((long long)(blockIdx.x >> 1) * 256 + rank * 128 + row) * n;
        for (int c = 0; c < n; c += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + (static_cast<uint32_t>(q * 32) << 16) + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[c + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem, 256);
    }
}

cudaError_t launch_pair_probe(const CUtensorMap& a_map, const CUtensorMap& b_map, int n, int kblocks, int iters,
                              float* d_out, long long* cycles, int pairs, cudaStream_t s) {
    const size_t smem = 1024 + (size_t)kblocks * (16384 + (n / 2) * 128) + 64;
    cudaError_t e = cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    pair_probe_kernel<<<2 * pairs, 256, smem, s>>>(a_map, b_map, n, kblocks, iters, d_out, cycles);
    return cudaGetLastError();
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
