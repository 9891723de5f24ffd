/* dev_api.h — entry points of the DEVELOPMENT library only (libb200unet3d_dev.so, built with -DB200_DEV by
 * build.build(dev=True)).  Not part of the product ABI (include/b200_unet3d.h). */
#ifndef B200_DEV_API_H
#define B200_DEV_API_H
#ifdef __cplusplus
extern "C" {
#endif
/* cycles for `iters` x 4 tcgen05.mma (M=128, N=n, K=16, SS mode) per CTA, operands cycling through `stages`
 * shared-memory slots; out_cycles[blocks] (int64) */
int b200_probe_mma(int n, int iters, int stages, long long* out_cycles, int blocks, void* stream);
/* same with UMMA M (64 or 128) and operand majorness (0: both K-major, 1: both MN-major) selectable */
int b200_probe_mma2(int m, int n, int mn_major, int iters, int stages, long long* out_cycles, int blocks,
                    void* stream);
/* CTA-pair (cta_group::2) primitives: d_out[pairs][256][n] fp32 = A[256][k] * B[n][k]^T (bf16, k contiguous) computed
 * by `pairs` clusters of two CTAs, the MMA chain repeated `iters` times; cycles[pairs] (int64) */
int b200_probe_pair(const void* a, const void* b, int n, int k, int iters, float* d_out, long long* cycles, int pairs,
                    void* stream);
/* kernel ablations for timing experiments (results are garbage, timings are not): igemm 1 = no TMA loads, 2 = no MMAs,
 * 3 = loads the MMAs do not wait for; dmarch 1 / 2 likewise; nopair = single-CTA igemm everywhere; stage_cap = ring depth */
int b200_dev_set_ablation(int igemm_ablate, int dmarch_ablate, int nopair, int stage_cap);
/* launcher variants under test (tools/bench_variants.py): which = 0 convT forward UMMA N, 1 convT forward staging tiles,
 * 4 = ablation of the first-layer marching kernels (tools/ablate_conv1_march.py): 1 no output stores / dy loads, 2 no input
 * loads, 3 no MMAs */
int b200_dev_set_variant(int which, int value);
#ifdef __cplusplus
}
#endif
#endif
