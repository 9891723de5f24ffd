// Launch declarations of the HBM-bound kernels (bandwidth.cu), called from api.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

struct View {
    __nv_bfloat16* p;
    long long n, d, h, w, c, ld;
    __host__ __device__ long long voxels() const { return n * d * h * w; }
};

// 32-bit divide by a runtime constant (numerator < 2^31)
struct FastDiv {
    uint32_t div, mul, shr;
    __host__ FastDiv() : div(1), mul(0), shr(0) {}
    __host__ explicit FastDiv(uint32_t d) : div(d), mul(0), shr(0) {
        if (d != 1) {
            uint32_t lg = 0;
            while ((1ull << lg) < d) ++lg;
            const uint32_t pw = 31 + lg;
            mul = (uint32_t)(((1ull << pw) + d - 1) / d);
            shr = pw - 32;
        }
    }
    __device__ __forceinline__ uint32_t quot(uint32_t x) const { return div != 1 ? (__umulhi(x, mul) >> shr) : x; }
};

constexpr int kBwdMaxBlocks = 148 * 4;  // rows of the BatchNorm-backward partial buffer
constexpr int kLossMaxBlocks = 1024;

cudaError_t launch_pack_input(const float* x, long long n, long long c, long long d, long long h, long long w,
                              View out, cudaStream_t s);
cudaError_t launch_im2col_input(const float* x, long long n, long long c, long long d, long long h, long long w,
                                View out, cudaStream_t s);
cudaError_t launch_pack_rows(const float* w, int rows, int k, int kpad, __nv_bfloat16* out, cudaStream_t s);
cudaError_t launch_transpose_taps(const __nv_bfloat16* src, int taps, int rows, int cols, __nv_bfloat16* dst,
                                  cudaStream_t s);
cudaError_t launch_pack_conv1_slices(const float* w, int cout, int cin, __nv_bfloat16* out, cudaStream_t s);
cudaError_t launch_pack_conv_weight(const float* w, int cout, int cin, int cin_pad, __nv_bfloat16* wf,
                                    cudaStream_t s);
cudaError_t launch_pack_convt_weight(const float* w, const float* bias, int cin, int cout, __nv_bfloat16* wf,
                                     __nv_bfloat16* wd, float* bias8, cudaStream_t s);
cudaError_t launch_bn_finalize(const float* partial, long long m_tiles, long long count, int c, const float* gamma,
                               const float* beta, float eps, float momentum, float* rm, float* rv, long long* nbt,
                               float* mean, float* rstd, float* scale, float* shift, cudaStream_t s);
cudaError_t launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                                const float* cbias, float eps, int c, float* scale, float* shift, cudaStream_t s);
cudaError_t launch_bn_apply_relu(View y, const float* scale, const float* shift, View out, int sms, cudaStream_t s);
cudaError_t launch_bn_bwd_reduce(View dout, View y, const float* scale, const float* shift, const float* mean,
                                 const float* rstd, float* partial, int* nblk, cudaStream_t s);
cudaError_t launch_bn_bwd_finalize(const float* partial, int nblk, int c, long long count, float* dgamma,
                                   float* dbeta, float* coef, cudaStream_t s);
cudaError_t launch_bn_bwd_apply(View dout, View y, const float* scale, const float* shift, const float* mean,
                                const float* rstd, const float* coef, View dy, float* dbias, cudaStream_t s);
cudaError_t launch_bn_apply_relu_pool(View y, const float* scale, const float* shift, View out, View pooled, int sms,
                                      cudaStream_t s);
cudaError_t launch_bn_bwd_head(bool apply, const float* dlogits, const float* w, int ncls, View y, const float* scale,
                               const float* shift, const float* mean, const float* rstd, const float* coef,
                               float* partial, int* nblk, View dy, float* dbias, float* dw, float* db, cudaStream_t s);
cudaError_t launch_maxpool_fwd(View x, View y, int sms, cudaStream_t s);
cudaError_t launch_maxpool_bwd(View x, View dy, const View* dskip, View dx, int sms, cudaStream_t s);
cudaError_t launch_head_fwd(View x, const float* w, const float* b, int ncls, float* logits, float* probs,
                            cudaStream_t s);
cudaError_t launch_head_bwd(View x, const float* w, int ncls, const float* dlogits, View dx, float* dw, float* db,
                            cudaStream_t s);
cudaError_t launch_loss_fwd(const float* z, const float* t, long long n, float bce_w, float dice_w, float smooth,
                            float* ws, float* sums, float* loss, int sms, cudaStream_t s);
cudaError_t launch_loss_bwd(const float* z, const float* t, long long n, float bce_w, float dice_w, float smooth,
                            const float* sums, const float* gout, float* dz, int sms, cudaStream_t s);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double b1, double b2,
                        double eps, double wd, long long step, double gscale, const float* found_inf,
                        __nv_bfloat16* shadow, const float* dyn, int sms, cudaStream_t s);
cudaError_t launch_cast_bf16(const float* x, long long n, __nv_bfloat16* out, int sms, cudaStream_t s);
cudaError_t launch_sumsq(const float* x, long long n, float* out, int sms, cudaStream_t s);
cudaError_t launch_fill_zero(View v, int sms, cudaStream_t s);
cudaError_t launch_channel_sum(View v, float* out, const int* box /* d0,h0,w0,bd,bh,bw or null */, cudaStream_t s);
cudaError_t launch_unpack_act(View v, float* out, cudaStream_t s);
cudaError_t launch_resample3d(const float* in, long long nvol, int di, int hi, int wi, float* out, int dout, int ho,
                              int wo, int nearest, int binarize, int sms, cudaStream_t s);
cudaError_t launch_minmax_normalize(float* x, long long nvol, long long per, uint32_t* keys, int sms, cudaStream_t s);
cudaError_t launch_seg_counts(const float* score, const float* label, long long nsmp, long long per, float threshold,
                              unsigned long long* counts, int sms, cudaStream_t s);

int splitk_finalize_blocks(long long nvox, long long c);
cudaError_t launch_splitk_finalize(const float* ws, int splits, View y, int mode, const float* v0, const float* v1, float* partial,
                                   cudaStream_t s);
cudaError_t launch_window_gather(const float* x, long long c, long long d, long long h, long long w, const int* org,
                                 int nwin, int wd, int wh, int ww, float* out, int sms, cudaStream_t s);
cudaError_t launch_window_accumulate(const float* lg, const int* org, int nwin, long long k, int wd, int wh, int ww,
                                     float* acc, long long d, long long h, long long w, int v_lo, int v_cnt, int sms,
                                     cudaStream_t s);
cudaError_t launch_window_finalize(float* acc, const int* cover, long long nk, long long d, long long h, long long w,
                                   float threshold, float* probs, float* mask, int sms, cudaStream_t s);

}  // namespace b200
