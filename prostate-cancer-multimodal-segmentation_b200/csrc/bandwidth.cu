// HBM-bound kernels of the 3D U-Net hot path (sm_100a): BatchNorm finalize / apply / backward, ReLU, MaxPool3d,
// 1x1x1 head, sigmoid-BCE+Dice loss, fused Adam, layout packing.  All of them stream NDHWC bf16 activations with
// 16-byte vector accesses (8 channels per thread), grid sized in multiples of the SM count, fp32 arithmetic.
//
// Reference call sites replaced: nn.BatchNorm3d / nn.ReLU (models/unet3d.py:31-33,37-39), nn.MaxPool3d(2) (:80),
// outc Conv3d k=1 (:222), DiceLoss / BCEDiceLoss (utils/losses.py:44-92,107-152), optim.Adam (utils/trainer.py:113).
#include "bandwidth.cuh"
#include "launch.cuh"

namespace b200 {

#define DEV __device__ __forceinline__

struct alignas(16) Bf8 {
    __nv_bfloat162 v[4];
};

DEV Bf8 ld8(const __nv_bfloat16* p) {
    Bf8 r;
    *reinterpret_cast<uint4*>(&r) = __ldg(reinterpret_cast<const uint4*>(p));
    return r;
}
DEV Bf8 ld8_stream(const __nv_bfloat16* p) {
    Bf8 r;
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                 : "l"(p));
    *reinterpret_cast<uint4*>(&r) = u;
    return r;
}
DEV void st8(__nv_bfloat16* p, const Bf8& v) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v); }
DEV void unpack8(const Bf8& b, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(b.v[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
DEV Bf8 pack8(const float (&f)[8]) {
    Bf8 b;
#pragma unroll
    for (int i = 0; i < 4; ++i) b.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return b;
}
DEV void ldf8(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

DEV uint32_t pack_bf16x2_(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

static int grid_for(long long work_items, int threads, int sms, int waves) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = (long long)sms * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ------------------------------------------------------------------------------------------------ pack input
__global__ void pack_input_kernel(const float* __restrict__ x, long long nvox_per_n, int c_in, View out) {
    pdl_wait();
    const long long total = out.n * nvox_per_n;
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        const long long nb = v / nvox_per_n, sp = v - nb * nvox_per_n;
        const float* src = x + nb * c_in * nvox_per_n + sp;
        __nv_bfloat16* dst = out.p + v * out.ld;
        for (int c0 = 0; c0 < (int)out.c; c0 += 8) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (c0 + j < c_in) ? __ldg(src + (long long)(c0 + j) * nvox_per_n) : 0.f;
            st8(dst + c0, pack8(f));
        }
    }
}
cudaError_t launch_pack_input(const float* x, long long n, long long c, long long d, long long h, long long w,
                              View out, cudaStream_t s) {
    const long long total = n * d * h * w;
    launch_k(pack_input_kernel, grid_for(total, 256, 148, 16), 256, 0, s, x, d * h * w, (int)c, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ im2col of the input
// First conv (Cin = n_modalities = 5): K = 27*5 = 135 is too thin for 27 separate 64-channel TMA boxes, so the input
// is expanded once to rows of kpad = pad16(27*C) bf16, column k = c*27 + tap (tap = kd*9+kh*3+kw): exactly the
// flattened (C,3,3,3) weight layout, so the layer becomes a plain GEMM (1 tap) for fprop and wgrad.
constexpr int kI2cVox = 128;
// 256 threads = 128 voxels x 2 channel groups; each thread gathers whole channels (27 taps unrolled, coalesced along w
// across the warp, L1 reuse 27x) into a bf16 tile [128][kpad + 2] in shared memory; the block then streams the tile
// out as contiguous 4-byte words (rows are contiguous in global memory when ld == kpad).
__global__ void __launch_bounds__(256) im2col_input_kernel(const float* __restrict__ x, int c_in, int D, int H, int W,
                                                           long long nvox, View out) {
    pdl_wait();
    extern __shared__ uint32_t i2c_smem[];  // [128][kpad/2 + 1] words
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(i2c_smem);
    const int kpad = (int)out.c, pitch = kpad / 2 + 1, pitch16 = 2 * pitch;
    const long long v0 = (long long)blockIdx.x * kI2cVox;
    const int vl = threadIdx.x & (kI2cVox - 1), half = threadIdx.x >> 7;
    const long long v = v0 + vl;
    const long long plane = (long long)D * H * W;
    if (v < nvox) {
        const long long nb = v / plane;
        long long r = v - nb * plane;
        const int d = (int)(r / ((long long)H * W));
        r -= (long long)d * H * W;
        const int h = (int)(r / W), w = (int)(r - (long long)h * W);
        const float* xb = x + nb * c_in * plane + ((long long)d * H + h) * W + w;
        bool okd[3], okh[3], okw[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            okd[i] = (unsigned)(d + i - 1) < (unsigned)D;
            okh[i] = (unsigned)(h + i - 1) < (unsigned)H;
            okw[i] = (unsigned)(w + i - 1) < (unsigned)W;
        }
        __nv_bfloat16* row = tile + vl * pitch16;
        for (int c = half; c < c_in; c += 2) {
            const float* xc = xb + c * plane;
#pragma unroll
            for (int t = 0; t < 27; ++t) {
                const int kd = t / 9, kh = (t / 3) % 3, kw = t % 3;
                float val = 0.f;
                if (okd[kd] && okh[kh] && okw[kw])
                    val = __ldg(xc + ((long long)(kd - 1) * H + (kh - 1)) * W + (kw - 1));
                row[c * 27 + t] = __float2bfloat16_rn(val);
            }
        }
        for (int k = 27 * c_in + half; k < kpad; k += 2) row[k] = __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    const int words_per_row = kpad / 2;
    const long long rows = (nvox - v0) < kI2cVox ? (nvox - v0) : kI2cVox;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out.p + v0 * out.ld);
    const int ld_words = (int)(out.ld / 2);
    for (int i = threadIdx.x; i < rows * words_per_row; i += blockDim.x) {
        const int row = i / words_per_row, col = i - row * words_per_row;
        dst[(long long)row * ld_words + col] = i2c_smem[row * pitch + col];
    }
}
// Specialisation for a compile-time channel count (the 5-modality input): the kernel above is issue-bound (two
// instructions per 2-byte element plus a divide per copied word).  Here thread (voxel, half) builds words
// [36*half, 36*half + 36) of its row: every (channel, tap) of a word is a compile-time constant, two taps are converted
// and stored as one 32-bit word, and the copy-out walks whole rows per warp without divisions.
template <int CIN>
__global__ void __launch_bounds__(256) im2col_input_fixed_kernel(const float* __restrict__ x, int D, int H, int W,
                                                                 long long nvox, View out) {
    pdl_wait();
    constexpr int K = 27 * CIN, KPAD = (K + 15) / 16 * 16, WORDS = KPAD / 2, PITCH = WORDS | 1, HW = WORDS / 2;
    extern __shared__ uint32_t i2c_smem[];  // [128][PITCH] words
    const long long v0 = (long long)blockIdx.x * kI2cVox;
    const int vl = threadIdx.x & (kI2cVox - 1), half = threadIdx.x >> 7;   // half is warp-uniform
    const long long v = v0 + vl;
    const int plane = D * H * W, hw = H * W;
    if (v < nvox) {
        const int nb = (int)(v / plane);
        int r = (int)(v - (long long)nb * plane);
        const int d = r / hw;
        r -= d * hw;
        const int h = r / W, w = r - h * W;
        const float* xb = x + (long long)nb * CIN * plane + (d * hw + h * W + w);
        bool okdh[9], okw[3];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            okdh[i] = (unsigned)(d + i / 3 - 1) < (unsigned)D && (unsigned)(h + i % 3 - 1) < (unsigned)H;
#pragma unroll
        for (int i = 0; i < 3; ++i) okw[i] = (unsigned)(w + i - 1) < (unsigned)W;
        uint32_t* row = i2c_smem + vl * PITCH;
        auto tap = [&](int k) -> float {
            if (k >= K) return 0.f;
            const int c = k / 27, t = k % 27, kd = t / 9, kh = (t / 3) % 3, kw = t % 3;
            if (!(okdh[kd * 3 + kh] && okw[kw])) return 0.f;
            return __ldg(xb + ((long long)c * plane + (kd - 1) * hw + (kh - 1) * W + (kw - 1)));
        };
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < HW; ++j) row[j] = pack_bf16x2_(tap(2 * j), tap(2 * j + 1));
        } else {
#pragma unroll
            for (int j = HW; j < WORDS; ++j) row[j] = pack_bf16x2_(tap(2 * j), tap(2 * j + 1));
        }
    }
    __syncthreads();
    const int rows = (int)((nvox - v0) < kI2cVox ? (nvox - v0) : kI2cVox);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out.p + v0 * out.ld);
    const int ld_words = (int)(out.ld / 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp; rr < rows; rr += 8) {
        const uint32_t* src = i2c_smem + rr * PITCH;
        uint32_t* o = dst + (long long)rr * ld_words;
#pragma unroll
        for (int c = 0; c < (WORDS + 31) / 32; ++c)
            if (c * 32 + lane < WORDS) o[c * 32 + lane] = src[c * 32 + lane];
    }
}
cudaError_t launch_im2col_input(const float* x, long long n, long long c, long long d, long long h, long long w,
                                View out, cudaStream_t s) {
    const long long nvox = n * d * h * w;
    const long long blocks = (nvox + kI2cVox - 1) / kI2cVox;
    if (c == 5 && out.c == 144 && d * h * w < (1LL << 31)) {
        const int smem = kI2cVox * 73 * 4;
        launch_k(im2col_input_fixed_kernel<5>, (unsigned)blocks, 256, smem, s, x, (int)d, (int)h, (int)w, nvox, out);
        return cudaGetLastError();
    }
    const int smem = kI2cVox * ((int)out.c / 2 + 1) * 4;
    launch_k(im2col_input_kernel, (unsigned)blocks, 256, smem, s, x, (int)c, (int)d, (int)h, (int)w, nvox, out);
    return cudaGetLastError();
}

// fp32 [rows][k] -> bf16 [rows][kpad] (zero padded): GEMM-form weights of the im2col'd first conv
__global__ void pack_rows_kernel(const float* __restrict__ w, int rows, int k, int kpad, __nv_bfloat16* out) {
    pdl_wait();
    const long long total = (long long)rows * kpad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / kpad), col = (int)(i - (long long)r * kpad);
        out[i] = __float2bfloat16_rn(col < k ? __ldg(w + (long long)r * k + col) : 0.f);
    }
}
cudaError_t launch_pack_rows(const float* w, int rows, int k, int kpad, __nv_bfloat16* out, cudaStream_t s) {
    launch_k(pack_rows_kernel, grid_for((long long)rows * kpad, 256, 148, 8), 256, 0, s, w, rows, k, kpad, out);
    return cudaGetLastError();
}

// first-layer weight (Cout, Cin, 3, 3, 3) fp32 -> bf16 [3 kd][Cout][64]: row k = c*9 + kh*3 + kw of depth tap kd
// (Cin*9 <= 64 real columns, zero padded) — the K-major B operand of conv1_march.cu
__global__ void pack_conv1_slices_kernel(const float* __restrict__ w, int cout, int cin, __nv_bfloat16* out) {
    pdl_wait();
    const int total = 3 * cout * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i & 63, co = (i >> 6) % cout, kd = (i >> 6) / cout;
        float v = 0.f;
        if (k < cin * 9) {
            const int c = k / 9, r = k - 9 * c;
            v = __ldg(w + ((long long)(co * cin + c) * 3 + kd) * 9 + r);
        }
        out[i] = __float2bfloat16_rn(v);
    }
}
cudaError_t launch_pack_conv1_slices(const float* w, int cout, int cin, __nv_bfloat16* out, cudaStream_t s) {
    launch_k(pack_conv1_slices_kernel, grid_for(3LL * cout * 64, 256, 148, 8), 256, 0, s, w, cout, cin, out);
    return cudaGetLastError();
}

// bf16 [taps][rows][cols] -> [taps][cols][rows] (32 x 32 tiles through shared memory): the K-major form of a packed
// conv weight for the other GEMM direction (dgrad on the CTA-pair depth-marching kernel)
__global__ void __launch_bounds__(256) transpose_taps_kernel(const __nv_bfloat16* __restrict__ src, int rows, int cols,
                                                             __nv_bfloat16* __restrict__ dst) {
    pdl_wait();
    __shared__ __nv_bfloat16 tile[32][33];
    const long long base = (long long)blockIdx.z * rows * cols;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = src[base + (long long)(r0 + i) * cols + c0 + tx];
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) dst[base + (long long)(c0 + i) * rows + r0 + tx] = tile[tx][i];
}
cudaError_t launch_transpose_taps(const __nv_bfloat16* src, int taps, int rows, int cols, __nv_bfloat16* dst,
                                  cudaStream_t s) {
    launch_k(transpose_taps_kernel, dim3((cols + 31) / 32, (rows + 31) / 32, taps), 256, 0, s, src, rows, cols, dst);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ pack weights
// one block = 32 output channels x 32 input channels x 27 taps; the tile is converted to bf16 on the way into
// shared memory ([27][32 co][32 ci], row pitch 34 to spread banks), then written out as [27][Cout][Cin] with 64-byte rows
// (the one packed layout serves fprop as a K-major and dgrad as an MN-major B operand).  55 KB of shared memory -> 4 blocks per SM.
constexpr int kPackPitch = 34;
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin,
                                                               int cin_pad, __nv_bfloat16* __restrict__ wf) {
    pdl_wait();
    extern __shared__ __nv_bfloat16 ptile[];  // [27][32][kPackPitch]
    const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
    const int tid = threadIdx.x;
    const int nci = min(32, cin - ci0);        // valid input channels of this tile (may be <= 0 in the padded tail)
    const int nco = min(32, cout - co0);
    const int row_len = max(nci, 0) * 27;      // contiguous floats per output channel
    const bool vec = (nci == 32) && ((((long long)cin * 27) & 3) == 0);
    if (vec) {
        for (int i = tid; i < 32 * 216; i += 256) {
            const int r = i / 216, c4 = i - r * 216;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nco) v = __ldg(reinterpret_cast<const float4*>(w + ((long long)(co0 + r) * cin + ci0) * 27) + c4);
            const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = c4 * 4 + e, ci = k / 27, t = k - ci * 27;
                ptile[(t * 32 + r) * kPackPitch + ci] = __float2bfloat16_rn(f[e]);
            }
        }
    } else {
        for (int i = tid; i < 32 * 864; i += 256) {
            const int r = i / 864, k = i - r * 864;
            const int ci = k / 27, t = k - ci * 27;
            float v = 0.f;
            if (r < nco && k < row_len) v = __ldg(w + ((long long)(co0 + r) * cin + ci0) * 27 + k);
            ptile[(t * 32 + r) * kPackPitch + ci] = __float2bfloat16_rn(v);
        }
    }
    __syncthreads();
    const int lane = tid & 31, wrp = tid >> 5;
    // packed tap order t = kd*9 + kw*3 + kh (kh fastest: the three kh taps of a (kd,kw) group are one TMA box);
    // native (torch) tap index tn = kd*9 + kh*3 + kw
    for (int job = wrp; job < 27 * 32; job += 8) {
        const int t = job >> 5, r = job & 31;
        const int tn = (t / 9) * 9 + (t % 3) * 3 + (t / 3) % 3;
        {  // [27][cout][cin_pad], ci contiguous: row r = output channel, lane = input channel
            const int co = co0 + r, ci = ci0 + lane;
            if (co < cout && ci < cin_pad) wf[((long long)t * cout + co) * cin_pad + ci] = ptile[(tn * 32 + r) * kPackPitch + lane];
        }
    }
}
cudaError_t launch_pack_conv_weight(const float* w, int cout, int cin, int cin_pad, __nv_bfloat16* wf,
                                    cudaStream_t s) {
    const int smem = 27 * 32 * kPackPitch * 2;
    {   // the > 48 KB opt-in is per device; setting it again is cheap and this launch is off the hot path
        cudaError_t e =
            cudaFuncSetAttribute(pack_conv_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((cin_pad + 31) / 32, (cout + 31) / 32);
    launch_k(pack_conv_weight_kernel, grid, 256, smem, s, w, cout, cin, cin_pad, wf);
    return cudaGetLastError();
}

// w (cin, cout, 8) -> wf [8*cout][cin] (ci contiguous), wd [8][cin][cout] (co contiguous)
__global__ void pack_convt_weight_kernel(const float* __restrict__ w, const float* __restrict__ bias, int cin,
                                         int cout, __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd,
                                         float* __restrict__ bias8) {
    pdl_wait();
    const long long total = (long long)cin * cout * 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        {  // wf index: (t, co, ci)
            const int ci = (int)(i % cin);
            const long long r = i / cin;
            const int co = (int)(r % cout), t = (int)(r / cout);
            wf[i] = __float2bfloat16_rn(__ldg(w + ((long long)ci * cout + co) * 8 + t));
        }
        {  // wd index: (t, ci, co)
            const int co = (int)(i % cout);
            const long long r = i / cout;
            const int ci = (int)(r % cin), t = (int)(r / cin);
            wd[i] = __float2bfloat16_rn(__ldg(w + ((long long)ci * cout + co) * 8 + t));
        }
        if (i < 8LL * cout) bias8[i] = __ldg(bias + (i % cout));
    }
}
cudaError_t launch_pack_convt_weight(const float* w, const float* bias, int cin, int cout, __nv_bfloat16* wf,
                                     __nv_bfloat16* wd, float* bias8, cudaStream_t s) {
    launch_k(pack_convt_weight_kernel, grid_for((long long)cin * cout * 8, 256, 148, 8), 256, 0, s, w, bias, cin, cout, wf,
                                                                                             wd, bias8);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ BN finalize
// partial[rows][c][2] -> per-channel (sum0, sum1) in double; block = 32 channels x kRedSlices row slices (1024
// threads: the kernel is a chain of dependent loads, so the rows are spread wide and loaded four at a time)
constexpr int kRedSlices = 32;
DEV void reduce_partials(const float* __restrict__ partial, int rows, int c, int ch, int slice,
                         double (&red)[kRedSlices][32][2], double& s0, double& s1) {
    double a = 0.0, b = 0.0;
    if (ch < c) {
        const float2* src = reinterpret_cast<const float2*>(partial) + ch;
        int r = slice;
        for (; r + 3 * kRedSlices < rows; r += 4 * kRedSlices) {
            const float2 v0 = __ldg(src + (long long)r * c);
            const float2 v1 = __ldg(src + (long long)(r + kRedSlices) * c);
            const float2 v2 = __ldg(src + (long long)(r + 2 * kRedSlices) * c);
            const float2 v3 = __ldg(src + (long long)(r + 3 * kRedSlices) * c);
            a += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
            b += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
        }
        for (; r < rows; r += kRedSlices) {
            const float2 v = __ldg(src + (long long)r * c);
            a += v.x;
            b += v.y;
        }
    }
    red[slice][threadIdx.x & 31][0] = a;
    red[slice][threadIdx.x & 31][1] = b;
    __syncthreads();
    s0 = s1 = 0.0;
    if (slice == 0) {
#pragma unroll
        for (int k = 0; k < kRedSlices; ++k) {
            s0 += red[k][threadIdx.x & 31][0];
            s1 += red[k][threadIdx.x & 31][1];
        }
    }
}

__global__ void __launch_bounds__(32 * kRedSlices) bn_finalize_kernel(const float* __restrict__ partial, int rows,
                                                          double inv_count, double unbias, int c,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, float momentum,
                                                          float* rm, float* rv, long long* nbt, float* mean,
                                                          float* rstd, float* scale, float* shift) {
    pdl_wait();
    __shared__ double red[kRedSlices][32][2];
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) nbt[0] += 1;   // num_batches_tracked
    const int ch = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
    double s, q;
    reduce_partials(partial, rows, c, ch, slice, red, s, q);
    if (slice != 0 || ch >= c) return;
    const double mu = s * inv_count;
    double var = q * inv_count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float muf = (float)mu;
    mean[ch] = muf;
    rstd[ch] = rs;
    const float sc = gamma[ch] * rs;
    scale[ch] = sc;
    shift[ch] = beta[ch] - muf * sc;
    if (rm) rm[ch] = (1.f - momentum) * rm[ch] + momentum * muf;
    if (rv) rv[ch] = (1.f - momentum) * rv[ch] + momentum * (float)(var * unbias);
}
cudaError_t launch_bn_finalize(const float* partial, long long rows, long long count, int c, const float* gamma,
                               const float* beta, float eps, float momentum, float* rm, float* rv, long long* nbt,
                               float* mean, float* rstd, float* scale, float* shift, cudaStream_t s) {
    const double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
    launch_k(bn_finalize_kernel, (c + 31) / 32, 32 * kRedSlices, 0, s, partial, (int)rows, 1.0 / (double)count, unbias, c, gamma, beta,
                                                    eps, momentum, rm, rv, nbt, mean, rstd, scale, shift);
    return cudaGetLastError();
}

__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                    const float* cbias, float eps, int c, float* scale, float* shift) {
    pdl_wait();
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const float sc = gamma[ch] / sqrtf(rv[ch] + eps);
    scale[ch] = sc;
    shift[ch] = beta[ch] + ((cbias ? cbias[ch] : 0.f) - rm[ch]) * sc;
}
cudaError_t launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                                const float* cbias, float eps, int c, float* scale, float* shift, cudaStream_t s) {
    launch_k(bn_fold_eval_kernel, (c + 127) / 128, 128, 0, s, gamma, beta, rm, rv, cbias, eps, c, scale, shift);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ BN apply + ReLU
// Four 16-byte vectors per thread are in flight before the first is used: with one load per thread the kernel sat at
// 5.5 TB/s (ncu: 91 % of the warps resident, 31 % issue, 68 % DRAM — bound by the bytes in flight, not by either pipe).
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(View y, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, View out, FastDiv c8d,
                                                            uint32_t total) {
    pdl_wait();
    constexpr int U = 4;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
        Bf8 in[U];
        uint32_t vox[U], c0[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i < total) {
                vox[u] = c8d.quot(i);
                c0[u] = (i - vox[u] * c8d.div) * 8;
                in[u] = ld8_stream(y.p + (long long)vox[u] * y.ld + c0[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * stride < total) {
                float f[8], sc[8], sh[8];
                unpack8(in[u], f);
                ldf8(scale + c0[u], sc);
                ldf8(shift + c0[u], sh);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
                st8(out.p + (long long)vox[u] * out.ld + c0[u], pack8(f));
            }
        }
    }
}
cudaError_t launch_bn_apply_relu(View y, const float* scale, const float* shift, View out, int sms, cudaStream_t s) {
    const long long total = y.voxels() * (y.c / 8);
    if (total >= (1LL << 31)) return cudaErrorInvalidValue;
    launch_k(bn_apply_relu_kernel, grid_for(total, 256, sms, 16), 256, 0, s, y, scale, shift, out,
                                                                      FastDiv((uint32_t)(y.c / 8)), (uint32_t)total);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ channel-lane skeleton
// 256 threads = rows x c8 lanes; a thread owns 8 fixed channels and walks voxels, so per-channel sums stay in
// registers; one shared-memory reduction per block at the end.
struct LaneMap {
    int c8, rows;
};
static LaneMap lane_map(long long c) {
    LaneMap m;
    m.c8 = (int)(c / 8);
    m.rows = 256 / m.c8;
    return m;
}

// dynamic shared memory of the final per-block reduction (floats): small on purpose — these kernels are meant to
// co-reside with a persistent GEMM CTA that leaves < 30 KB of the SM's shared memory
static size_t reduce_smem_bytes(const LaneMap& m, int nacc) {
    const bool warp_pre = m.c8 <= 32 && (m.c8 & (m.c8 - 1)) == 0;
    const int slabs = warp_pre ? 8 : m.rows;
    return (size_t)slabs * m.c8 * nacc * 8 * sizeof(float);
}

template <int NACC>
DEV void block_reduce_store(float (&acc)[NACC][8], int c8, int rows, int row, int cv, bool active, float* smem,
                            float* dst /* [c][NACC] or atomics */, int c, bool atomic) {
    const bool warp_pre = c8 <= 32 && (c8 & (c8 - 1)) == 0;
    int slabs = rows, slab = row;
    if (warp_pre) {
        // rows that share a warp are folded with shuffles first (lane = row_in_warp * c8 + cv)
        for (int off = c8; off < 32; off <<= 1) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[a][j] += __shfl_xor_sync(0xffffffffu, acc[a][j], off);
        }
        slabs = 8;
        slab = threadIdx.x >> 5;
        active = (threadIdx.x & 31) < c8;  // one lane per channel group holds the warp's total
    }
    // smem [slabs][c8][NACC*8]
    if (active) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) smem[((slab * c8 + cv) * NACC + a) * 8 + j] = acc[a][j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < c * NACC; t += blockDim.x) {
        const int ch = t / NACC, a = t - ch * NACC;
        const int cvv = ch >> 3, j = ch & 7;
        float s = 0.f;
        for (int r = 0; r < slabs; ++r) s += smem[((r * c8 + cvv) * NACC + a) * 8 + j];
        if (atomic)
            atomicAdd(dst + ch * NACC + a, s);
        else
            dst[ch * NACC + a] = s;
    }
}

// ------------------------------------------------------------------------------------------------ BN backward
// Both backward passes are software-pipelined: a thread issues the loads of its next group of kBwdU voxels before it
// computes on the current one, so every resident warp has loads in flight all the time.  (The straight loops ran at
// 5.1 TB/s with 24-28 % of the warps resident — registers — and 45-49 % issue: bound by the bytes in flight.)
constexpr int kBwdU = 4;
struct BwdGroup {
    Bf8 g[kBwdU], x[kBwdU];
};
DEV void bwd_load(BwdGroup& t, const View& dout, const View& y, long long v0, long long step, long long nvox, int c0) {
#pragma unroll
    for (int u = 0; u < kBwdU; ++u) {
        const long long v = v0 + u * step;
        if (v < nvox) {
            t.g[u] = ld8_stream(dout.p + v * dout.ld + c0);
            t.x[u] = ld8_stream(y.p + v * y.ld + c0);
        }
    }
}
// walks groups v0, v0 + kBwdU * step, ... with two register sets (ping-pong: no copies)
template <class Compute>
DEV void bwd_pipeline(const View& dout, const View& y, long long v0, long long step, long long nvox, int c0,
                      Compute&& compute) {
    BwdGroup a, b;
    const long long big = kBwdU * step;
    if (v0 < nvox) bwd_load(a, dout, y, v0, step, nvox, c0);
    while (v0 < nvox) {
        const long long v1 = v0 + big;
        if (v1 < nvox) bwd_load(b, dout, y, v1, step, nvox, c0);
        compute(a, v0);
        if (v1 >= nvox) break;
        const long long v2 = v1 + big;
        if (v2 < nvox) bwd_load(a, dout, y, v2, step, nvox, c0);
        compute(b, v1);
        v0 = v2;
    }
}

__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(View dout, View y, const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, float* partial, int c8,
                                                               int rows, long long nvox) {
    pdl_wait();
    extern __shared__ float smem[];
    const int row = threadIdx.x / c8, cv = threadIdx.x - row * c8;
    const bool active = row < rows;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
    if (active) {
        float sc[8], sh[8], rs[8], nmr[8];   // x-hat = x * rstd - mean * rstd
        ldf8(scale + cv * 8, sc);
        ldf8(shift + cv * 8, sh);
        ldf8(rstd + cv * 8, rs);
        ldf8(mean + cv * 8, nmr);
#pragma unroll
        for (int j = 0; j < 8; ++j) nmr[j] = -nmr[j] * rs[j];
        const long long step = (long long)gridDim.x * rows;
        bwd_pipeline(dout, y, (long long)blockIdx.x * rows + row, step, nvox, cv * 8,
                     [&](const BwdGroup& t, long long v0) {
#pragma unroll
                         for (int u = 0; u < kBwdU; ++u) {
                             if (v0 + u * step < nvox) {
                                 float g[8], x[8];
                                 unpack8(t.g[u], g);
                                 unpack8(t.x[u], x);
#pragma unroll
                                 for (int j = 0; j < 8; ++j) {
                                     const float gm = fmaf(x[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
                                     acc[0][j] += gm;
                                     acc[1][j] = fmaf(gm, fmaf(x[j], rs[j], nmr[j]), acc[1][j]);
                                 }
                             }
                         }
                     });
    }
    block_reduce_store<2>(acc, c8, rows, row, cv, active, smem, partial + (long long)blockIdx.x * y.c * 2, (int)y.c,
                          false);
}
cudaError_t launch_bn_bwd_reduce(View dout, View y, const float* scale, const float* shift, const float* mean,
                                 const float* rstd, float* partial, int* nblk, cudaStream_t s) {
    const LaneMap m = lane_map(y.c);
    const long long nvox = y.voxels();
    long long blocks = (nvox + m.rows * 8LL - 1) / (m.rows * 8LL);
    if (blocks > kBwdMaxBlocks) blocks = kBwdMaxBlocks;
    if (blocks < 1) blocks = 1;
    *nblk = (int)blocks;
    const size_t smem = reduce_smem_bytes(m, 2);
    launch_k(bn_bwd_reduce_kernel, (int)blocks, 256, smem, s, dout, y, scale, shift, mean, rstd, partial, m.c8, m.rows,
                                                       nvox);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(32 * kRedSlices) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int c,
                                                              double inv_count, float* dgamma, float* dbeta,
                                                              float* coef) {
    pdl_wait();
    __shared__ double red[kRedSlices][32][2];
    const int ch = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
    double s1, s2;
    reduce_partials(partial, nblk, c, ch, slice, red, s1, s2);
    if (slice != 0 || ch >= c) return;
    if (dbeta) dbeta[ch] += (float)s1;
    if (dgamma) dgamma[ch] += (float)s2;
    coef[2 * ch] = (float)(s1 * inv_count);
    coef[2 * ch + 1] = (float)(s2 * inv_count);
}
cudaError_t launch_bn_bwd_finalize(const float* partial, int nblk, int c, long long count, float* dgamma,
                                   float* dbeta, float* coef, cudaStream_t s) {
    launch_k(bn_bwd_finalize_kernel, (c + 31) / 32, 32 * kRedSlices, 0, s, partial, nblk, c, 1.0 / (double)count, dgamma, dbeta, coef);
    return cudaGetLastError();
}

// dy = gamma * rstd * (dy_m - coef0 - x-hat * coef1) = scale * dy_m + (P * x + Q) with per-channel
// P = -scale * coef1 * rstd, Q = scale * (coef1 * mean * rstd - coef0): two FMAs per element
DEV void bwd_apply_consts(const float* scale, const float* mean, const float* rstd, const float* coef, int cv,
                          float (&sc)[8], float (&P)[8], float (&Q)[8]) {
    float mu[8], rs[8], t[8], u[8];
    ldf8(scale + cv * 8, sc);
    ldf8(mean + cv * 8, mu);
    ldf8(rstd + cv * 8, rs);
    ldf8(coef + cv * 16, t);       // coef[c][2] = (sum dy_m, sum dy_m * x-hat) / count
    ldf8(coef + cv * 16 + 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float k0 = j < 4 ? t[2 * j] : u[2 * j - 8], k1 = j < 4 ? t[2 * j + 1] : u[2 * j - 7];
        P[j] = -sc[j] * k1 * rs[j];
        Q[j] = sc[j] * (k1 * mu[j] * rs[j] - k0);
    }
}

__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(View dout, View y, const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ rstd,
                                                              const float* __restrict__ coef, View dy, float* dbias,
                                                              int c8, int rows, long long nvox) {
    pdl_wait();
    extern __shared__ float smem[];
    const int row = threadIdx.x / c8, cv = threadIdx.x - row * c8;
    const bool active = row < rows;
    float acc[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
    if (active) {
        float sc[8], sh[8], P[8], Q[8];
        ldf8(shift + cv * 8, sh);
        bwd_apply_consts(scale, mean, rstd, coef, cv, sc, P, Q);
        const long long step = (long long)gridDim.x * rows;
        bwd_pipeline(dout, y, (long long)blockIdx.x * rows + row, step, nvox, cv * 8,
                     [&](const BwdGroup& t, long long v0) {
#pragma unroll
                         for (int u = 0; u < kBwdU; ++u) {
                             const long long v = v0 + u * step;
                             if (v < nvox) {
                                 float g[8], x[8], o[8];
                                 unpack8(t.g[u], g);
                                 unpack8(t.x[u], x);
#pragma unroll
                                 for (int j = 0; j < 8; ++j) {
                                     const float gm = fmaf(x[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;
                                     o[j] = fmaf(sc[j], gm, fmaf(x[j], P[j], Q[j]));
                                 }
                                 const Bf8 ob = pack8(o);
                                 st8(dy.p + v * dy.ld + cv * 8, ob);
                                 float r[8];
                                 unpack8(ob, r);
#pragma unroll
                                 for (int j = 0; j < 8; ++j) acc[0][j] += r[j];
                             }
                         }
                     });
    }
    if (dbias) block_reduce_store<1>(acc, c8, rows, row, cv, active, smem, dbias, (int)y.c, true);
}
cudaError_t launch_bn_bwd_apply(View dout, View y, const float* scale, const float* shift, const float* mean,
                                const float* rstd, const float* coef, View dy, float* dbias, cudaStream_t s) {
    const LaneMap m = lane_map(y.c);
    const long long nvox = y.voxels();
    long long blocks = (nvox + m.rows * 8LL - 1) / (m.rows * 8LL);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    const size_t smem = reduce_smem_bytes(m, 1);
    launch_k(bn_bwd_apply_kernel, (int)blocks, 256, smem, s, dout, y, scale, shift, mean, rstd, coef, dy, dbias, m.c8, m.rows,
                                                      nvox);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ channel sum
// per-channel sum over a box [d0, d0+bd) x [h0, h0+bh) x [w0, w0+bw) of every sample (the whole volume by default):
// the bias gradient of a transposed convolution is the sum of the output gradient over the un-padded core of the
// concat half (the F.pad border, models/unet3d.py:149-151, carries no bias)
__global__ void __launch_bounds__(256) channel_sum_kernel(View v, float* out, int c8, int rows, long long nvox,
                                                          int d0, int h0, int w0, FastDiv bwd, FastDiv bhd,
                                                          FastDiv bdd, int boxed) {
    pdl_wait();
    extern __shared__ float smem[];
    const int row = threadIdx.x / c8, cv = threadIdx.x - row * c8;
    const bool active = row < rows;
    float acc[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
    if (active) {
        constexpr int U = 4;   // loads in flight per thread (one per iteration ran at half the copy bandwidth)
        const long long step = (long long)gridDim.x * rows;
        for (long long i0 = (long long)blockIdx.x * rows + row; i0 < nvox; i0 += U * step) {
            Bf8 t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * step;
                if (i < nvox) {
                    long long vox = i;
                    if (boxed) {
                        uint32_t r = (uint32_t)i, q = bwd.quot(r);
                        const uint32_t w = r - q * bwd.div;
                        r = q; q = bhd.quot(r);
                        const uint32_t h = r - q * bhd.div;
                        r = q; q = bdd.quot(r);
                        const uint32_t d = r - q * bdd.div;
                        vox = (((long long)q * v.d + d0 + d) * v.h + h0 + h) * v.w + w0 + w;
                    }
                    t[u] = ld8_stream(v.p + vox * v.ld + cv * 8);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * step < nvox) {
                    float x[8];
                    unpack8(t[u], x);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[0][j] += x[j];
                }
            }
        }
    }
    block_reduce_store<1>(acc, c8, rows, row, cv, active, smem, out, (int)v.c, true);
}
cudaError_t launch_channel_sum(View v, float* out, const int* box, cudaStream_t s) {
    const LaneMap m = lane_map(v.c);
    long long nvox = v.voxels();
    int boxed = 0, d0 = 0, h0 = 0, w0 = 0;
    FastDiv bw, bh, bd;
    if (box) {
        boxed = 1;
        d0 = box[0]; h0 = box[1]; w0 = box[2];
        nvox = v.n * (long long)box[3] * box[4] * box[5];
        if (nvox >= (1LL << 31)) return cudaErrorInvalidValue;
        bd = FastDiv((uint32_t)box[3]); bh = FastDiv((uint32_t)box[4]); bw = FastDiv((uint32_t)box[5]);
    }
    long long blocks = (nvox + m.rows * 8LL - 1) / (m.rows * 8LL);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    const size_t smem = reduce_smem_bytes(m, 1);
    launch_k(channel_sum_kernel, (int)blocks, 256, smem, s, v, out, m.c8, m.rows, nvox, d0, h0, w0, bw, bh, bd, boxed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ MaxPool3d(2)
// torch semantics: window max with NaN propagation; backward routes the gradient to the first maximum in
// d, h, w scan order (ATen max_pool3d_with_indices: `val > max || isnan(val)` keeps the earliest equal value).
DEV __nv_bfloat162 max2_nan(__nv_bfloat162 a, __nv_bfloat162 b) { return __hmax2_nan(a, b); }

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(View x, View y, FastDiv c8d, FastDiv owd, FastDiv ohd,
                                                          FastDiv odd, uint32_t total) {
    pdl_wait();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t r = c8d.quot(i);
        const uint32_t c0 = (i - r * c8d.div) * 8;
        const uint32_t ovox = r;
        uint32_t q = owd.quot(r);
        const uint32_t ow = r - q * owd.div;
        r = q; q = ohd.quot(r);
        const uint32_t oh = r - q * ohd.div;
        r = q; q = odd.quot(r);
        const uint32_t od = r - q * odd.div;
        const uint32_t nb = q;
        const long long base = (((long long)nb * x.d + 2 * od) * x.h + 2 * oh) * x.w + 2 * ow;
        Bf8 m = ld8_stream(x.p + base * x.ld + c0);
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            const long long off = base + ((k >> 2) & 1) * x.h * x.w + ((k >> 1) & 1) * x.w + (k & 1);
            const Bf8 t = ld8_stream(x.p + off * x.ld + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) m.v[j] = max2_nan(m.v[j], t.v[j]);
        }
        st8(y.p + (long long)ovox * y.ld + c0, m);
    }
}
cudaError_t launch_maxpool_fwd(View x, View y, int sms, cudaStream_t s) {
    const long long total = y.voxels() * (y.c / 8);
    if (total >= (1LL << 31)) return cudaErrorInvalidValue;
    if (total == 0) return cudaSuccess;
    launch_k(maxpool_fwd_kernel, grid_for(total, 256, sms, 16), 256, 0, s, x, y, FastDiv((uint32_t)(y.c / 8)),
                                                                    FastDiv((uint32_t)y.w), FastDiv((uint32_t)y.h),
                                                                    FastDiv((uint32_t)y.d), (uint32_t)total);
    return cudaGetLastError();
}

// one thread = one 2x2x2 cell of the input grid (cells cover ceil(dim/2)) x 8 channels; all selection logic stays in
// packed bf16x2 (compare masks), so the eight input vectors cost 32 registers instead of 64
DEV uint32_t bits(const __nv_bfloat162& v) { return *reinterpret_cast<const uint32_t*>(&v); }
DEV __nv_bfloat162 from_bits(uint32_t u) { return *reinterpret_cast<__nv_bfloat162*>(&u); }
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(View x, View dy, View dskip, int has_skip, View dx,
                                                          FastDiv c8d, FastDiv cwd, FastDiv chd, FastDiv cdd,
                                                          uint32_t total) {
    pdl_wait();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t r = c8d.quot(i);
        const uint32_t c0 = (i - r * c8d.div) * 8;
        uint32_t q = cwd.quot(r);
        const uint32_t cw = r - q * cwd.div;
        r = q; q = chd.quot(r);
        const uint32_t ch = r - q * chd.div;
        r = q; q = cdd.quot(r);
        const uint32_t cd = r - q * cdd.div;
        const uint32_t nb = q;
        const bool full = (2 * cd + 1 < x.d) && (2 * ch + 1 < x.h) && (2 * cw + 1 < x.w);
        const long long v000 = (((long long)nb * x.d + 2 * cd) * x.h + 2 * ch) * x.w + 2 * cw;
        Bf8 xin[8], sk[8], g, mx;
        if (full) {
            const long long ov = (((long long)nb * dy.d + cd) * dy.h + ch) * dy.w + cw;
            g = ld8_stream(dy.p + ov * dy.ld + c0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long v = v000 + ((k >> 2) & 1) * x.h * x.w + ((k >> 1) & 1) * x.w + (k & 1);
                xin[k] = ld8_stream(x.p + v * x.ld + c0);
            }
            if (has_skip) {   // the skip gradients are in flight together with the window's inputs
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const long long v = v000 + ((k >> 2) & 1) * x.h * x.w + ((k >> 1) & 1) * x.w + (k & 1);
                    sk[k] = ld8_stream(dskip.p + v * dskip.ld + c0);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 m = xin[0].v[j];
#pragma unroll
                for (int k = 1; k < 8; ++k) m = __hmax2(m, xin[k].v[j]);
                mx.v[j] = m;
            }
        }
        uint32_t taken[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int kd = (k >> 2) & 1, kh = (k >> 1) & 1, kw = k & 1;
            const uint32_t d = 2 * cd + kd, h = 2 * ch + kh, w = 2 * cw + kw;
            if (d < x.d && h < x.h && w < x.w) {
                const long long v = v000 + (long long)kd * x.h * x.w + kh * x.w + kw;
                Bf8 o;
                if (has_skip) {
                    if (full) o = sk[k];
                    else o = ld8_stream(dskip.p + v * dskip.ld + c0);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) o.v[j] = from_bits(0u);
                }
                if (full) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // first position (d, h, w scan order) equal to the window maximum takes the gradient
                        const uint32_t hit = __heq2_mask(xin[k].v[j], mx.v[j]) & ~taken[j];
                        taken[j] |= hit;
                        o.v[j] = __hadd2(o.v[j], from_bits(bits(g.v[j]) & hit));
                    }
                }
                st8(dx.p + v * dx.ld + c0, o);
            }
        }
    }
}
cudaError_t launch_maxpool_bwd(View x, View dy, const View* dskip, View dx, int sms, cudaStream_t s) {
    const long long cw = (x.w + 1) / 2, ch = (x.h + 1) / 2, cd = (x.d + 1) / 2;
    const long long total = x.n * cd * ch * cw * (x.c / 8);
    if (total >= (1LL << 31)) return cudaErrorInvalidValue;
    View sk = dskip ? *dskip : dx;
    launch_k(maxpool_bwd_kernel, grid_for(total, 256, sms, 16), 256, 0, s, 
        x, dy, sk, dskip ? 1 : 0, dx, FastDiv((uint32_t)(x.c / 8)), FastDiv((uint32_t)cw), FastDiv((uint32_t)ch),
        FastDiv((uint32_t)cd), (uint32_t)total);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ head 1x1x1
constexpr int kMaxCls = 8;
constexpr int kMaxHeadC = 256;
__global__ void __launch_bounds__(256) head_fwd_kernel(View x, const float* __restrict__ w,
                                                       const float* __restrict__ b, int ncls, float* logits,
                                                       float* probs, long long nvox_per_n) {
    pdl_wait();
    __shared__ float ws[kMaxCls * kMaxHeadC];
    __shared__ float bs[kMaxCls];
    const int c = (int)x.c;
    for (int i = threadIdx.x; i < ncls * c; i += blockDim.x) ws[i] = w[i];
    if (threadIdx.x < ncls) bs[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const long long total = x.voxels();
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        float acc[kMaxCls];
#pragma unroll
        for (int k = 0; k < kMaxCls; ++k) acc[k] = 0.f;
        const __nv_bfloat16* src = x.p + v * x.ld;
        for (int c0 = 0; c0 < c; c0 += 8) {
            float f[8];
            unpack8(ld8_stream(src + c0), f);
#pragma unroll
            for (int k = 0; k < kMaxCls; ++k) {
                if (k < ncls) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], ws[k * c + c0 + j], acc[k]);
                }
            }
        }
        const long long nb = v / nvox_per_n, sp = v - nb * nvox_per_n;
#pragma unroll
        for (int k = 0; k < kMaxCls; ++k) {
            if (k < ncls) {
                const float z = acc[k] + bs[k];
                const long long o = (nb * ncls + k) * nvox_per_n + sp;
                logits[o] = z;
                if (probs) probs[o] = 1.f / (1.f + __expf(-z));
            }
        }
    }
}
// Coalesced form for power-of-two channel counts (c/8 lanes per voxel, 16 bytes per lane: a warp reads whole 128-byte
// lines; the one-thread-per-voxel kernel above touches 32 lines per load instruction and ran at ~1.1 TB/s in the step).
// Four voxels per thread are in flight; the class dot products are reduced across the c/8 lanes with shuffles.
template <int NCLS>
__global__ void __launch_bounds__(256) head_fwd_vec_kernel(View x, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* logits, float* probs,
                                                           int c8, long long nvox, long long nvox_per_n) {
    pdl_wait();
    constexpr int U = 4;
    const int rows = 256 / c8, g = threadIdx.x / c8, cv = threadIdx.x - g * c8;
    const int c = (int)x.c;
    float wk[NCLS][8], bk[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) {
        ldf8(w + k * c + cv * 8, wk[k]);
        bk[k] = __ldg(b + k);
    }
    const long long step = (long long)gridDim.x * rows * U;
    for (long long v0 = (long long)blockIdx.x * rows * U; v0 < nvox; v0 += step) {
        Bf8 a[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long v = v0 + u * rows + g;
            ok[u] = v < nvox;
            if (ok[u]) a[u] = ld8_stream(x.p + v * x.ld + cv * 8);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float f[8];
            if (ok[u]) unpack8(a[u], f);
            float acc[NCLS];
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
                acc[k] = 0.f;
                if (ok[u]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], wk[k][j], acc[k]);
                }
                for (int o = c8 >> 1; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
            if (ok[u] && cv == 0) {
                const long long v = v0 + u * rows + g;
                const long long nb = v / nvox_per_n, sp = v - nb * nvox_per_n;
#pragma unroll
                for (int k = 0; k < NCLS; ++k) {
                    const float z = acc[k] + bk[k];
                    const long long o = (nb * NCLS + k) * nvox_per_n + sp;
                    logits[o] = z;
                    if (probs) probs[o] = 1.f / (1.f + __expf(-z));
                }
            }
        }
    }
}
template <int NCLS>
static cudaError_t head_fwd_vec_launch(View x, const float* w, const float* b, float* logits, float* probs,
                                       cudaStream_t s) {
    const int c8 = (int)(x.c / 8);
    const long long nvox = x.voxels();
    long long blocks = (nvox + (256 / c8) * 4 - 1) / ((256 / c8) * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    launch_k(head_fwd_vec_kernel<NCLS>, (int)blocks, 256, 0, s, x, w, b, logits, probs, c8, nvox, x.d * x.h * x.w);
    return cudaGetLastError();
}
cudaError_t launch_head_fwd(View x, const float* w, const float* b, int ncls, float* logits, float* probs,
                            cudaStream_t s) {
    const long long c8 = x.c / 8;
    if (x.c % 8 == 0 && c8 >= 1 && c8 <= 32 && (c8 & (c8 - 1)) == 0 && ncls <= 2) {
        // summation order differs from the scalar kernel only in the grouping of the fp32 partial dot products
        return ncls == 1 ? head_fwd_vec_launch<1>(x, w, b, logits, probs, s)
                         : head_fwd_vec_launch<2>(x, w, b, logits, probs, s);
    }
    launch_k(head_fwd_kernel, grid_for(x.voxels(), 256, 148, 16), 256, 0, s, x, w, b, ncls, logits, probs,
                                                                      x.d * x.h * x.w);
    return cudaGetLastError();
}

// dx[v, c] = sum_k dl[v,k] w[k,c] ; dw[k,c] += sum_v dl[v,k] x[v,c] ; db[k] += sum_v dl[v,k]
template <int NCLS>
__global__ void __launch_bounds__(256) head_bwd_kernel(View x, const float* __restrict__ w,
                                                       const float* __restrict__ dl, View dx, float* dw, float* db,
                                                       int c8, int rows, long long nvox, long long nvox_per_n) {
    pdl_wait();
    extern __shared__ float smem[];
    const int row = threadIdx.x / c8, cv = threadIdx.x - row * c8;
    const bool active = row < rows;
    const int c = (int)x.c;
    float acc[NCLS][8];
    float dbacc[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) {
        dbacc[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
    }
    if (active) {
        float wk[NCLS][8];
#pragma unroll
        for (int k = 0; k < NCLS; ++k) ldf8(w + k * c + cv * 8, wk[k]);
        for (long long v = (long long)blockIdx.x * rows + row; v < nvox; v += (long long)gridDim.x * rows) {
            const long long nb = v / nvox_per_n, sp = v - nb * nvox_per_n;
            float g[NCLS];
#pragma unroll
            for (int k = 0; k < NCLS; ++k) g[k] = __ldg(dl + (nb * NCLS + k) * nvox_per_n + sp);
            float a[8], o[8];
            unpack8(ld8_stream(x.p + v * x.ld + cv * 8), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
            for (int k = 0; k < NCLS; ++k) {
                if (cv == 0) dbacc[k] += g[k];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = fmaf(g[k], wk[k][j], o[j]);
                    acc[k][j] = fmaf(g[k], a[j], acc[k][j]);
                }
            }
            st8(dx.p + v * dx.ld + cv * 8, pack8(o));
        }
    }
    // reduce dw: smem [rows][c8][NCLS*8]
    if (active) {
#pragma unroll
        for (int k = 0; k < NCLS; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) smem[((row * c8 + cv) * NCLS + k) * 8 + j] = acc[k][j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < c * NCLS; t += blockDim.x) {
        const int k = t / c, ch = t - k * c;
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += smem[((r * c8 + (ch >> 3)) * NCLS + k) * 8 + (ch & 7)];
        atomicAdd(dw + k * c + ch, s);
    }
    __syncthreads();
    if (active && cv == 0) {
#pragma unroll
        for (int k = 0; k < NCLS; ++k) smem[row * NCLS + k] = dbacc[k];
    }
    __syncthreads();
    if (threadIdx.x < NCLS) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += smem[r * NCLS + threadIdx.x];
        atomicAdd(db + threadIdx.x, s);
    }
}
template <int NCLS>
static cudaError_t head_bwd_launch(View x, const float* w, const float* dl, View dx, float* dw, float* db,
                                   cudaStream_t s) {
    const LaneMap m = lane_map(x.c);
    const long long nvox = x.voxels();
    long long blocks = (nvox + m.rows * 8LL - 1) / (m.rows * 8LL);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)m.rows * m.c8 * NCLS * 8 * sizeof(float);
    launch_k(head_bwd_kernel<NCLS>, (int)blocks, 256, smem, s, x, w, dl, dx, dw, db, m.c8, m.rows, nvox, x.d * x.h * x.w);
    return cudaGetLastError();
}
cudaError_t launch_head_bwd(View x, const float* w, int ncls, const float* dlogits, View dx, float* dw, float* db,
                            cudaStream_t s) {
    switch (ncls) {
        case 1: return head_bwd_launch<1>(x, w, dlogits, dx, dw, db, s);
        case 2: return head_bwd_launch<2>(x, w, dlogits, dx, dw, db, s);
        case 3: return head_bwd_launch<3>(x, w, dlogits, dx, dw, db, s);
        case 4: return head_bwd_launch<4>(x, w, dlogits, dx, dw, db, s);
        default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------ fused BatchNorm passes
// The BatchNorm passes are bound by HBM, so they get cheaper only by moving fewer bytes:
//   * forward, encoder blocks: BatchNorm + ReLU + MaxPool3d(2) in one pass (the pooled tensor comes from registers).
//   * backward, last block (up4): the gradient entering its last BatchNorm is dout = dlogits . w_head, an elementwise
//     function of a 1/64-size tensor, so it is never written: both backward passes recompute it.  head_bwd + reduce +
//     apply move 7 tensor-sized streams (a -> dout; dout, y; dout, y -> dy); the fused passes move 3 (y; y -> dy) and
//     produce the head's dw / db on the way (a = relu(bn(y)) is recomputed with bn_apply_relu_kernel's arithmetic).
//     Every intermediate is rounded exactly where the unfused kernels round it (dout to bf16, a to bf16).
//   (The same treatment of the encoder blocks — dout = dskip + MaxPool3d-backward(dpool) recomputed in both passes, 5
//    streams instead of 8 — was built and measured: 0.93 ms against 0.79 ms for maxpool_bwd + bn_bwd at 2 x 128^3 x 64.
//    Seventeen 16-byte loads per 2x2x2 cell and thread cost 170 registers, i.e. 8 resident warps per SM, and the
//    recomputed pooling decisions ~3 instructions per byte: bound by issue at low occupancy, not by HBM.  Dropped.)

DEV void cell_decode(uint32_t cell, const FastDiv& cwd, const FastDiv& chd, const FastDiv& cdd, uint32_t& cw,
                     uint32_t& ch, uint32_t& cd, uint32_t& nb) {
    uint32_t q = cwd.quot(cell);
    cw = cell - q * cwd.div;
    uint32_t r = q;
    q = chd.quot(r);
    ch = r - q * chd.div;
    r = q;
    q = cdd.quot(r);
    cd = r - q * cdd.div;
    nb = q;
}
DEV Bf8 bn_relu8(const Bf8& yb, const float (&sc)[8], const float (&sh)[8]) {
    float f[8];
    unpack8(yb, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
    return pack8(f);
}

// one thread = one 2x2x2 cell (cells cover ceil(extent / 2)) x 8 channels
__global__ void __launch_bounds__(256) bn_apply_relu_pool_kernel(View y, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, View out,
                                                                 View pooled, FastDiv c8d, FastDiv cwd, FastDiv chd,
                                                                 FastDiv cdd, uint32_t total) {
    pdl_wait();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t cell = c8d.quot(i);
        const uint32_t c0 = (i - cell * c8d.div) * 8;
        uint32_t cw, ch, cd, nb;
        cell_decode(cell, cwd, chd, cdd, cw, ch, cd, nb);
        const bool full = (2 * cd + 1 < y.d) && (2 * ch + 1 < y.h) && (2 * cw + 1 < y.w);
        const long long v000 = (((long long)nb * y.d + 2 * cd) * y.h + 2 * ch) * y.w + 2 * cw;
        float sc[8], sh[8];
        ldf8(scale + c0, sc);
        ldf8(shift + c0, sh);
        if (full) {
            Bf8 xin[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long v = v000 + ((k >> 2) & 1) * y.h * y.w + ((k >> 1) & 1) * y.w + (k & 1);
                xin[k] = ld8_stream(y.p + v * y.ld + c0);
            }
            Bf8 m;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long v = v000 + ((k >> 2) & 1) * y.h * y.w + ((k >> 1) & 1) * y.w + (k & 1);
                const Bf8 a = bn_relu8(xin[k], sc, sh);
                st8(out.p + v * out.ld + c0, a);
                if (k == 0) m = a;
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) m.v[j] = max2_nan(m.v[j], a.v[j]);
                }
            }
            const long long ov = (((long long)nb * pooled.d + cd) * pooled.h + ch) * pooled.w + cw;
            st8(pooled.p + ov * pooled.ld + c0, m);
        } else {   // odd extent: the cell sticks out of the volume, its voxels belong to no pooling window
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t kd = (k >> 2) & 1, kh = (k >> 1) & 1, kw = k & 1;
                if (2 * cd + kd < y.d && 2 * ch + kh < y.h && 2 * cw + kw < y.w) {
                    const long long v = v000 + (long long)kd * y.h * y.w + kh * y.w + kw;
                    st8(out.p + v * out.ld + c0, bn_relu8(ld8_stream(y.p + v * y.ld + c0), sc, sh));
                }
            }
        }
    }
}
cudaError_t launch_bn_apply_relu_pool(View y, const float* scale, const float* shift, View out, View pooled, int sms,
                                      cudaStream_t s) {
    const long long cw = (y.w + 1) / 2, ch = (y.h + 1) / 2, cd = (y.d + 1) / 2;
    const long long total = y.n * cd * ch * cw * (y.c / 8);
    if (total >= (1LL << 31)) return cudaErrorInvalidValue;
    if (total == 0) return cudaSuccess;
    launch_k(bn_apply_relu_pool_kernel, grid_for(total, 256, sms, 16), 256, 0, s, 
        y, scale, shift, out, pooled, FastDiv((uint32_t)(y.c / 8)), FastDiv((uint32_t)cw), FastDiv((uint32_t)ch),
        FastDiv((uint32_t)cd), (uint32_t)total);
    return cudaGetLastError();
}

// BatchNorm backward with dout = dlogits . w_head recomputed on the fly (the network's last BatchNorm).  Pass 1 also
// accumulates the head's weight / bias gradients from a = relu(bn(y)) recomputed the same way.  Software-pipelined like
// bn_bwd_*_kernel (the group carries the voxels' dlogits too).  Pass 1 reads ONE stream and does ~14 instructions per
// element, so it is bound by instruction issue, not by HBM (ncu: 71 % issue, 32 % DRAM before the diet below).
constexpr int kHeadU = 4;
template <int NCLS>
struct HeadGroup {
    Bf8 x[kHeadU];
    float g[kHeadU][NCLS];
};
template <int NCLS>
DEV void head_load(HeadGroup<NCLS>& t, const float* __restrict__ dl, const View& y, long long v0, long long step,
                   long long nvox, long long nvox_per_n, int c0) {
#pragma unroll
    for (int u = 0; u < kHeadU; ++u) {
        const long long v = v0 + u * step;
        if (v < nvox) {
            t.x[u] = ld8_stream(y.p + v * y.ld + c0);
            if (NCLS == 1) {
                t.g[u][0] = __ldg(dl + v);   // (N, 1, D, H, W) is the voxel order itself
            } else {
                const long long nb = v / nvox_per_n, sp = v - nb * nvox_per_n;
#pragma unroll
                for (int k = 0; k < NCLS; ++k) t.g[u][k] = __ldg(dl + (nb * NCLS + k) * nvox_per_n + sp);
            }
        }
    }
}
template <int NCLS, bool APPLY>
__global__ void __launch_bounds__(256, 2) bn_bwd_head_kernel(const float* __restrict__ dl, const float* __restrict__ w,
                                                             View y, const float* __restrict__ scale,
                                                             const float* __restrict__ shift,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd,
                                                             const float* __restrict__ coef, float* partial, View dy,
                                                             float* dbias, float* dw, float* db, int c8, int rows,
                                                             long long nvox, long long nvox_per_n) {
    pdl_wait();
    extern __shared__ float smem[];
    const int row = threadIdx.x / c8, cv = threadIdx.x - row * c8;
    const bool active = row < rows;
    const int c = (int)y.c;
    constexpr int NACC = APPLY ? 1 : 2;
    constexpr int NW = APPLY ? 1 : NCLS;
    float acc[NACC][8], wacc[NW][8], dbacc[NW];
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        dbacc[k] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) wacc[k][j] = 0.f;
    }
    if (active) {
        const int c0 = cv * 8;
        float sc[8], sh[8], A[8], B[8], wk[NCLS][8];   // reduce: A = rstd, B = -mean * rstd; apply: A = P, B = Q
        ldf8(shift + c0, sh);
#pragma unroll
        for (int k = 0; k < NCLS; ++k) ldf8(w + k * c + c0, wk[k]);
        if (APPLY) {
            bwd_apply_consts(scale, mean, rstd, coef, cv, sc, A, B);
        } else {
            ldf8(scale + c0, sc);
            ldf8(rstd + c0, A);
            ldf8(mean + c0, B);
#pragma unroll
            for (int j = 0; j < 8; ++j) B[j] = -B[j] * A[j];
        }
        const long long step = (long long)gridDim.x * rows;
        auto compute = [&](const HeadGroup<NCLS>& t, long long v0) {
#pragma unroll
            for (int u = 0; u < kHeadU; ++u) {
                const long long v = v0 + u * step;
                if (v >= nvox) continue;
                float x[8], o[8], gq[8], z[8];
                unpack8(t.x[u], x);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = 0.f;
                    z[j] = fmaf(x[j], sc[j], sh[j]);
                }
#pragma unroll
                for (int k = 0; k < NCLS; ++k)
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = fmaf(t.g[u][k], wk[k][j], o[j]);
                unpack8(pack8(o), gq);   // dout of this voxel, rounded to bf16 as head_bwd_kernel stores it
                if (APPLY) {
                    float r[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        r[j] = fmaf(sc[j], z[j] > 0.f ? gq[j] : 0.f, fmaf(x[j], A[j], B[j]));
                    const Bf8 ob = pack8(r);
                    st8(dy.p + v * dy.ld + c0, ob);
                    unpack8(ob, r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[0][j] += r[j];
                } else {
                    float a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = fmaxf(z[j], 0.f);
                    unpack8(pack8(a), a);   // the head's input as bn_apply_relu_kernel stored it
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float gm = z[j] > 0.f ? gq[j] : 0.f;
                        acc[0][j] += gm;
                        acc[NACC - 1][j] = fmaf(gm, fmaf(x[j], A[j], B[j]), acc[NACC - 1][j]);
                    }
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        if (cv == 0) dbacc[k] += t.g[u][k];
#pragma unroll
                        for (int j = 0; j < 8; ++j) wacc[k][j] = fmaf(t.g[u][k], a[j], wacc[k][j]);
                    }
                }
            }
        };
        HeadGroup<NCLS> ga, gb;
        long long v0 = (long long)blockIdx.x * rows + row;
        const long long big = kHeadU * step;
        if (v0 < nvox) head_load<NCLS>(ga, dl, y, v0, step, nvox, nvox_per_n, c0);
        while (v0 < nvox) {
            const long long v1 = v0 + big;
            if (v1 < nvox) head_load<NCLS>(gb, dl, y, v1, step, nvox, nvox_per_n, c0);
            compute(ga, v0);
            if (v1 >= nvox) break;
            const long long v2 = v1 + big;
            if (v2 < nvox) head_load<NCLS>(ga, dl, y, v2, step, nvox, nvox_per_n, c0);
            compute(gb, v1);
            v0 = v2;
        }
    }
    if (APPLY) {
        if (dbias) block_reduce_store<NACC>(acc, c8, rows, row, cv, active, smem, dbias, (int)y.c, true);
        return;
    }
    block_reduce_store<NACC>(acc, c8, rows, row, cv, active, smem, partial + (long long)blockIdx.x * y.c * NACC,
                             (int)y.c, false);
    __syncthreads();
    // head dw[k][c] / db[k]: smem [rows][c8][NW * 8], as head_bwd_kernel
    if (active) {
#pragma unroll
        for (int k = 0; k < NW; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) smem[((row * c8 + cv) * NW + k) * 8 + j] = wacc[k][j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < c * NW; t += blockDim.x) {
        const int k = t / c, ch = t - k * c;
        float sum = 0.f;
        for (int r = 0; r < rows; ++r) sum += smem[((r * c8 + (ch >> 3)) * NW + k) * 8 + (ch & 7)];
        atomicAdd(dw + k * c + ch, sum);
    }
    __syncthreads();
    if (active && cv == 0) {
#pragma unroll
        for (int k = 0; k < NW; ++k) smem[row * NW + k] = dbacc[k];
    }
    __syncthreads();
    if (threadIdx.x < NW) {
        float sum = 0.f;
        for (int r = 0; r < rows; ++r) sum += smem[r * NW + threadIdx.x];
        atomicAdd(db + threadIdx.x, sum);
    }
}
template <int NCLS>
static cudaError_t bn_bwd_head_launch(bool apply, const float* dl, const float* w, View y, const float* scale,
                                      const float* shift, const float* mean, const float* rstd, const float* coef,
                                      float* partial, int* nblk, View dy, float* dbias, float* dw, float* db,
                                      cudaStream_t s) {
    const LaneMap m = lane_map(y.c);
    const long long nvox = y.voxels();
    long long blocks = (nvox + m.rows * 8LL - 1) / (m.rows * 8LL);
    const long long cap = apply ? 148 * 8 : kBwdMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (apply) {
        launch_k(bn_bwd_head_kernel<NCLS, true>, (int)blocks, 256, reduce_smem_bytes(m, 1), s, 
            dl, w, y, scale, shift, mean, rstd, coef, nullptr, dy, dbias, nullptr, nullptr, m.c8, m.rows, nvox,
            y.d * y.h * y.w);
    } else {
        *nblk = (int)blocks;
        size_t smem = reduce_smem_bytes(m, 2);
        const size_t head = (size_t)m.rows * m.c8 * NCLS * 8 * sizeof(float);
        if (head > smem) smem = head;
        launch_k(bn_bwd_head_kernel<NCLS, false>, (int)blocks, 256, smem, s, dl, w, y, scale, shift, mean, rstd, nullptr,
                                                                      partial, y, nullptr, dw, db, m.c8, m.rows, nvox,
                                                                      y.d * y.h * y.w);
    }
    return cudaGetLastError();
}
cudaError_t launch_bn_bwd_head(bool apply, const float* dlogits, const float* w, int ncls, View y, const float* scale,
                               const float* shift, const float* mean, const float* rstd, const float* coef,
                               float* partial, int* nblk, View dy, float* dbias, float* dw, float* db,
                               cudaStream_t s) {
#define B200_HEAD_CASE(N)                                                                                            \
    case N:                                                                                                          \
        return bn_bwd_head_launch<N>(apply, dlogits, w, y, scale, shift, mean, rstd, coef, partial, nblk, dy, dbias, \
                                     dw, db, s)
    switch (ncls) {
        B200_HEAD_CASE(1);
        B200_HEAD_CASE(2);
        B200_HEAD_CASE(3);
        B200_HEAD_CASE(4);
        default: return cudaErrorInvalidValue;
    }
#undef B200_HEAD_CASE
}

// ------------------------------------------------------------------------------------------------ loss
DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEV float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

__global__ void __launch_bounds__(256) loss_partial_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                           long long n, float* ws) {
    pdl_wait();
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z) + i);
        const float4 tt = __ldg(reinterpret_cast<const float4*>(t) + i);
        const float zs[4] = {zz.x, zz.y, zz.z, zz.w}, ts[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float sg = sigmoidf_(zs[j]);
            s[0] += fmaxf(zs[j], 0.f) - zs[j] * ts[j] + log1pf(expf(-fabsf(zs[j])));
            s[1] += sg * ts[j];
            s[2] += sg;
            s[3] += ts[j];
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const float zj = z[i], tj = t[i], sg = sigmoidf_(zj);
            s[0] += fmaxf(zj, 0.f) - zj * tj + log1pf(expf(-fabsf(zj)));
            s[1] += sg * tj;
            s[2] += sg;
            s[3] += tj;
        }
    }
    __shared__ float red[8][4];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v = warp_sum(s[j]);
        if (lane == 0) red[wrp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
        ws[blockIdx.x * 4 + threadIdx.x] = v;
    }
}
__global__ void loss_finalize_kernel(const float* __restrict__ ws, int nblk, double inv_n, float bce_w, float dice_w,
                                     float smooth, float* sums, float* loss) {
    pdl_wait();
    __shared__ double red[4][4];
    const int j = threadIdx.x & 3, part = threadIdx.x >> 2;  // 16 threads: 4 sums x 4 parts
    double a = 0.0;
    for (int b = part; b < nblk; b += 4) a += ws[b * 4 + j];
    red[part][j] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s[4];
        for (int k = 0; k < 4; ++k) s[k] = red[0][k] + red[1][k] + red[2][k] + red[3][k];
        for (int k = 0; k < 4; ++k) sums[k] = (float)s[k];
        const double dice = (2.0 * s[1] + smooth) / (s[2] + s[3] + smooth);
        loss[0] = (float)(bce_w * s[0] * inv_n + dice_w * (1.0 - dice));
    }
}
cudaError_t launch_loss_fwd(const float* z, const float* t, long long n, float bce_w, float dice_w, float smooth,
                            float* ws, float* sums, float* loss, int sms, cudaStream_t s) {
    int blocks = grid_for(n / 4 + 1, 256, sms, 4);
    if (blocks > kLossMaxBlocks) blocks = kLossMaxBlocks;
    launch_k(loss_partial_kernel, blocks, 256, 0, s, z, t, n, ws);
    launch_k(loss_finalize_kernel, 1, 16, 0, s, ws, blocks, 1.0 / (double)n, bce_w, dice_w, smooth, sums, loss);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                       long long n, float bce_w, float dice_w, float smooth,
                                                       const float* __restrict__ sums, const float* __restrict__ gout,
                                                       float* dz) {
    pdl_wait();
    const float g = gout[0];
    const float I = sums[1], P = sums[2], T = sums[3];
    const float B = P + T + smooth, A = 2.f * I + smooth;
    const float kb = g * bce_w / (float)n;
    const float k1 = g * dice_w * (-2.f / B), k2 = g * dice_w * (A / (B * B));
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float zj = __ldg(z + i), tj = __ldg(t + i);
        const float sg = sigmoidf_(zj);
        dz[i] = kb * (sg - tj) + sg * (1.f - sg) * (k1 * tj + k2);
    }
}
cudaError_t launch_loss_bwd(const float* z, const float* t, long long n, float bce_w, float dice_w, float smooth,
                            const float* sums, const float* gout, float* dz, int sms, cudaStream_t s) {
    launch_k(loss_bwd_kernel, grid_for(n, 256, sms, 8), 256, 0, s, z, t, n, bce_w, dice_w, smooth, sums, gout, dz);
    return cudaGetLastError();
}

DEV uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, long long n,
                                                        __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    const long long n8 = n >> 3;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
        uint4 o;
        o.x = pack2(a.x, a.y); o.y = pack2(a.z, a.w); o.z = pack2(b.x, b.y); o.w = pack2(b.z, b.w);
        reinterpret_cast<uint4*>(out)[i] = o;
    }
    if (blockIdx.x == 0)
        for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) out[i] = __float2bfloat16_rn(x[i]);
}
cudaError_t launch_cast_bf16(const float* x, long long n, __nv_bfloat16* out, int sms, cudaStream_t s) {
    launch_k(cast_bf16_kernel, grid_for(n / 8 + 1, 256, sms, 8), 256, 0, s, x, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float step_size, float b2, float w1, float w2, float eps,
                                                   float wd, float bc2_sqrt, float gscale, const float* found_inf,
                                                   __nv_bfloat16* __restrict__ shadow, const float* __restrict__ dyn) {
    pdl_wait();
    // torch.optim.Adam arithmetic: m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); p.addcdiv_(m, sqrt(v)/bc2+eps)
    if (found_inf && *found_inf != 0.f) return;
    if (dyn) {   // step-dependent scalars from device memory: the launch can be replayed from a CUDA graph
        step_size = __ldg(dyn);
        bc2_sqrt = __ldg(dyn + 1);
        gscale = __ldg(dyn + 2);
    }
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float* pa = reinterpret_cast<float*>(&pp);
        const float* ga = reinterpret_cast<const float*>(&gg);
        float* ma = reinterpret_cast<float*>(&mm);
        float* va = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = fmaf(wd, pa[j], ga[j] * gscale);
            ma[j] = fmaf(w1, gr - ma[j], ma[j]);
            va[j] = fmaf(w2 * gr, gr, b2 * va[j]);
            const float denom = sqrtf(va[j]) / bc2_sqrt + eps;
            pa[j] -= step_size * (ma[j] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow) {  // bf16 operand shadow of the weights, same flat index
            uint2 sv;
            sv.x = pack2(pa[0], pa[1]);
            sv.y = pack2(pa[2], pa[3]);
            reinterpret_cast<uint2*>(shadow)[i] = sv;
        }
    }
    if (blockIdx.x == 0) {
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const float gr = fmaf(wd, p[i], g[i] * gscale);
            m[i] = fmaf(w1, gr - m[i], m[i]);
            v[i] = fmaf(w2 * gr, gr, b2 * v[i]);
            const float denom = sqrtf(v[i]) / bc2_sqrt + eps;
            p[i] -= step_size * (m[i] / denom);
            if (shadow) shadow[i] = __float2bfloat16_rn(p[i]);
        }
    }
}
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double b1, double b2,
                        double eps, double wd, long long step, double gscale, const float* found_inf,
                        __nv_bfloat16* shadow, const float* dyn, int sms, cudaStream_t s) {
    const double bc1 = 1.0 - pow(b1, (double)step);
    const double bc2 = 1.0 - pow(b2, (double)step);
    launch_k(adam_kernel, grid_for(n / 4 + 1, 256, sms, 8), 256, 0, s, p, g, m, v, n, (float)(lr / bc1), (float)b2,
                                                                (float)(1.0 - b1), (float)(1.0 - b2), (float)eps,
                                                                (float)wd, (float)sqrt(bc2), (float)gscale, found_inf, shadow, dyn);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* out) {
    pdl_wait();
    float s = 0.f;
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        s = fmaf(v, v, s);
        bad |= !isfinite(v);
    }
    s = warp_sum(s);
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (lane == 0) red[wrp] = s;
    if (__any_sync(0xffffffffu, bad) && lane == 0) out[1] = 1.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int k = 0; k < 8; ++k) a += red[k];
        atomicAdd(out, a);
    }
}
cudaError_t launch_sumsq(const float* x, long long n, float* out, int sms, cudaStream_t s) {
    launch_k(sumsq_kernel, grid_for(n, 256, sms, 4), 256, 0, s, x, n, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ helpers
__global__ void __launch_bounds__(256) fill_zero_kernel(View v, FastDiv c8d, uint32_t total) {
    pdl_wait();
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t vox = c8d.quot(i);
        const uint32_t c0 = (i - vox * c8d.div) * 8;
        *reinterpret_cast<uint4*>(v.p + (long long)vox * v.ld + c0) = z;
    }
}
cudaError_t launch_fill_zero(View v, int sms, cudaStream_t s) {
    const long long total = v.voxels() * (v.c / 8);
    if (total >= (1LL << 31)) return cudaErrorInvalidValue;
    if (total == 0) return cudaSuccess;
    launch_k(fill_zero_kernel, grid_for(total, 256, sms, 16), 256, 0, s, v, FastDiv((uint32_t)(v.c / 8)), (uint32_t)total);
    return cudaGetLastError();
}

__global__ void unpack_act_kernel(View v, float* out, long long nvox_per_n) {
    pdl_wait();
    const long long total = v.voxels() * v.c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        // i indexes the NCDHW output
        const long long sp = i % nvox_per_n;
        const long long r = i / nvox_per_n;
        const long long ch = r % v.c, nb = r / v.c;
        out[i] = __bfloat162float(v.p[(nb * nvox_per_n + sp) * v.ld + ch]);
    }
}
cudaError_t launch_unpack_act(View v, float* out, cudaStream_t s) {
    launch_k(unpack_act_kernel, grid_for(v.voxels() * v.c, 256, 148, 16), 256, 0, s, v, out, v.d * v.h * v.w);
    return cudaGetLastError();
}


// ================================================================================================ "next" rows (8f)
// Input pipeline and validation metrics on the device.  Reference call sites: resample-to-target of every modality
// (script/data_loader.py:240-283, sitk.ResampleImageFilter, linear) and of the label (:395-409, nearest, then > 0),
// per-modality min-max normalisation (script/predict.py:69-75), hard Dice / IoU of thresholded predictions
// (script/validate_model.py:24-95).

// ITK semantics of that resample (identity direction, same origin, output spacing = in_size*spacing/out_size): output
// index i maps to the continuous input index i * (in/out) per axis (pixel centres at integer indices).  A point is
// inside the buffer iff index < size - 0.5 on every axis, else the default pixel value 0 is written.  Linear: base =
// floor, the upper neighbour is clamped to the last index.  Nearest: floor(index + 0.5).
__global__ void __launch_bounds__(256) resample3d_kernel(const float* __restrict__ in, int di, int hi, int wi,
                                                         float* __restrict__ out, int dout, int ho, int wo,
                                                         long long nvol, int nearest, int binarize) {
    pdl_wait();
    const double sd = (double)di / dout, sh = (double)hi / ho, sw = (double)wi / wo;
    const long long per = (long long)dout * ho * wo, total = nvol * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long vol = i / per;
        long long r = i - vol * per;
        const int od = (int)(r / ((long long)ho * wo));
        r -= (long long)od * ho * wo;
        const int oh = (int)(r / wo), ow = (int)(r - (long long)oh * wo);
        const double cd = od * sd, ch = oh * sh, cw = ow * sw;
        const float* src = in + vol * (long long)di * hi * wi;
        float v = 0.f;
        if (cd < di - 0.5 && ch < hi - 0.5 && cw < wi - 0.5) {
            if (nearest) {
                const int zd = min((int)floor(cd + 0.5), di - 1), zh = min((int)floor(ch + 0.5), hi - 1),
                          zw = min((int)floor(cw + 0.5), wi - 1);
                v = __ldg(src + ((long long)zd * hi + zh) * wi + zw);
            } else {
                const int d0 = (int)floor(cd), h0 = (int)floor(ch), w0 = (int)floor(cw);
                const int d1 = min(d0 + 1, di - 1), h1 = min(h0 + 1, hi - 1), w1 = min(w0 + 1, wi - 1);
                const float fd = (float)(cd - d0), fh = (float)(ch - h0), fw = (float)(cw - w0);
                auto at = [&](int z, int y, int x) { return __ldg(src + ((long long)z * hi + y) * wi + x); };
                const float c00 = at(d0, h0, w0) + fw * (at(d0, h0, w1) - at(d0, h0, w0));
                const float c01 = at(d0, h1, w0) + fw * (at(d0, h1, w1) - at(d0, h1, w0));
                const float c10 = at(d1, h0, w0) + fw * (at(d1, h0, w1) - at(d1, h0, w0));
                const float c11 = at(d1, h1, w0) + fw * (at(d1, h1, w1) - at(d1, h1, w0));
                const float c0 = c00 + fh * (c01 - c00), c1 = c10 + fh * (c11 - c10);
                v = c0 + fd * (c1 - c0);
            }
        }
        out[i] = binarize ? (v > 0.f ? 1.f : 0.f) : v;
    }
}
cudaError_t launch_resample3d(const float* in, long long nvol, int di, int hi, int wi, float* out, int dout, int ho,
                              int wo, int nearest, int binarize, int sms, cudaStream_t s) {
    const long long total = nvol * dout * ho * wo;
    if (total == 0) return cudaSuccess;
    launch_k(resample3d_kernel, grid_for(total, 256, sms, 16), 256, 0, s, in, di, hi, wi, out, dout, ho, wo, nvol, nearest,
                                                                   binarize);
    return cudaGetLastError();
}

// per-volume min / max (order-preserving unsigned keys so that one atomicMin / atomicMax per block suffices)
DEV uint32_t f2key(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
DEV float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__global__ void minmax_init_kernel(uint32_t* keys, int nvol) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nvol) { keys[2 * i] = 0xffffffffu; keys[2 * i + 1] = 0u; }
}
__global__ void __launch_bounds__(256) minmax_reduce_kernel(const float* __restrict__ x, long long per, uint32_t* keys) {
    pdl_wait();
    const int vol = blockIdx.y;
    const float* src = x + vol * per;
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(src + i);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys + 2 * vol, f2key(lo));
        atomicMax(keys + 2 * vol + 1, f2key(hi));
    }
}
__global__ void __launch_bounds__(256) minmax_apply_kernel(float* __restrict__ x, long long per,
                                                           const uint32_t* __restrict__ keys) {
    pdl_wait();
    const int vol = blockIdx.y;
    const float lo = key2f(keys[2 * vol]), hi = key2f(keys[2 * vol + 1]);
    const bool flat = !(hi > lo);          // constant volume -> zeros (predict.py's division guard)
    const float range = hi - lo;
    float* dst = x + vol * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x)
        dst[i] = flat ? 0.f : (dst[i] - lo) / range;
}
cudaError_t launch_minmax_normalize(float* x, long long nvol, long long per, uint32_t* keys, int sms, cudaStream_t s) {
    if (nvol == 0 || per == 0) return cudaSuccess;
    launch_k(minmax_init_kernel, (int)((nvol + 127) / 128), 128, 0, s, keys, (int)nvol);
    int bx = grid_for(per, 256, sms, 8);
    if (bx * nvol > (long long)sms * 16) bx = (int)((sms * 16 + nvol - 1) / nvol);
    dim3 grid(bx, (unsigned)nvol);
    launch_k(minmax_reduce_kernel, grid, 256, 0, s, x, per, keys);
    launch_k(minmax_apply_kernel, grid, 256, 0, s, x, per, keys);
    return cudaGetLastError();
}

// counts[sample][3] += (|pred & target|, |pred|, |target|) with pred = score > threshold, target = label > 0.5:
// exact integer arithmetic (the reference sums 0/1 floats, which is exact up to 2^24 only)
__global__ void __launch_bounds__(256) seg_counts_kernel(const float* __restrict__ score,
                                                         const float* __restrict__ label, long long per,
                                                         float threshold, unsigned long long* counts) {
    pdl_wait();
    const int smp = blockIdx.y;
    const float* sc = score + smp * per;
    const float* lb = label + smp * per;
    unsigned int inter = 0, np = 0, nt = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const bool p = __ldg(sc + i) > threshold, t = __ldg(lb + i) > 0.5f;
        inter += p && t;
        np += p;
        nt += t;
    }
    inter = __reduce_add_sync(0xffffffffu, inter);
    np = __reduce_add_sync(0xffffffffu, np);
    nt = __reduce_add_sync(0xffffffffu, nt);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counts + 3 * smp, (unsigned long long)inter);
        atomicAdd(counts + 3 * smp + 1, (unsigned long long)np);
        atomicAdd(counts + 3 * smp + 2, (unsigned long long)nt);
    }
}
cudaError_t launch_seg_counts(const float* score, const float* label, long long nsmp, long long per, float threshold,
                              unsigned long long* counts, int sms, cudaStream_t s) {
    if (nsmp == 0 || per == 0) return cudaSuccess;
    int bx = grid_for(per, 256, sms, 8);
    if (bx * nsmp > (long long)sms * 16) bx = (int)((sms * 16 + nsmp - 1) / nsmp);
    if (bx < 1) bx = 1;
    launch_k(seg_counts_kernel, dim3(bx, (unsigned)nsmp), 256, 0, s, score, label, per, threshold, counts);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ split-K finalize
// ws[split][voxel][c] holds the fp32 partial tiles of a deep-level convolution computed by `splits` work units per tile
// (igemm EPI_SPLITK).  One pass adds them in split order (deterministic) and writes the bf16 output — mode 0: plain
// (dgrad), 1: + bias and per-block BatchNorm partial sums of the ROUNDED values (train fprop; partial[block][c][2], the
// layout the conv epilogue writes), 2: relu(acc * scale + shift) (eval fprop).  The slices are L2-resident (<= 36 MB).
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int splits, View y, int mode,
                                                              const float* __restrict__ v0,
                                                              const float* __restrict__ v1, float* partial, int c8,
                                                              int rows, long long nvox) {
    pdl_wait();
    extern __shared__ float smem[];
    const int row = threadIdx.x / c8, cv = threadIdx.x - row * c8;
    const bool active = row < rows;
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
    if (active) {
        float a0[8], a1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
        if (mode != 0) ldf8(v0 + cv * 8, a0);
        if (mode == 2) ldf8(v1 + cv * 8, a1);
        const long long slice = nvox * y.c;
        for (long long v = (long long)blockIdx.x * rows + row; v < nvox; v += (long long)gridDim.x * rows) {
            const float* src = ws + v * y.c + cv * 8;
            float f[8];
            ldf8(src, f);
            for (int sp = 1; sp < splits; ++sp) {
                float t[8];
                ldf8(src + sp * slice, t);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += t[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (mode == 1) f[j] += a0[j];
                else if (mode == 2) f[j] = fmaxf(fmaf(f[j], a0[j], a1[j]), 0.f);
            }
            const Bf8 ob = pack8(f);
            st8(y.p + v * y.ld + cv * 8, ob);
            if (mode == 1) {
                float r[8];
                unpack8(ob, r);
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[0][j] += r[j]; acc[1][j] = fmaf(r[j], r[j], acc[1][j]); }
            }
        }
    }
    if (mode == 1)
        block_reduce_store<2>(acc, c8, rows, row, cv, active, smem, partial + (long long)blockIdx.x * y.c * 2, (int)y.c,
                              false);
}
int splitk_finalize_blocks(long long nvox, long long c) {
    const LaneMap m = lane_map(c);
    long long blocks = (nvox + m.rows - 1) / m.rows;
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
cudaError_t launch_splitk_finalize(const float* ws, int splits, View y, int mode, const float* v0, const float* v1,
                                   float* partial, cudaStream_t s) {
    const LaneMap m = lane_map(y.c);
    const long long nvox = y.voxels();
    const size_t smem = mode == 1 ? reduce_smem_bytes(m, 2) : 0;
    launch_k(splitk_finalize_kernel, splitk_finalize_blocks(nvox, y.c), 256, smem, s, ws, splits, y, mode, v0, v1, partial,
                                                                                m.c8, m.rows, nvox);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ sliding windows
// Sliding-window inference (BASELINE configs[3]; the reference's script/predict.py:152-172 predicts whole volumes, a
// window schedule is the new capability): windows (volume, d0, h0, w0) of a batch of fp32 volumes are gathered into
// one contiguous batch, their logits are summed back into the volume in schedule order, and the sums are divided by
// the number of covering windows (a product of per-axis counts), squashed and thresholded in one pass.
__global__ void __launch_bounds__(256) window_gather_kernel(const float* __restrict__ x, long long C, long long D,
                                                            long long H, long long W, const int* __restrict__ org,
                                                            int wd, int wh, int ww, float* __restrict__ out,
                                                            long long per_win) {
    pdl_wait();
    const int win = blockIdx.y;
    const int v = org[4 * win], d0 = org[4 * win + 1], h0 = org[4 * win + 2], w0 = org[4 * win + 3];
    const float* src = x + (long long)v * C * D * H * W;
    float* dst = out + (long long)win * per_win;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_win;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int l = (int)(r % ww); r /= ww;
        const int j = (int)(r % wh); r /= wh;
        const int k = (int)(r % wd); r /= wd;   // r = channel
        dst[i] = __ldg(src + ((r * D + d0 + k) * H + h0 + j) * W + w0 + l);
    }
}
cudaError_t launch_window_gather(const float* x, long long c, long long d, long long h, long long w, const int* org,
                                 int nwin, int wd, int wh, int ww, float* out, int sms, cudaStream_t s) {
    const long long per = c * wd * wh * ww;
    dim3 grid((unsigned)grid_for(per, 256, sms, 8), (unsigned)nwin);
    launch_k(window_gather_kernel, grid, 256, 0, s, x, c, d, h, w, org, wd, wh, ww, out, per);
    return cudaGetLastError();
}
// one thread per voxel of the volumes [v_lo, v_lo + v_cnt): the windows of the launch that cover it are added in
// schedule order (deterministic: no atomics although windows overlap)
__global__ void __launch_bounds__(256) window_accumulate_kernel(const float* __restrict__ lg, const int* __restrict__ org,
                                                                int nwin, long long K, int wd, int wh, int ww,
                                                                float* __restrict__ acc, long long D, long long H,
                                                                long long W, int v_lo, long long total) {
    pdl_wait();
    extern __shared__ int s_org[];
    for (int i = threadIdx.x; i < 4 * nwin; i += blockDim.x) s_org[i] = org[i];
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H); r /= H;
        const int d = (int)(r % D); r /= D;
        const int k = (int)(r % K); r /= K;
        const int v = v_lo + (int)r;
        float a = 0.f;
        bool hit = false;
        for (int q = 0; q < nwin; ++q) {
            const int dd = d - s_org[4 * q + 1], hh = h - s_org[4 * q + 2], wl = w - s_org[4 * q + 3];
            if (s_org[4 * q] == v && (unsigned)dd < (unsigned)wd && (unsigned)hh < (unsigned)wh &&
                (unsigned)wl < (unsigned)ww) {
                a += __ldg(lg + ((((long long)q * K + k) * wd + dd) * wh + hh) * ww + wl);
                hit = true;
            }
        }
        if (hit) {
            float* dst = acc + ((((long long)v * K + k) * D + d) * H + h) * W + w;
            *dst += a;
        }
    }
}
cudaError_t launch_window_accumulate(const float* lg, const int* org, int nwin, long long k, int wd, int wh, int ww,
                                     float* acc, long long d, long long h, long long w, int v_lo, int v_cnt, int sms,
                                     cudaStream_t s) {
    const long long total = (long long)v_cnt * k * d * h * w;
    launch_k(window_accumulate_kernel, grid_for(total, 256, sms, 16), 256, 4 * nwin * sizeof(int), s, 
        lg, org, nwin, k, wd, wh, ww, acc, d, h, w, v_lo, total);
    return cudaGetLastError();
}
__global__ void __launch_bounds__(256) window_finalize_kernel(float* __restrict__ acc, const int* __restrict__ cover,
                                                              long long D, long long H, long long W, float threshold,
                                                              float* __restrict__ probs, float* __restrict__ mask,
                                                              long long total) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H); r /= H;
        const int d = (int)(r % D);
        const float cnt = (float)(cover[d] * cover[D + h] * cover[D + H + w]);
        const float z = acc[i] / cnt;
        acc[i] = z;
        const float pr = 1.f / (1.f + __expf(-z));
        if (probs) probs[i] = pr;
        if (mask) mask[i] = pr > threshold ? 1.f : 0.f;
    }
}
cudaError_t launch_window_finalize(float* acc, const int* cover, long long nk, long long d, long long h, long long w,
                                   float threshold, float* probs, float* mask, int sms, cudaStream_t s) {
    const long long total = nk * d * h * w;
    launch_k(window_finalize_kernel, grid_for(total, 256, sms, 16), 256, 0, s, acc, cover, d, h, w, threshold, probs, mask,
                                                                        total);
    return cudaGetLastError();
}

}  // namespace b200
