// First layer of the 5-modality network, forward: Conv3d(5 -> Cout <= 64, 3x3x3, pad 1) read straight from the fp32
// (N, 5, D, H, W) input (models/unet3d.py:194 inc = DoubleConv3D(5, 64); nn.Conv3d call site models/unet3d.py:29).
//
// The layer is 0.3 % of the network's FLOPs and writes 128 B per voxel for 20 B read: its floor is HBM (0.09 ms at
// 2 x 128^3), not the tensor pipe.  The generic direct kernel (igemm_im2col5_kernel) rebuilt the whole 144-column
// im2col image of every 128-voxel brick (~4500 warp instructions per brick in the builders) and ran one epilogue
// warpgroup (~1.8 us per brick): 0.55 ms.  This kernel marches along depth instead, like dmarch.cu:
//
//   * a CTA owns an 8 w x 16 h brick column (or a depth segment of one).  For every INPUT slice z it builds one "slice
//     image" S(z): 45 (+3 zero) rows k = c*9 + kh*3 + kw, each row the 128 voxels of the brick shifted by (kh-1, kw-1)
//     — 12 KB, voxel-contiguous and 128B-swizzled, i.e. an MN-major A operand (two 64-voxel halves LBO apart).  An
//     input row of 10 floats is loaded ONCE and written to its 9 (kh, kw) places: 90 row tasks per slice, where the
//     generic kernel ran 720 tasks per brick.
//   * output slice d = sum over kd of S(d + kd - 1) x W[kd]  (W[kd]: [Cout][48] K-major, resident in shared memory for
//     the whole kernel): 9 MMAs of 128 x 64 x 16, every slice image feeds three output slices.
//   * two epilogue warpgroups take the output slices in turn (TMEM ring of eight 64-column slots), each with its own
//     staging tile and TMA store; BatchNorm partial sums stay in registers until the CTA ends.
//
// Warp roles (512 threads, 1 CTA / SM): warp 0 loads the weights (once), warp 1 issues the MMAs, warp 2 allocates
// TMEM, warps 4-11 are the two epilogue warpgroups, warps 12-13 / 14-15 build the even / odd slice images.
#include <cuda_bf16.h>
#include "igemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {

namespace {
struct C1Unit {
    int nb, w0, h0, ds, de, z0, z1;
};
}  // namespace

extern "C" __global__ void __launch_bounds__(kC1Threads, 1)
    conv1_march_kernel(const __grid_constant__ Conv1MarchParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    const uint32_t smem_img = smem_base;                               // kC1Imgs slice images
    const uint32_t smem_b = smem_img + kC1Imgs * kC1ImgBytes;          // W[kd]: 3 x [64 rows][128 B]
    const uint32_t smem_c = smem_b + 3 * 8192;                         // one 16 KB staging tile per epilogue warpgroup
    const uint32_t bar_base = smem_c + kC1EpiWGs * kBoxBytes;
    auto ifull = [&](uint32_t s) { return bar_base + 8 * s; };
    auto iempty = [&](uint32_t s) { return bar_base + 8 * (kC1Imgs + s); };
    auto tfull = [&](uint32_t s) { return bar_base + 8 * (2 * kC1Imgs + s); };
    auto tempty = [&](uint32_t s) { return bar_base + 8 * (2 * kC1Imgs + kC1Slots + s); };
    const uint32_t bfull = bar_base + 8 * (2 * kC1Imgs + 2 * kC1Slots);
    const uint32_t tmem_ptr_smem = bfull + 8;
    const uint32_t f_off = (tmem_ptr_smem + 16 - smem_base + 15u) & ~15u;
    float* red = reinterpret_cast<float*>(smem_gen + f_off);   // [4 * kC1EpiWGs epilogue warps][64 cols][2]
    float* colvec = red + 4 * kC1EpiWGs * 64 * 2;                          // [2][64]: bias | scale, shift

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.b_map);
        prefetch_tmap(&p.c_map);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < kC1Imgs; ++s) { mbar_init(ifull(s), 2); mbar_init(iempty(s), 1); }   // 2 builder warps
        for (uint32_t s = 0; s < kC1Slots; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 128); }
        mbar_init(bfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    if (warp >= kC1BuilderWarp0) {
        // rows 45..47 of every slice image are never written again: zero (the MMAs read them against zero weights)
        for (int i = threadIdx.x - 32 * kC1BuilderWarp0; i < kC1Imgs * 2 * 3 * 8; i += 128) {
            const int chunk = i & 7, row = 45 + (i >> 3) % 3, half = (i / 24) & 1, img = i / 48;
            st_shared_v4(smem_img + img * kC1ImgBytes + half * kC1HalfBytes + row * 128 + (chunk << 4), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();   // launch.cuh: the predecessor's results are complete and visible from here on
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    const int units = p.nbatch * p.nbw * p.nbh * p.nseg;
    const int unit0 = blockIdx.x, unit_stride = gridDim.x;
    // unit -> (batch, brick column, depth segment); the segments of one column are consecutive units
    auto decode = [&](int unit) {
        C1Unit u;
        int col = unit / p.nseg;
        const int seg = unit - col * p.nseg;
        const int bw = col % p.nbw; col /= p.nbw;
        const int bh = col % p.nbh; col /= p.nbh;
        u.nb = col;
        u.w0 = bw * 8;
        u.h0 = bh * 16;
        u.ds = seg * p.seg_len;
        u.de = min(p.D, u.ds + p.seg_len);
        u.z0 = max(u.ds - 1, 0);          // input slices this unit reads: z0 .. z1
        u.z1 = min(u.de, p.D - 1);
        return u;
    };

    if (warp == 0) {
        // ===================================================================== weights: loaded once, stay resident
        if (elect_one()) {
            mbar_arrive_expect_tx(bfull, 3 * 8192);
            for (int kd = 0; kd < 3; ++kd) tma_load_3d(smem_b + kd * 8192, &p.b_map, bfull, 0, 0, kd);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // A: slice image, MN-major (M = voxels): 64-voxel halves LBO = 6 KB apart, 8-row K groups SBO = 1 KB apart, one
        // K step = 16 rows = 2 KB.  B: W[kd] K-major, one K step = 32 B inside the 128-byte swizzle row.
        const uint32_t idesc = make_idesc_bf16(128, 64, 1u, 0u);
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_img, kC1HalfBytes, 1024);
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_b, 0, 1024);
        mbar_wait(bfull, 0);
        uint32_t ibase = 0;    // slice images of the units before this one
        uint32_t iready = 0;   // slice images known to be complete
        uint32_t ucnt = 0;     // output slices so far
        for (int unit = unit0; unit < units; unit += unit_stride) {
            const C1Unit u = decode(unit);
            for (int d = u.ds; d < u.de; ++d, ++ucnt) {
                const uint32_t slot = ucnt % kC1Slots;
                mbar_wait(tempty(slot), ((ucnt / kC1Slots) & 1) ^ 1);   // previous use drained
                const uint32_t need = ibase + (uint32_t)(min(d + 1, u.z1) - u.z0);
                while (iready <= need) {
                    mbar_wait(ifull(iready % kC1Imgs), (iready / kC1Imgs) & 1);
                    ++iready;
                }
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d_tmem = tmem_base + slot * 64;
                    uint32_t accum = 0;
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        const int z = d + kd - 1;
                        if (z < 0 || z >= p.D) continue;   // the convolution's zero padding along depth
                        const uint32_t img = (ibase + (uint32_t)(z - u.z0)) % kC1Imgs;
                        const uint64_t a_desc = a_desc0 + img * (kC1ImgBytes >> 4);
                        const uint64_t b_desc = b_desc0 + kd * (8192 >> 4);
#pragma unroll
                        for (int s = 0; s < kC1Rows / 16; ++s) {
                            if (B200_ABLATE(p) != 3) umma_f16(d_tmem, a_desc + s * (2048 >> 4), b_desc + s * 2, idesc, accum);
                            accum = 1u;
                        }
                    }
                    umma_commit(tfull(slot));
                    // slice images nobody reads any more go back to the builders
                    if (d - 1 >= u.z0) umma_commit(iempty((ibase + (uint32_t)(d - 1 - u.z0)) % kC1Imgs));
                    if (d == u.de - 1)
                        for (int z = max(u.z0, d); z <= u.z1; ++z)
                            umma_commit(iempty((ibase + (uint32_t)(z - u.z0)) % kC1Imgs));
                }
                __syncwarp();
            }
            ibase += (uint32_t)(u.z1 - u.z0 + 1);
        }
    } else if (warp >= kC1BuilderWarp0) {
        // ===================================================================== slice-image builders
        // group g (two warps, 64 threads) builds the images with sequence number = g mod 2.  One task = one input row
        // (channel c, row hs = h0 - 1 + hr, hr = 0..17): 10 floats x[w0 - 1 .. w0 + 8] -> the chunks (8 voxels of
        // output row hr - kh) of the 9 image rows (c, kh, kw).  The loads of a group's next image are in flight while
        // it writes the current one (two register sets, ping-pong).
        const int g = (warp - kC1BuilderWarp0) >> 1;
        const int gt = threadIdx.x - 32 * kC1BuilderWarp0 - 64 * g;   // 0..63
        const long long hw = (long long)p.H * p.W;
        const bool vec_ok = (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
        struct Rows {
            float f[2][10];
        };
        struct Cursor {
            int unit, z;
            C1Unit u;
        };
        auto start = [&](Cursor& c, int unit) -> bool {
            if (unit >= units) return false;
            c.unit = unit;
            c.u = decode(unit);
            c.z = c.u.z0;
            return true;
        };
        auto advance = [&](Cursor& c) -> bool {
            if (c.z < c.u.z1) { ++c.z; return true; }
            return start(c, c.unit + unit_stride);
        };
        auto load = [&](Rows& r, const Cursor& c) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float (&f)[10] = r.f[i];
#pragma unroll
                for (int e = 0; e < 10; ++e) f[e] = 0.f;
                const int t = gt + 64 * i;
                if (t >= 18 * kC1Cin || B200_ABLATE(p) == 2) continue;
                const int ch = t / 18, hr = t - 18 * ch;
                const int hs = c.u.h0 - 1 + hr, w0 = c.u.w0;
                if ((unsigned)hs >= (unsigned)p.H || w0 >= p.W) continue;   // outside the volume: zero padding
                const float* src = p.x + (((long long)c.u.nb * kC1Cin + ch) * p.D + c.z) * hw + (long long)hs * p.W + w0;
                const int nv = min(8, p.W - w0);
                if (vec_ok && nv == 8) {
                    const float4 lo = __ldg(reinterpret_cast<const float4*>(src));
                    const float4 hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    f[1] = lo.x; f[2] = lo.y; f[3] = lo.z; f[4] = lo.w;
                    f[5] = hi.x; f[6] = hi.y; f[7] = hi.z; f[8] = hi.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (e < nv) f[1 + e] = __ldg(src + e);
                }
                if (w0 > 0) f[0] = __ldg(src - 1);
                if (w0 + 8 < p.W) f[9] = __ldg(src + 8);
            }
        };
        uint32_t ic = (uint32_t)g;   // sequence number of the image this group builds next
        auto store = [&](const Rows& r) {
            const uint32_t slot = ic % kC1Imgs;
            mbar_wait(iempty(slot), ((ic / kC1Imgs) & 1) ^ 1);
            const uint32_t img = smem_img + slot * kC1ImgBytes;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int t = gt + 64 * i;
                if (t >= 18 * kC1Cin) continue;
                const int ch = t / 18, hr = t - 18 * ch;
                const float (&f)[10] = r.f[i];
                uint32_t pk[3][4];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[kw][e] = pack_bf16x2(f[kw + 2 * e], f[kw + 2 * e + 1]);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int cj = hr - kh;   // output row of the brick this input row feeds through tap kh
                    if (cj < 0 || cj >= 16) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int k = ch * 9 + kh * 3 + kw;
                        st_shared_v4(img + (uint32_t)(cj >> 3) * kC1HalfBytes + (uint32_t)k * 128u +
                                         ((uint32_t)((cj & 7) ^ (k & 7)) << 4),
                                     pk[kw][0], pk[kw][1], pk[kw][2], pk[kw][3]);
                    }
                }
            }
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's (async-proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(ifull(slot));
            ic += 2;
        };
        // this group's cursor: image g, g + 2, g + 4, ... of the CTA's image sequence
        auto advance2 = [&](Cursor& c) -> bool { return advance(c) && advance(c); };
        Rows ra, rb;
        Cursor cur;
        bool have = start(cur, unit0);
        if (have && g == 1) have = advance(cur);
        if (have) load(ra, cur);
        while (have) {
            Cursor nxt = cur;
            const bool hn = advance2(nxt);
            if (hn) load(rb, nxt);
            store(ra);
            if (!hn) break;
            cur = nxt;
            have = advance2(cur);
            if (have) load(ra, cur);
            store(rb);
        }
    } else if (warp >= 4 && warp < kC1BuilderWarp0) {
        // ===================================================================== epilogue: warpgroup wg takes output
        // slices wg, wg + kC1EpiWGs, ... of the CTA's sequence
        const int wg = (warp - 4) >> 2;
        const int q = (warp - 4) & 3;           // == warp % 4: TMEM lane quarter
        const int row = q * 32 + lane;          // voxel of the brick: w = row & 7, h = row >> 3
        const int et = threadIdx.x - 128 - 128 * wg;   // 0..127 inside the warpgroup
        const int rw = row & 7, rh = row >> 3;
        const int mode = p.mode;
        const uint32_t cbuf = smem_c + wg * kBoxBytes;
        const uint32_t row_smem = cbuf + row * 128;
        const uint32_t sw = row & 7;
        const uint32_t bar_id = 1 + wg;
        if (wg == 0 && mode != EPI_PLAIN && et < 64) {
            const bool ok = et < p.ncols;
            colvec[et] = ok ? __ldg(p.vec0 + et) : 0.f;
            colvec[64 + et] = (ok && mode == EPI_AFFINE_RELU) ? __ldg(p.vec1 + et) : 0.f;
        }
        named_bar_sync(1 + kC1EpiWGs, 128 * kC1EpiWGs);
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;   // BatchNorm partial sums: columns 2*lane, 2*lane + 1, rows 32q..
        uint32_t ucnt = 0;
        // 32 accumulator columns -> bias / affine + ReLU -> bf16 -> this row's chunks 4*jj .. 4*jj+3 of the staging tile
        auto stage_half = [&](const uint32_t (&v)[32], int jj, bool row_ok) {
            const float* cv = colvec + jj * 32;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
                if (mode == EPI_AFFINE_RELU) {
                    a = fmaxf(fmaf(a, cv[2 * i], cv[64 + 2 * i]), 0.f);
                    b = fmaxf(fmaf(b, cv[2 * i + 1], cv[64 + 2 * i + 1]), 0.f);
                } else if (mode != EPI_PLAIN) {
                    a += cv[2 * i];
                    b += cv[2 * i + 1];
                }
                pk[i] = row_ok ? pack_bf16x2(a, b) : 0u;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
                st_shared_v4(row_smem + (((jj * 4 + c) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        };
        for (int unit = unit0; unit < units; unit += unit_stride) {
            const C1Unit u = decode(unit);
            const bool row_ok = (u.w0 + rw) < p.W && (u.h0 + rh) < p.H;
            for (int d = u.ds; d < u.de; ++d, ++ucnt) {
                if ((int)(ucnt % kC1EpiWGs) != wg) continue;
                const uint32_t slot = ucnt % kC1Slots, par = (ucnt / kC1Slots) & 1;
                if (et == 0) bulk_wait_read0();   // the previous TMA store has finished reading the staging tile
                named_bar_sync(bar_id, 128);
                mbar_wait(tfull(slot), par);
                tc_fence_after();
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 64;
                {   // (one half of the columns at a time: 640 threads leave 96 registers each)
                    uint32_t v[32];
                    tmem_ld32(t_addr, v);
                    tmem_ld_wait();
                    stage_half(v, 0, row_ok);
                    tmem_ld32(t_addr + 32, v);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(tempty(slot));   // accumulator drained into registers
                    stage_half(v, 1, row_ok);
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (et == 0 && B200_ABLATE(p) != 1) {
                    tma_store_5d(&p.c_map, cbuf, 0, u.w0, u.h0, d, u.nb);
                    bulk_commit();
                }
                if (mode == EPI_BIAS_STATS) {
                    // statistics of the STORED (bf16-rounded) values; rows outside the volume were staged as zeros.
                    // warp q sums rows 32q .. 32q+31 of column pair `lane` (conflict-free word reads of the tile)
                    const uint32_t base = cbuf + (q * 32) * 128 + (lane & 3) * 4;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const uint32_t wv = ld_shared_b32(base + r * 128 + ((((uint32_t)lane >> 2) ^ (r & 7)) << 4));
                        const float lo = __uint_as_float(wv << 16), hi = __uint_as_float(wv & 0xffff0000u);
                        s0 += lo; q0 = fmaf(lo, lo, q0);
                        s1 += hi; q1 = fmaf(hi, hi, q1);
                    }
                }
            }
        }
        if (et == 0) bulk_wait0();   // every output tile is in global memory before the CTA exits
        if (mode == EPI_BIAS_STATS) {
            // one partial row per CTA: stats[blockIdx.x][ncols][2], the epilogue warps' sums added in a fixed order
            *reinterpret_cast<float4*>(red + ((wg * 4 + q) * 64 + 2 * lane) * 2) = make_float4(s0, q0, s1, q1);
            named_bar_sync(1 + kC1EpiWGs, 128 * kC1EpiWGs);
            const int i = threadIdx.x - 128;   // (column, sum | sum of squares)
            if (i < 2 * p.ncols && i < 128) {
                float a = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 4 * kC1EpiWGs; ++w8) a += red[w8 * 128 + i];
                p.stats[(long long)blockIdx.x * p.ncols * 2 + i] = a;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of the same layer:  dW[co][c][kd][kh][kw] += sum over voxels of dy[voxel][co] * x[c][voxel + tap]
// (autograd of the nn.Conv3d call site models/unet3d.py:29 for inc, models/unet3d.py:194).  Same march: the slice
// image S(z) (48 rows x 128 voxels) is now the K-major B operand (K = voxels), the 8 x 16 dy brick of output slice d
// ([128 voxels][64 co], loaded by TMA) the MN-major A operand, and
//     G[:, kd*48 .. kd*48+47] += dy(d)^T x S(d + kd - 1)^T   for kd = 0..2
// is ONE MMA chain of N = 144 over the three images of the window: the image ring is laid out [half][slot][48 rows], so
// consecutive ring slots are consecutive B rows (a window that wraps the ring is issued as two runs).  The 64 x 144
// accumulator stays in TMEM over every slice the CTA visits and is added to dW once, at the end.
// Warp roles (512 threads): warp 0 TMA producer (dy bricks), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7
// epilogue, warps 8-9 / 10-11 build the even / odd slice images.
extern "C" __global__ void __launch_bounds__(kC1WgThreads, 1)
    conv1_march_wgrad_kernel(const __grid_constant__ Conv1MarchWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    constexpr uint32_t kHalfAll = kC1WgImgs * kC1HalfBytes;   // all slots' rows of one 64-voxel half
    constexpr uint32_t kPSlot = kBoxBytes;                  // dy brick [128 voxels][64 co]
    const uint32_t smem_q = smem_base;                      // [2 halves][kC1WgImgs slots][48 rows][128 B]
    const uint32_t smem_p = smem_q + 2 * kHalfAll;
    const uint32_t smem_z = smem_p + kC1WgPSlots * kPSlot;  // one zero box: the upper 64 rows of every MMA (M = 128)
    const uint32_t bar_base = smem_z + kBoxBytes;
    auto ifull = [&](uint32_t s) { return bar_base + 8 * s; };
    auto iempty = [&](uint32_t s) { return bar_base + 8 * (kC1WgImgs + s); };
    auto pfull = [&](uint32_t s) { return bar_base + 8 * (2 * kC1WgImgs + s); };
    auto pempty = [&](uint32_t s) { return bar_base + 8 * (2 * kC1WgImgs + kC1WgPSlots + s); };
    const uint32_t tfull = bar_base + 8 * (2 * kC1WgImgs + 2 * kC1WgPSlots);
    const uint32_t tmem_ptr_smem = tfull + 8;

    if (warp == 0 && lane == 0) prefetch_tmap(&p.p_map);
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < kC1WgImgs; ++s) { mbar_init(ifull(s), 2); mbar_init(iempty(s), 1); }
        for (uint32_t s = 0; s < kC1WgPSlots; ++s) { mbar_init(pfull(s), 1); mbar_init(pempty(s), 1); }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, 256);
        tmem_relinquish();
    }
    if (warp >= 8) {
        // rows 45..47 of every slice image are never written again: zero (finite operands for the unused columns)
        for (int i = threadIdx.x - 256; i < kC1WgImgs * 2 * 3 * 8; i += 64 * kC1WgGroups) {
            const int chunk = i & 7, row = 45 + (i >> 3) % 3, half = (i / 24) & 1, img = i / 48;
            st_shared_v4(smem_q + half * kHalfAll + img * kC1HalfBytes + row * 128 + (chunk << 4), 0u, 0u, 0u, 0u);
        }
        for (int i = threadIdx.x - 256; i < kBoxBytes / 16; i += 64 * kC1WgGroups) st_shared_v4(smem_z + i * 16, 0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();   // launch.cuh: the predecessor's results are complete and visible from here on
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    const int units = p.nbatch * p.nbw * p.nbh * p.nseg;
    const int unit0 = blockIdx.x, unit_stride = gridDim.x;
    auto decode = [&](int unit) {
        C1Unit u;
        int col = unit / p.nseg;
        const int seg = unit - col * p.nseg;
        const int bw = col % p.nbw; col /= p.nbw;
        const int bh = col % p.nbh; col /= p.nbh;
        u.nb = col;
        u.w0 = bw * 8;
        u.h0 = bh * 16;
        u.ds = seg * p.seg_len;
        u.de = min(p.D, u.ds + p.seg_len);
        u.z0 = max(u.ds - 1, 0);
        u.z1 = min(u.de, p.D - 1);
        return u;
    };

    if (warp == 0) {
        // ===================================================================== TMA producer: one dy brick per slice
        uint32_t pc = 0;
        for (int unit = unit0; unit < units; unit += unit_stride) {
            const C1Unit u = decode(unit);
            for (int d = u.ds; d < u.de; ++d, ++pc) {
                const uint32_t slot = pc % kC1WgPSlots;
                mbar_wait(pempty(slot), ((pc / kC1WgPSlots) & 1) ^ 1);
                if (elect_one()) {
                    const uint32_t fb = pfull(slot);
                    if (B200_ABLATE(p) == 1) {
                        mbar_arrive(fb);
                    } else {
                        mbar_arrive_expect_tx(fb, kPSlot);
                        tma_load_5d(smem_p + slot * kPSlot, &p.p_map, fb, 0, u.w0, u.h0, d, u.nb);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // A = dy brick, MN-major (M = co): 8-voxel groups SBO = 1 KB apart, one K step (16 voxels) = 2 KB; the second
        // 64-channel atom of the M = 128 rows is the shared zero box (LBO = its distance from the slot).  B = slice images, K-major (rows = image rows, 128 B = 64 voxels): one K step = 32 B
        // inside the swizzle row, voxels 64..127 one half (kHalfAll) further.
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_q, 0, 1024);
        const uint32_t idesc0 = make_idesc_bf16(128, 0, 1u, 0u);   // N field added per run
        uint32_t ibase = 0, iready = 0, pc = 0;
        uint32_t fresh = 7u;   // depth taps whose 48 accumulator columns have not been written yet
        for (int unit = unit0; unit < units; unit += unit_stride) {
            const C1Unit u = decode(unit);
            for (int d = u.ds; d < u.de; ++d, ++pc) {
                const uint32_t pslot = pc % kC1WgPSlots;
                const uint32_t need = ibase + (uint32_t)(min(d + 1, u.z1) - u.z0);
                while (iready <= need) {
                    mbar_wait(ifull(iready % kC1WgImgs), (iready / kC1WgImgs) & 1);
                    ++iready;
                }
                mbar_wait(pfull(pslot), (pc / kC1WgPSlots) & 1);
                tc_fence_after();
                // every lane walks the runs (uniform control flow, `fresh` stays warp-uniform); one elected lane issues
                const uint32_t leader = elect_one();
                const uint64_t a_desc = make_smem_desc_sw128(smem_p + pslot * kPSlot, smem_z - (smem_p + pslot * kPSlot), 1024);
                // runs of taps whose images sit in consecutive ring slots (and share their freshness)
                int kd = 0;
                while (kd < 3) {
                    const int z = d + kd - 1;
                    if (z < 0 || z >= p.D) { ++kd; continue; }   // zero padding along depth
                    const uint32_t slot = (ibase + (uint32_t)(z - u.z0)) % kC1WgImgs;
                    const bool fr = (fresh >> kd) & 1u;
                    int len = 1;
                    while (kd + len < 3 && d + kd + len - 1 < p.D && slot + len < (uint32_t)kC1WgImgs &&
                           (((fresh >> (kd + len)) & 1u) != 0) == fr)
                        ++len;
                    if (leader) {
                        const uint32_t idesc = idesc0 | ((uint32_t)(len * kC1Rows >> 3) << 17);
                        const uint32_t d_tmem = tmem_base + kd * kC1Rows;
                        const uint64_t b_desc = b_desc0 + slot * (kC1HalfBytes >> 4);
#pragma unroll
                        for (int j = 0; j < 8 && B200_ABLATE(p) != 3; ++j)
                            umma_f16(d_tmem, a_desc + 128 * j, b_desc + (j >> 2) * (kHalfAll >> 4) + (j & 3) * 2, idesc,
                                     (fr && j == 0) ? 0u : 1u);
                    }
                    fresh &= ~(((1u << len) - 1u) << kd);
                    kd += len;
                }
                if (leader) {
                    umma_commit(pempty(pslot));
                    if (d - 1 >= u.z0) umma_commit(iempty((ibase + (uint32_t)(d - 1 - u.z0)) % kC1WgImgs));
                    if (d == u.de - 1)
                        for (int z = max(u.z0, d); z <= u.z1; ++z)
                            umma_commit(iempty((ibase + (uint32_t)(z - u.z0)) % kC1WgImgs));
                }
                __syncwarp();
            }
            ibase += (uint32_t)(u.z1 - u.z0 + 1);
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else if (warp >= 8) {
        // ===================================================================== slice-image builders (as in the forward)
        const int g = (warp - 8) >> 1;
        const int gt = threadIdx.x - 256 - 64 * g;
        const long long hw = (long long)p.H * p.W;
        const bool vec_ok = (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
        struct Rows {
            float f[2][10];
        };
        struct Cursor {
            int unit, z;
            C1Unit u;
        };
        auto start = [&](Cursor& c, int unit) -> bool {
            if (unit >= units) return false;
            c.unit = unit;
            c.u = decode(unit);
            c.z = c.u.z0;
            return true;
        };
        auto advance = [&](Cursor& c) -> bool {
            if (c.z < c.u.z1) { ++c.z; return true; }
            return start(c, c.unit + unit_stride);
        };
        auto load = [&](Rows& r, const Cursor& c) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float (&f)[10] = r.f[i];
#pragma unroll
                for (int e = 0; e < 10; ++e) f[e] = 0.f;
                const int t = gt + 64 * i;
                if (t >= 18 * kC1Cin || B200_ABLATE(p) == 2) continue;
                const int ch = t / 18, hr = t - 18 * ch;
                const int hs = c.u.h0 - 1 + hr, w0 = c.u.w0;
                if ((unsigned)hs >= (unsigned)p.H || w0 >= p.W) continue;
                const float* src = p.x + (((long long)c.u.nb * kC1Cin + ch) * p.D + c.z) * hw + (long long)hs * p.W + w0;
                const int nv = min(8, p.W - w0);
                if (vec_ok && nv == 8) {
                    const float4 lo = __ldg(reinterpret_cast<const float4*>(src));
                    const float4 hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
                    f[1] = lo.x; f[2] = lo.y; f[3] = lo.z; f[4] = lo.w;
                    f[5] = hi.x; f[6] = hi.y; f[7] = hi.z; f[8] = hi.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (e < nv) f[1 + e] = __ldg(src + e);
                }
                if (w0 > 0) f[0] = __ldg(src - 1);
                if (w0 + 8 < p.W) f[9] = __ldg(src + 8);
            }
        };
        uint32_t ic = (uint32_t)g;
        auto store = [&](const Rows& r) {
            const uint32_t slot = ic % kC1WgImgs;
            mbar_wait(iempty(slot), ((ic / kC1WgImgs) & 1) ^ 1);
            const uint32_t img = smem_q + slot * kC1HalfBytes;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int t = gt + 64 * i;
                if (t >= 18 * kC1Cin) continue;
                const int ch = t / 18, hr = t - 18 * ch;
                const float (&f)[10] = r.f[i];
                uint32_t pk[3][4];
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[kw][e] = pack_bf16x2(f[kw + 2 * e], f[kw + 2 * e + 1]);
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    const int cj = hr - kh;
                    if (cj < 0 || cj >= 16) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int k = ch * 9 + kh * 3 + kw;
                        st_shared_v4(img + (uint32_t)(cj >> 3) * kHalfAll + (uint32_t)k * 128u +
                                         ((uint32_t)((cj & 7) ^ (k & 7)) << 4),
                                     pk[kw][0], pk[kw][1], pk[kw][2], pk[kw][3]);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(ifull(slot));
            ic += kC1WgGroups;
        };
        auto advance_n = [&](Cursor& c) -> bool {   // this group's next image: kC1WgGroups further in the sequence
            for (int i = 0; i < kC1WgGroups; ++i)
                if (!advance(c)) return false;
            return true;
        };
        Rows ra, rb;
        Cursor cur;
        bool have = start(cur, unit0);
        for (int i = 0; i < g && have; ++i) have = advance(cur);
        if (have) load(ra, cur);
        while (have) {
            Cursor nxt = cur;
            const bool hn = advance_n(nxt);
            if (hn) load(rb, nxt);
            store(ra);
            if (!hn) break;
            cur = nxt;
            have = advance_n(cur);
            if (have) load(ra, cur);
            store(rb);
        }
    } else if (warp >= 4 && warp < 6) {
        // ===================================================================== epilogue: dW += accumulator (rows = co;
        // TMEM lanes 64..127 hold the zero-filled upper channel atom)
        const int co = (warp - 4) * 32 + lane;
        // depth taps this CTA's slices contributed to (a tap whose slices all fell outside the volume was never
        // accumulated: its TMEM columns are not initialised)
        uint32_t written = 0;
        for (int unit = unit0; unit < units; unit += unit_stride) {
            const C1Unit u = decode(unit);
            if (u.ds >= u.de) continue;
            written |= 2u;
            if (u.de >= 2) written |= 1u;          // some slice d >= 1 reads input slice d - 1
            if (u.ds <= p.D - 2) written |= 4u;    // some slice d <= D - 2 reads input slice d + 1
        }
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>((warp - 4) * 32) << 16);
        for (int kd = 0; kd < 3; ++kd) {
            if (!((written >> kd) & 1u)) continue;
#pragma unroll
            for (int c16 = 0; c16 < 3; ++c16) {
                uint32_t v[16];
                tmem_ld16(t_addr + kd * kC1Rows + c16 * 16, v);
                tmem_ld_wait();
                if (co < p.ncols) {
                    float* dst = p.dw + (long long)co * (27 * kC1Cin) + kd * 9;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = c16 * 16 + j;   // c*9 + kh*3 + kw
                        if (k < 9 * kC1Cin) atomicAdd(dst + (k / 9) * 27 + (k % 9), __uint_as_float(v[j]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace b200
