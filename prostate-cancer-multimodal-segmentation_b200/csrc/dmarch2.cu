// Depth-marching implicit GEMM on CTA pairs (tcgen05.mma.cta_group::2, M = 256): the 64-column 3x3x3 convolutions at
// full resolution (fprop with Cout = 64 — models/unet3d.py:29,35 at inc / up4 — and, given transposed weights, dgrad
// with Cin = 64).  Same march as dmarch.cu: for input slice z one MMA A(z, kh, kw) x [W(kd=2) | W(kd=1) | W(kd=0)]
// (N = 192) adds into the accumulators of output slices z-1, z, z+1.
//
// Why another kernel.  dmarch_kernel is bound by shared-memory bandwidth: every 128 x 192 x 16 MMA reads 4 KB of A and
// 6 KB of B (107 B/clk) while TMA fills 47 B/clk more — 151 B/clk against the 128 B/clk of an SM, which is the 68-75 %
// tensor-pipe ceiling ncu shows for it.  With cta_group::2 two CTAs march two adjacent brick columns in lockstep, each
// supplies its own A rows and HALF of B (N columns 0..95 / 96..191 = rows of the K-major weight tile), so an SM reads
// 4 + 3 KB per MMA and fills 31 B/clk: 104 B/clk.  dmarch_pair_kernel already shared the weight stream by multicast
// (L2 reads / 2), but every SM still received and read all of B.
//
// What the N split forces (an MMA must always cover all three slabs, its D columns are contiguous):
//   * uniform windows: a unit processes input slices ds-1 .. de with full N = 192 MMAs; the four slices ds-2, ds-1,
//     de, de+1 that fall outside the unit are dummies whose accumulators are thrown away (input slices outside the
//     volume are zero-filled by TMA);
//   * no fresh-slab MMAs: every MMA accumulates, the epilogue zeroes a slot (tcgen05.st) after draining it;
//   * a ring that never wraps inside a window: output slice s lives in logical slot s mod 6; the window starting at
//     slice f covers physical slots (f mod 6) + {0, 1, 2} <= 7 of the eight 64-column TMEM slots, so slices with
//     s mod 6 = 0 / 1 collect part of their sum in the mirror slots 6 / 7 and the epilogue adds the two parts.
//
// Warp roles (256 threads, clusters of 2): warp 0 TMA producer (own A boxes, own half of B; bytes counted on the
// leader's barriers), warp 1 of the leader issues every MMA and commits to both CTAs' barriers, warp 2 TMEM allocator,
// warps 4..7 epilogue (own TMEM rows -> own output column; release on the leader's barrier).
#include <cuda_bf16.h>
#include "igemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {

namespace {
struct Ring2 {
    uint32_t stage = 0, phase = 0;
    DEV void advance(uint32_t n) {
        if (++stage == n) { stage = 0; phase ^= 1; }
    }
};
}  // namespace

extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    dmarch2_kernel(const __grid_constant__ DmarchParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();

    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + kDmAStages * kDmAStageBytes;
    const uint32_t smem_c = smem_b + kDm2BStages * kDm2BBytes;   // 16 KB output staging tile
    const uint32_t bar_base = smem_c + kBoxBytes;
    auto afull = [&](uint32_t s) { return bar_base + 8 * s; };
    auto aempty = [&](uint32_t s) { return bar_base + 8 * (kDmAStages + s); };
    auto bfull = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + s); };
    auto bempty = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + kDm2BStages + s); };
    auto tfull = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + 2 * kDm2BStages + s); };
    auto tempty = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + 2 * kDm2BStages + kDm2Slots + s); };
    const uint32_t tmem_ptr_smem = bar_base + 8 * (2 * kDmAStages + 2 * kDm2BStages + 2 * kDm2Slots);
    const uint32_t f_off = (tmem_ptr_smem + 16 - smem_base + 15u) & ~15u;
    float* scratch = reinterpret_cast<float*>(smem_gen + f_off);  // [4 warps][64][2]
    float* colacc = scratch + 4 * 64 * 2;                         // [64][2]
    float* colvec = colacc + 128;                                 // [2][64]

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.a_map);
        prefetch_tmap(&p.b_map);
        prefetch_tmap(&p.c_map);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < kDmAStages; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
        for (uint32_t s = 0; s < kDm2BStages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        // an accumulator is handed back by the epilogue threads of BOTH CTAs, on the leader's barrier
        for (uint32_t s = 0; s < kDm2Slots; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 256); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_ptr_smem, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    pdl_wait();   // launch.cuh: the predecessor's results are complete and visible from here on
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    const int columns = p.nbatch * p.nbw * p.nbh;
    // Work = (pair of adjacent columns, output slice) steps in one linear order (column pair major).  Cluster c of n
    // takes the contiguous range [total * c / n, total * (c + 1) / n) and walks it in pieces that end at column ends:
    // every cluster gets the same number of slices (no wave of left-over units), and the two boundary input slices are
    // paid once per piece (~2 pieces per cluster) instead of once per fixed depth segment.  CTA `rank` owns column
    // 2 * cpair + rank (a column past the last one has every coordinate out of range: zero-filled loads, clipped stores)
    const long long total = (long long)((columns + 1) / 2) * p.D;
    const long long ncl = gridDim.x / 2, cl = blockIdx.x / 2;
    const long long l_begin = total * cl / ncl, l_end = total * (cl + 1) / ncl;
    const int kc_blocks = p.kc_blocks;
    const int sign = p.sign;

    // next piece of this cluster's range: (batch, brick column origin, output slices [ds, de)); false when done
    auto next_piece = [&](long long& l, int& nb, int& w0, int& h0, int& ds, int& de) -> bool {
        if (l >= l_end) return false;
        const int cp = (int)(l / p.D);
        ds = (int)(l - (long long)cp * p.D);
        de = (int)min((long long)p.D, (long long)ds + (l_end - l));
        l += de - ds;
        int c = cp * 2 + rank;
        const int bw = c % p.nbw; c /= p.nbw;
        const int bh = c % p.nbh; c /= p.nbh;
        nb = c;
        w0 = bw * 8;
        h0 = bh * 16;
        return true;
    };

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        Ring2 ra, rb;
        int nb, w0, h0, ds, de;
        for (long long l = l_begin; next_piece(l, nb, w0, h0, ds, de);) {
            for (int dz = ds - 1; dz <= de; dz += kDmG) {
                const int nin = min(kDmG, de - dz + 1);   // input slices that share this pass over the weights
                for (int kw = 0; kw < 3; ++kw) {
                    for (int kc = 0; kc < kc_blocks; ++kc) {
                        mbar_wait(aempty(ra.stage), ra.phase ^ 1);
                        if (elect_one()) {
                            const uint32_t fb = afull(ra.stage);
                            if (rank == 0) mbar_arrive_expect_tx(fb, 2 * nin * kDmABytes);   // both CTAs' boxes
                            for (int si = 0; si < nin; ++si)
                                tma_load_5d_pair(smem_a + ra.stage * kDmAStageBytes + si * kDmABytes, &p.a_map, fb,
                                                 kc * 64, w0 + sign * (kw - 1), h0 - 1, dz + si, nb);
                        }
                        __syncwarp();
                        ra.advance(kDmAStages);
                        for (int kh = 0; kh < 3; ++kh) {
                            mbar_wait(bempty(rb.stage), rb.phase ^ 1);
                            if (elect_one()) {
                                const uint32_t fb = bfull(rb.stage);
                                if (rank == 0) mbar_arrive_expect_tx(fb, 2 * kDm2BBytes);
                                // this CTA's 96 of the 192 B rows: 32-row boxes rank*3 .. rank*3+2 of [slab 0 | 1 | 2]
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const int r32 = rank * 3 + i;
                                    const int j = r32 >> 1;                 // slab j feeds output slice dz - 1 + j
                                    const int kd = sign > 0 ? 2 - j : j;
                                    const int tap = kd * 9 + kw * 3 + kh;   // packed tap order
                                    tma_load_3d_pair(smem_b + rb.stage * kDm2BBytes + i * 4096, &p.b_map, fb, kc * 64,
                                                     (r32 & 1) * 32, tap);
                                }
                            }
                            __syncwarp();
                            rb.advance(kDm2BStages);
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================================================================== MMA issuer (leader CTA only)
        Ring2 ra, rb;
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_a, 0, 1024);
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_b, 0, 1024);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
        const uint32_t idesc = make_idesc_bf16(256, 192, 0u, 0u);
        const int nk_last = ((p.cin - (kc_blocks - 1) * 64) + 15) >> 4;
        uint32_t sbase = 0;      // slices (dummies included) of the units before this one
        uint32_t acquired = 0;   // slices whose slot is known to be zeroed and free
        int nb, w0, h0, ds, de;
        for (long long l = l_begin; next_piece(l, nb, w0, h0, ds, de);) {
            for (int dz0 = ds - 1; dz0 <= de; dz0 += kDmG) {
                const int nin = min(kDmG, de - dz0 + 1);
                // the window of input slice dz starts at slice index f = sbase + (dz - ds + 1) (slice ds-2 is sbase)
                const uint32_t f0 = sbase + (uint32_t)(dz0 - ds + 1);
                while (acquired <= f0 + (uint32_t)nin + 1u) {
                    mbar_wait(tempty(acquired % kDm2Slots), (acquired / kDm2Slots) & 1);
                    ++acquired;
                }
                uint32_t d_tm[kDmG];
#pragma unroll
                for (int si = 0; si < kDmG; ++si) d_tm[si] = tmem_base + ((f0 + (uint32_t)si) % kDm2Slots) * 64;
                tc_fence_after();
                for (int kw = 0; kw < 3; ++kw) {
                    for (int kc = 0; kc < kc_blocks; ++kc) {
                        const int nk = (kc == kc_blocks - 1) ? nk_last : 4;
                        mbar_wait(afull(ra.stage), ra.phase);
                        const uint32_t a_st = a_lo0 + ra.stage * (kDmAStageBytes >> 4);
                        for (int kh = 0; kh < 3; ++kh) {
                            mbar_wait(bfull(rb.stage), rb.phase);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t b_lo = b_lo0 + rb.stage * (kDm2BBytes >> 4);
#pragma unroll
                                for (int si = 0; si < kDmG; ++si) {
                                    if (si < nin) {
                                        // tap kh reads the halo box at row offset kh (fprop) or 2 - kh (dgrad): 8 rows = 1 KB
                                        const uint32_t a_lo = a_st + si * (kDmABytes >> 4) +
                                                              (uint32_t)((sign > 0 ? kh : 2 - kh) * (1024 >> 4));
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            if (k < nk)
                                                umma_f16_lohi_pair(d_tm[si], a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc,
                                                                   1u);
                                    }
                                }
                                umma_commit_pair(bempty(rb.stage), 3u);   // frees the slot in both CTAs
                            }
                            __syncwarp();
                            rb.advance(kDm2BStages);
                        }
                        if (elect_one()) umma_commit_pair(aempty(ra.stage), 3u);
                        __syncwarp();
                        ra.advance(kDmAStages);
                    }
                }
                // the first slice of every window has now received its last contribution; so have the two trailing
                // dummies after the unit's last input slice
                if (elect_one()) {
                    for (int si = 0; si < nin; ++si) umma_commit_pair(tfull((f0 + (uint32_t)si) % kDm2Slots), 3u);
                    if (dz0 + nin > de) {
                        umma_commit_pair(tfull((f0 + (uint32_t)nin) % kDm2Slots), 3u);
                        umma_commit_pair(tfull((f0 + (uint32_t)nin + 1u) % kDm2Slots), 3u);
                    }
                }
                __syncwarp();
            }
            sbase += (uint32_t)(de - ds + 4);
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (one slice at a time)
        const int q = warp - 4;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;
        const int rw = row & 7, rh = row >> 3;
        const int mode = p.mode;
        const uint32_t row_smem = smem_c + row * 128;
        const uint32_t sw = row & 7;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (mode == EPI_BIAS_STATS) {
            for (int i = et; i < 128; i += 128) colacc[i] = 0.f;
        }
        if (mode != EPI_PLAIN) {
            if (et < 64) {   // columns past ncols (32-column layers run as 64 with zero weights) get neutral values
                const bool ok = et < p.ncols;
                colvec[et] = ok ? __ldg(p.vec0 + et) : 0.f;
                if (mode == EPI_AFFINE_RELU) colvec[64 + et] = ok ? __ldg(p.vec1 + et) : 0.f;
            }
        }
        // every accumulator starts from zero: all MMAs accumulate
#pragma unroll 4
        for (int c = 0; c < 512; c += 16) tmem_st16_zero(t_lane + c);
        tmem_st_wait();
        tc_fence_before();
        for (uint32_t s = 0; s < kDm2Slots; ++s) mbar_arrive_leader(tempty(s));
        named_bar_sync(1, 128);
        uint32_t sidx = 0;   // slice index in the CTA's sequence
        int nb, w0, h0, ds, de;
        for (long long l = l_begin; next_piece(l, nb, w0, h0, ds, de);) {
            const bool row_ok = (w0 + rw) < p.W && (h0 + rh) < p.H && nb < p.nbatch;
            for (int d = ds - 2; d <= de + 1; ++d, ++sidx) {
                const uint32_t m = sidx % kDm2Slots, par = (sidx / kDm2Slots) & 1;
                const bool real = d >= ds && d < de;
                if (real) {
                    if (et == 0) bulk_wait_read0();  // previous TMA store finished reading the staging tile
                    named_bar_sync(1, 128);
                }
                mbar_wait(tfull(m), par);
                tc_fence_after();
                const uint32_t t_main = t_lane + m * 64, t_mirror = t_lane + (kDm2Slots + m) * 64;
                uint32_t v0[32], v1[32];
                if (real) {
                    tmem_ld32(t_main, v0);
                    tmem_ld32(t_main + 32, v1);
                    tmem_ld_wait();
                    if (m < 2) {   // the part of the sum that was accumulated in the mirror slot
                        uint32_t u0[32];
                        tmem_ld32(t_mirror, u0);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v0[i] = __float_as_uint(__uint_as_float(v0[i]) + __uint_as_float(u0[i]));
                        tmem_ld32(t_mirror + 32, u0);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v1[i] = __float_as_uint(__uint_as_float(v1[i]) + __uint_as_float(u0[i]));
                    }
                }
                // zero the slot(s) for their next slice and hand them back
#pragma unroll
                for (int c = 0; c < 64; c += 16) tmem_st16_zero(t_main + c);
                if (m < 2) {
#pragma unroll
                    for (int c = 0; c < 64; c += 16) tmem_st16_zero(t_mirror + c);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive_leader(tempty(m));
                if (!real) continue;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const uint32_t (&v)[32] = jj == 0 ? v0 : v1;
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
                        st_shared_v4(row_smem + (((jj * 4 + c) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2],
                                     pk[4 * c + 3]);
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (et == 0) {
                    tma_store_5d(&p.c_map, smem_c, 0, w0, h0, d, nb);
                    bulk_commit();
                }
                if (mode == EPI_BIAS_STATS) {
                    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                    const uint32_t base = smem_c + (q * 32) * 128 + (lane & 3) * 4;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) {
                        const uint32_t wv = ld_shared_b32(base + r * 128 + ((((uint32_t)lane >> 2) ^ (r & 7)) << 4));
                        const float lo = __uint_as_float(wv << 16), hi = __uint_as_float(wv & 0xffff0000u);
                        s0 += lo; q0 = fmaf(lo, lo, q0);
                        s1 += hi; q1 = fmaf(hi, hi, q1);
                    }
                    *reinterpret_cast<float4*>(scratch + (q * 64 + 2 * lane) * 2) = make_float4(s0, q0, s1, q1);
                    named_bar_sync(1, 128);
                    if (et < 64) {
                        float a = 0.f, b2 = 0.f;
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            a += scratch[(w4 * 64 + et) * 2 + 0];
                            b2 += scratch[(w4 * 64 + et) * 2 + 1];
                        }
                        colacc[2 * et] += a;
                        colacc[2 * et + 1] += b2;
                    }
                }
            }
        }
        if (et == 0) bulk_wait0();
        if (mode == EPI_BIAS_STATS) {
            named_bar_sync(1, 128);
            float* dst = p.stats + (long long)blockIdx.x * 2 * p.ncols;   // [gridDim.x][ncols][2]
            for (int i = et; i < 2 * p.ncols; i += 128) dst[i] = colacc[i];
        }
    }

    tc_fence_before();
    cluster_sync_all();   // a peer may still signal this CTA's barriers until here
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace b200
