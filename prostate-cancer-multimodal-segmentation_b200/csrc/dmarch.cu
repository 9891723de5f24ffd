// Depth-marching tcgen05 implicit GEMM (see DmarchParams in igemm.cuh): conv3d fprop with Cout = 64 and dgrad with
// Cin = 64 — the full-resolution layers that hold 52 % of the network's FLOPs (models/unet3d.py:29,35 at inc / up4).
//
// Warp roles (256 threads, 1 CTA / SM): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..7
// epilogue.  Three pipelines: A ring (h-halo boxes, one per (kw, channel block)), B ring (weights of one (kh, kw,
// channel block): three kd slabs), TMEM ring (one 64-column slot per output slice).
#include <cuda_bf16.h>
#include "igemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {

struct Ring {
    uint32_t stage = 0, phase = 0;
    DEV void advance(uint32_t n) {
        if (++stage == n) { stage = 0; phase ^= 1; }
    }
};

// kPair: launched as clusters of two CTAs that march two adjacent brick columns through the same depth segment in
// lockstep and SHARE the weight stream: every B stage is fetched once, by the CTA whose rank equals the parity of the
// stage count, and multicast into both CTAs' shared memory.  Every large GEMM here runs against the L2 -> SM fabric
// (~10-11 TB/s measured in all ncu captures); the weights are 2/3 of this kernel's fabric bytes.
template <bool kPair>
DEV void dmarch_body(const DmarchParams& p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + kDmAStages * kDmAStageBytes;
    const uint32_t smem_c = smem_b + kDmBStages * kDmBBytes;   // 16 KB output staging tile
    const uint32_t bar_base = smem_c + kBoxBytes;
    auto afull = [&](uint32_t s) { return bar_base + 8 * s; };
    auto aempty = [&](uint32_t s) { return bar_base + 8 * (kDmAStages + s); };
    auto bfull = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + s); };
    auto bempty = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + kDmBStages + s); };
    auto tfull = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + 2 * kDmBStages + s); };
    auto tempty = [&](uint32_t s) { return bar_base + 8 * (2 * kDmAStages + 2 * kDmBStages + kDmSlots + s); };
    const uint32_t tmem_ptr_smem = bar_base + 8 * (2 * kDmAStages + 2 * kDmBStages + 2 * kDmSlots);
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
        // a B stage is rewritten in both CTAs of a pair at once: both MMA issuers release it
        for (uint32_t s = 0; s < kDmBStages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), kPair ? 2 : 1); }
        for (uint32_t s = 0; s < kDmSlots; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 128); }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    pdl_wait();   // launch.cuh: the predecessor's results are complete and visible from here on
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    const int columns = p.nbatch * p.nbw * p.nbh;
    const int cmul = kPair ? 2 : 1;
    const int rank = kPair ? (int)cluster_ctarank() : 0;
    // pair mode: a unit is (pair of adjacent columns, depth segment); CTA `rank` owns column 2 * cpair + rank (a
    // column past the last one has every coordinate out of range: zero-filled loads, clipped stores)
    const int units = ((columns + cmul - 1) / cmul) * p.nseg;
    const int unit0 = blockIdx.x / cmul, unit_stride = gridDim.x / cmul;
    const int kc_blocks = p.kc_blocks;
    const int sign = p.sign;

    // unit -> (batch, brick column, depth segment); segments of one column are consecutive units
    auto decode = [&](int unit, int& nb, int& w0, int& h0, int& ds, int& de) {
        const int cu = unit / p.nseg, seg = unit - cu * p.nseg;
        const int col = cu * cmul + rank;
        int c = col;
        const int bw = c % p.nbw; c /= p.nbw;
        const int bh = c % p.nbh; c /= p.nbh;
        nb = c;
        w0 = bw * 8;
        h0 = bh * 16;
        ds = seg * p.seg_len;
        de = min(p.D, ds + p.seg_len);
    };

    if (warp == 0) {
        // ===================================================================== TMA producer
        Ring ra, rb;
        uint32_t bcount = 0;   // B stages so far: in pair mode the CTA of that parity fetches the stage for both
        for (int unit = unit0; unit < units; unit += unit_stride) {
            int nb, w0, h0, ds, de;
            decode(unit, nb, w0, h0, ds, de);
            if (ds >= de) continue;
            const int z0 = max(ds - 1, 0), z1 = min(de, p.D - 1);
            for (int dz = z0; dz <= z1; dz += kDmG) {
                const int nin = min(kDmG, z1 - dz + 1);   // input slices that share this pass over the weights
                for (int kw = 0; kw < 3; ++kw) {
                    for (int kc = 0; kc < kc_blocks; ++kc) {
                        mbar_wait(aempty(ra.stage), ra.phase ^ 1);
                        if (elect_one()) {
                            const uint32_t fb = afull(ra.stage);
                            if (B200_ABLATE(p) == 1) { mbar_arrive(fb); }
                            else mbar_arrive_expect_tx(fb, nin * kDmABytes);
                            for (int si = 0; si < nin && B200_ABLATE(p) != 1; ++si)
                                tma_load_5d(smem_a + ra.stage * kDmAStageBytes + si * kDmABytes, &p.a_map, fb, kc * 64,
                                            w0 + sign * (kw - 1), h0 - 1, dz + si, nb);
                        }
                        __syncwarp();
                        ra.advance(kDmAStages);
                        for (int kh = 0; kh < 3; ++kh) {
                            mbar_wait(bempty(rb.stage), rb.phase ^ 1);
                            if (elect_one()) {
                                const uint32_t fb = bfull(rb.stage);
                                if (B200_ABLATE(p) == 1) { mbar_arrive(fb); }
                                else mbar_arrive_expect_tx(fb, kDmBBytes);
                                if (B200_ABLATE(p) != 1 && (!kPair || (int)(bcount & 1u) == rank)) {
#pragma unroll
                                    for (int j = 0; j < 3; ++j) {
                                        // slab j feeds output slice dz - 1 + j
                                        const int kd = sign > 0 ? 2 - j : j;
                                        const int tap = kd * 9 + kw * 3 + kh;  // packed tap order
                                        const uint32_t dst = smem_b + rb.stage * kDmBBytes + j * 8192;
                                        const int c0 = p.b_mn ? 0 : kc * 64, c1 = p.b_mn ? kc * 64 : 0;
                                        if (kPair) tma_load_3d_mc(dst, &p.b_map, fb, c0, c1, tap, 3u);
                                        else tma_load_3d(dst, &p.b_map, fb, c0, c1, tap);
                                    }
                                }
                            }
                            __syncwarp();
                            ++bcount;
                            rb.advance(kDmBStages);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // The issuing lane must stay lean (a dependent chain of uniform-datapath instructions per MMA is what bounds a
        // short-N pipeline): descriptors are 32-bit low words + a constant high word, the per-slice plan (which TMEM
        // columns, which B slabs, which instruction descriptor) is computed once per input slice.
        Ring ra, rb;
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_a, 0, 1024);
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_b, p.b_mn ? 8192 : 0, 1024);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
        const uint32_t kinc_b = p.b_mn ? 128u : 2u;
        const uint32_t bmaj = p.b_mn ? 1u : 0u;
        const uint32_t idesc0 = make_idesc_bf16(128, 0, 0, bmaj);   // N field added per run: (64 * slabs) >> 3 at bit 17
        const uint32_t idesc64 = idesc0 | (8u << 17);
        const int nk_last = ((p.cin - (kc_blocks - 1) * 64) + 15) >> 4;
        uint32_t ubase = 0;  // TMEM-slot use index of output slice ds of the current unit
        for (int unit = unit0; unit < units; unit += unit_stride) {
            int nb, w0, h0, ds, de;
            decode(unit, nb, w0, h0, ds, de);
            if (ds >= de) continue;
            const int z0 = max(ds - 1, 0), z1 = min(de, p.D - 1);
            for (int dz0 = z0; dz0 <= z1; dz0 += kDmG) {
                // kDmG consecutive input slices share one pass over the 27 tap weights (the B stream, 216 KB per
                // pass and channel block, is the larger half of the shared-memory fill traffic of this kernel)
                const int nin = min(kDmG, z1 - dz0 + 1);
                uint32_t tm[kDmG][3], acc0[kDmG][3];
                int jlo[kDmG], jhi[kDmG];
                uint32_t r_tm0[kDmG], r_bo0[kDmG], r_id0[kDmG], r_tm1[kDmG], r_bo1[kDmG], r_id1[kDmG];
                bool two[kDmG];
#pragma unroll
                for (int si = 0; si < kDmG; ++si) {
                    const int dz = dz0 + si;
                    // slabs j (output slice d = dz - 1 + j) that belong to this unit
                    jlo[si] = max(0, ds - (dz - 1));
                    jhi[si] = min(2, (de - 1) - (dz - 1));
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int d = dz - 1 + j;
                        const uint32_t u = ubase + (uint32_t)(d - ds);
                        const uint32_t sl = u % kDmSlots;
                        tm[si][j] = tmem_base + sl * 64;
                        const bool fresh = si < nin && (j >= jlo[si] && j <= jhi[si]) && (dz == max(d - 1, 0));
                        acc0[si][j] = fresh ? 0u : 1u;
                        if (fresh) mbar_wait(tempty(sl), ((u / kDmSlots) & 1) ^ 1);  // previous use drained
                    }
                    // regular k-steps: one MMA over all slabs, or two when the ring wraps inside the window
                    const int lo = jlo[si], hi = jhi[si];
                    const uint32_t tm_lo = lo == 0 ? tm[si][0] : (lo == 1 ? tm[si][1] : tm[si][2]);
                    r_tm0[si] = tm_lo; r_bo0[si] = lo * (8192u >> 4); r_id0[si] = idesc0 | ((uint32_t)(hi - lo + 1) << 20);
                    r_tm1[si] = 0; r_bo1[si] = 0; r_id1[si] = 0;
                    two[si] = false;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j >= lo && j < hi && tm[si][j + 1] != tm[si][j] + 64) {
                            two[si] = true;
                            r_id0[si] = idesc0 | ((uint32_t)(j + 1 - lo) << 20);
                            r_tm1[si] = tm[si][j + 1];
                            r_bo1[si] = (j + 1) * (8192u >> 4);
                            r_id1[si] = idesc0 | ((uint32_t)(hi - j) << 20);
                        }
                    }
                }
                tc_fence_after();
                bool first = true;
                for (int kw = 0; kw < 3; ++kw) {
                    for (int kc = 0; kc < kc_blocks; ++kc) {
                        const int nk = (kc == kc_blocks - 1) ? nk_last : 4;
                        mbar_wait(afull(ra.stage), ra.phase);
                        const uint32_t a_st = a_lo0 + ra.stage * (kDmAStageBytes >> 4);
                        for (int kh = 0; kh < 3; ++kh) {
                            mbar_wait(bfull(rb.stage), rb.phase);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint32_t b_lo = b_lo0 + rb.stage * (kDmBBytes >> 4);
#pragma unroll
                                for (int si = 0; si < kDmG; ++si) {
                                    if (si < nin && jlo[si] <= jhi[si] && B200_ABLATE(p) != 2) {
                                        // tap kh reads the halo box at row offset kh (fprop) or 2 - kh (dgrad): 8 rows = 1 KB
                                        const uint32_t a_lo = a_st + si * (kDmABytes >> 4) +
                                                              (uint32_t)((sign > 0 ? kh : 2 - kh) * (1024 >> 4));
                                        int k0 = 0;
                                        if (first) {
                                            // first k-step of the slice: per-slab MMAs (a fresh slab must not accumulate)
#pragma unroll
                                            for (int j = 0; j < 3; ++j)
                                                if (j >= jlo[si] && j <= jhi[si])
                                                    umma_f16_lohi(tm[si][j], a_lo, a_hi, b_lo + j * (8192u >> 4), b_hi,
                                                                  idesc64, acc0[si][j]);
                                            k0 = 1;
                                        }
                                        if (!two[si]) {
#pragma unroll
                                            for (int k = 0; k < 4; ++k)
                                                if (k >= k0 && k < nk)
                                                    umma_f16_lohi(r_tm0[si], a_lo + 2 * k, a_hi,
                                                                  b_lo + r_bo0[si] + k * kinc_b, b_hi, r_id0[si], 1u);
                                        } else {
#pragma unroll
                                            for (int k = 0; k < 4; ++k)
                                                if (k >= k0 && k < nk) {
                                                    umma_f16_lohi(r_tm0[si], a_lo + 2 * k, a_hi,
                                                                  b_lo + r_bo0[si] + k * kinc_b, b_hi, r_id0[si], 1u);
                                                    umma_f16_lohi(r_tm1[si], a_lo + 2 * k, a_hi,
                                                                  b_lo + r_bo1[si] + k * kinc_b, b_hi, r_id1[si], 1u);
                                                }
                                        }
                                    }
                                }
                                if (kPair) umma_commit_mc(bempty(rb.stage), 3u);   // both CTAs' producers
                                else umma_commit(bempty(rb.stage));
                            }
                            __syncwarp();
                            first = false;
                            rb.advance(kDmBStages);
                        }
                        if (elect_one()) umma_commit(aempty(ra.stage));
                        __syncwarp();
                        ra.advance(kDmAStages);
                    }
                }
                // output slices whose last contribution came from one of these input slices are complete
                if (elect_one()) {
                    const int dlo = max(ds, dz0 - 1), dhi = min(de - 1, dz0 + nin);
                    for (int d = dlo; d <= dhi; ++d) {
                        const int last = min(d + 1, p.D - 1);
                        if (last >= dz0 && last < dz0 + nin) {
                            const uint32_t u = ubase + (uint32_t)(d - ds);
                            umma_commit(tfull(u % kDmSlots));
                        }
                    }
                }
                __syncwarp();
            }
            ubase += (uint32_t)(de - ds);
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (one output slice at a time)
        const int q = warp - 4;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;
        const int rw = row & 7, rh = row >> 3;
        const int mode = p.mode;
        const uint32_t row_smem = smem_c + row * 128;
        const uint32_t sw = row & 7;
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
        named_bar_sync(1, 128);
        uint32_t u = 0;
        for (int unit = unit0; unit < units; unit += unit_stride) {
            int nb, w0, h0, ds, de;
            decode(unit, nb, w0, h0, ds, de);
            const bool row_ok = (w0 + rw) < p.W && (h0 + rh) < p.H && nb < p.nbatch;
            for (int d = ds; d < de; ++d, ++u) {
                const uint32_t slot = u % kDmSlots, par = (u / kDmSlots) & 1;
                if (et == 0) bulk_wait_read0();  // previous TMA store finished reading the staging tile
                named_bar_sync(1, 128);
                mbar_wait(tfull(slot), par);
                tc_fence_after();
                if (B200_ABLATE(p) == 4) {   // dev: the epilogue only hands the accumulator back
                    tc_fence_before();
                    mbar_arrive(tempty(slot));
                    continue;
                }
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + slot * 64;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + jj * 32, v);
                    tmem_ld_wait();
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
                tc_fence_before();
                mbar_arrive(tempty(slot));
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
    if (kPair) cluster_sync_all(); else __syncthreads();   // a peer may still signal this CTA's barriers until here
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

extern "C" __global__ void __launch_bounds__(kThreads, 1) dmarch_kernel(const __grid_constant__ DmarchParams p) {
    dmarch_body<false>(p);
}
extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    dmarch_pair_kernel(const __grid_constant__ DmarchParams p) {
    dmarch_body<true>(p);
}

}  // namespace b200
