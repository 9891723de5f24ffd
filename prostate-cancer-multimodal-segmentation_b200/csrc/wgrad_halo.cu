// Weight-gradient GEMM with h-halo reuse of the shifted operand (see WgradHaloParams in igemm.cuh).
// G[tap][p][q] += sum_voxel P[voxel][p] * Q[voxel + s*off(tap)][q], both operands voxel-major (MN-major UMMA operands).
#include <cuda_bf16.h>
#include "igemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {

struct RingW {
    uint32_t stage = 0, phase = 0;
    DEV void advance(uint32_t n) {
        if (++stage == n) { stage = 0; phase ^= 1; }
    }
};

extern "C" __global__ void __launch_bounds__(kThreads, 1) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    // P ring: kWhPBoxes boxes of 16 KB.  A 128-channel P tile takes two boxes per slot (3 slots); a tile of <= 64
    // channels takes one box per slot (5 slots) and the last box, zeroed once, stands in for the empty upper half of
    // every slot (its distance from the slot is the descriptor's leading-dimension offset).  The ring must be deep:
    // a slot is only released by the MMAs of its brick, and with one or two 768-cycle units per brick two slots did
    // not cover the L2 latency (r1 profile: tensor pipe busy 57 % of the kernel, CTAs with one unit twice as slow).
    const uint32_t smem_p = smem_base;
    const uint32_t smem_q = smem_p + kWhPBoxes * kBoxBytes;
    const uint32_t bar_base = smem_q + kWhQStages * kWhQBytes;
    constexpr uint32_t kMaxNP = kWhPBoxes - 1;
    auto pfull = [&](uint32_t s) { return bar_base + 8 * s; };
    auto pempty = [&](uint32_t s) { return bar_base + 8 * (kMaxNP + s); };
    auto qfull = [&](uint32_t s) { return bar_base + 8 * (2 * kMaxNP + s); };
    auto qempty = [&](uint32_t s) { return bar_base + 8 * (2 * kMaxNP + kWhQStages + s); };
    const uint32_t tfull = bar_base + 8 * (2 * kMaxNP + 2 * kWhQStages);
    const uint32_t tmem_ptr_smem = tfull + 8;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.p_map);
        prefetch_tmap(&p.q_map);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < kMaxNP; ++s) { mbar_init(pfull(s), 1); mbar_init(pempty(s), 1); }
        for (uint32_t s = 0; s < kWhQStages; ++s) { mbar_init(qfull(s), 1); mbar_init(qempty(s), 1); }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();   // launch.cuh: the predecessor's results are complete and visible from here on
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    // work decode: CTAs [0, (n_groups-1)*splits*p_tiles) cover the full groups, the rest the last group, which may be
    // split fewer times (it holds one unit when n_units is odd) so that every CTA carries the same MMA work
    int bid = blockIdx.x;
    const int ptile = bid % p.p_tiles; bid /= p.p_tiles;
    const int full = (p.n_groups - 1) * p.splits;
    int group, split, nsplit;
    if (bid < full) {
        group = bid % (p.n_groups - 1);
        split = bid / (p.n_groups - 1);
        nsplit = p.splits;
    } else {
        group = p.n_groups - 1;
        split = bid - full;
        nsplit = p.last_splits;
    }
    const int p0 = ptile * 128;
    const int u0 = group * p.units_per_group;
    const int nun = min(p.units_per_group, p.n_units - u0);
    const int nbricks = p.nbw * p.nbh * p.nbd * p.nbatch;
    const int sgn = p.sgn;
    const bool pair = p.pair != 0;
    const bool one_box = !pair && p.p_extent - p0 <= 64;
    const uint32_t kNP = one_box ? kMaxNP : kWhPBoxes / 2;
    const uint32_t kPSlot = one_box ? kBoxBytes : 2 * kBoxBytes;
    const uint32_t smem_zero = smem_p + kMaxNP * kBoxBytes;

    if (one_box) {
        uint4* z = reinterpret_cast<uint4*>(smem_gen + (smem_zero - smem_base));
        for (int i = threadIdx.x; i < kBoxBytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        __syncthreads();
    }

    if (warp == 0) {
        // ===================================================================== TMA producer
        RingW pp, qp;
        for (int b = split; b < nbricks; b += nsplit) {
            int mt = b;
            const int bw = mt % p.nbw; mt /= p.nbw;
            const int bh = mt % p.nbh; mt /= p.nbh;
            const int bd = mt % p.nbd; mt /= p.nbd;
            const int nb = mt;
            const int w0 = bw * 8, h0 = bh * 16, d0 = pair ? bd - 1 : bd;
            mbar_wait(pempty(pp.stage), pp.phase ^ 1);
            if (elect_one()) {
                const uint32_t fb = pfull(pp.stage);
                mbar_arrive_expect_tx(fb, one_box ? kBoxBytes : 2 * kBoxBytes);
                tma_load_5d(smem_p + pp.stage * kPSlot, &p.p_map, fb, p0, w0, h0, d0, nb);
                if (pair) tma_load_5d(smem_p + pp.stage * kPSlot + kBoxBytes, &p.p_map, fb, p0, w0, h0, d0 + 1, nb);
                else if (!one_box)
                    tma_load_5d(smem_p + pp.stage * kPSlot + kBoxBytes, &p.p_map, fb, p0 + 64, w0, h0, d0, nb);
            }
            __syncwarp();
            pp.advance(kNP);
            for (int ui = 0; ui < nun; ++ui) {
                const int u = u0 + ui;
                int kd, kw, qc;
                if (pair) {
                    const int g2 = u >> 1;
                    kw = g2 / p.q_chunks; qc = g2 - kw * p.q_chunks;
                    kd = (u & 1) ? 0 : 2;   // Q slice d0 - 1 or d0 + 1
                } else {
                    const int tg = u / p.q_chunks;
                    qc = u - tg * p.q_chunks;
                    kd = tg / 3; kw = tg - kd * 3;
                }
                mbar_wait(qempty(qp.stage), qp.phase ^ 1);
                if (elect_one()) {
                    const uint32_t fb = qfull(qp.stage);
                    mbar_arrive_expect_tx(fb, kWhQBytes);
                    tma_load_5d(smem_q + qp.stage * kWhQBytes, &p.q_map, fb, qc * 64, w0 + sgn * (kw - 1), h0 - 1,
                                d0 + sgn * (kd - 1), nb);
                }
                __syncwarp();
                qp.advance(kWhQStages);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        RingW pp, qp;
        // P: two 64-channel atoms 16 KB apart.  Q: three atoms (kh taps) 1 KB (one h line of 8 voxels) apart.
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_p, kBoxBytes, 1024);
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_q, 1024, 1024);
        const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), b_hi = (uint32_t)(b_desc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
        // leading-dimension offset (second 64-channel atom of the M side), bits [16, 30) of the low word: the next box
        // of the slot, or the shared zero box for a one-box tile (distance depends on the slot)
        const uint32_t lbo_mask = 0x3FFFu << 16;
        const uint32_t idesc = make_idesc_bf16(128, 192, 1, 1);
        uint32_t accum = 0;
        for (int b = split; b < nbricks; b += nsplit) {
            mbar_wait(pfull(pp.stage), pp.phase);
            tc_fence_after();
            uint32_t a_lo = a_lo0 + pp.stage * (kPSlot >> 4);
            if (one_box) a_lo = (a_lo & ~lbo_mask) | ((((smem_zero - (smem_p + pp.stage * kPSlot)) >> 4) & 0x3FFFu) << 16);
            for (int ui = 0; ui < nun; ++ui) {
                mbar_wait(qfull(qp.stage), qp.phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t b_lo = b_lo0 + qp.stage * (kWhQBytes >> 4);
                    const uint32_t d_tmem = tmem_base + ui * 192;
                    // 16 voxels = 16 rows x 128 B = 2048 B along K: +128 in the (>>4) address field
                    umma_f16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accum);
#pragma unroll
                    for (int k = 1; k < 8; ++k)
                        umma_f16_lohi(d_tmem, a_lo + 128 * k, a_hi, b_lo + 128 * k, b_hi, idesc, 1u);
                    umma_commit(qempty(qp.stage));
                }
                __syncwarp();
                qp.advance(kWhQStages);
            }
            if (elect_one()) umma_commit(pempty(pp.stage));
            __syncwarp();
            accum = 1u;
            pp.advance(kNP);
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else if (warp >= 4) {
        // ===================================================================== epilogue
        const int q = warp - 4;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float* tile = reinterpret_cast<float*>(smem_gen + (tmem_ptr_smem + 16 - smem_base)) + q * (32 * 33);
        // first P channel of this warp's 32 accumulator rows (pair mode: both row halves hold channels 0..63)
        const int prow0 = p0 + (pair ? (q & 1) * 32 : q * 32);
        const int pidx = prow0 + lane;
        for (int ui = 0; ui < nun; ++ui) {
            const int u = u0 + ui;
            int kd, kw, qc;
            if (pair) {
                const int g2 = u >> 1;
                kw = g2 / p.q_chunks; qc = g2 - kw * p.q_chunks;
                if (u & 1) {
                    if (q >= 2) continue;   // upper rows of a trailing-slice unit pair P(d+1) with Q(d-1): no such tap
                    kd = 0;
                } else {
                    kd = q < 2 ? 2 : 1;
                }
            } else {
                const int tg = u / p.q_chunks;
                qc = u - tg * p.q_chunks;
                kd = tg / 3; kw = tg - kd * 3;
            }
            for (int j = 0; j < 3; ++j) {
                const int kh = sgn > 0 ? j : 2 - j;
                const int tap = kd * 9 + kh * 3 + kw;  // native tap index
                float* obase = p.out + (long long)p.tap_out[tap] * p.st;
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + ui * 192 + j * 64 + half * 32, v);
                    tmem_ld_wait();
                    if (p.sq == 1) {
                        // transpose through shared memory: one RED instruction = 128 contiguous bytes of one row
#pragma unroll
                        for (int c = 0; c < 32; ++c) tile[lane * 33 + c] = __uint_as_float(v[c]);
                        __syncwarp();
                        const int qi = qc * 64 + half * 32 + lane;
                        if (qi < p.q_extent) {
                            const int nrows = min(32, p.p_extent - prow0);
                            float* o = obase + (long long)prow0 * p.sp + qi;
                            for (int rr = 0; rr < nrows; ++rr) atomicAdd(o + (long long)rr * p.sp, tile[rr * 33 + lane]);
                        }
                        __syncwarp();
                    } else if (pidx < p.p_extent) {
                        float* dst = obase + (long long)pidx * p.sp;
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const int qi = qc * 64 + half * 32 + c;
                            if (qi < p.q_extent) atomicAdd(dst + (long long)qi * p.sq, __uint_as_float(v[c]));
                        }
                    }
                }
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

}  // namespace b200
