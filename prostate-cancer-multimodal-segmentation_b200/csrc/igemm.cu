// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a.
//
//   igemm_kernel : 3x3x3 conv fprop and dgrad, 2x2x2-stride-2 transposed conv forward and dgrad
//                  (reference: nn.Conv3d / nn.ConvTranspose3d call sites, models/unet3d.py:29,35,120).
//   wgrad_kernel : weight gradients of both (autograd of the same call sites).
//
// Warp roles (256 threads, 1 CTA / SM): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one
// lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM lane quarter = warp % 4).
#include <cuda_bf16.h>
#include "igemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// Shared-memory carve-up helpers
// ------------------------------------------------------------------------------------------------
struct PipeState {
    uint32_t stage = 0, phase = 0;
    DEV void advance(uint32_t nstages) {
        if (++stage == nstages) { stage = 0; phase ^= 1; }
    }
};


// ------------------------------------------------------------------------------------------------
// First-layer operand built in shared memory (CIN modalities, K = 27 * CIN im2col columns, never materialised).
// The image is VOXEL-contiguous: one 128-byte row per im2col column k = c*27 + kd*9 + kh*3 + kw holding 64 consecutive
// voxels of the brick (SWIZZLE_128B: 16-byte chunk j of row r at ((j ^ (r & 7)) << 4)), two such halves for the 128
// voxels of a brick.  The tensor core reads it as an MN-major A operand in the forward GEMM (M = voxels) and as a
// K-major B operand in the weight-gradient GEMM (K = voxels) — the same bytes.
// Why this layout: bricks are at least 8 voxels wide, so a 16-byte chunk is 8 consecutive voxels of one image row and
// the three kw columns of a (channel, kd, kh) are the same 10 input floats shifted by one — one task loads them once
// (two aligned 16-byte loads + the two halo floats) and writes three chunks.  45 x 16 tasks per brick, ~50 instructions
// each, against one thread per voxel gathering 72 predicated scalars (~1000 instructions per thread and brick, which
// made both first-layer kernels issue-bound: ncu 1.9 IPC, tensor pipe 7 %).
// ------------------------------------------------------------------------------------------------
struct BrickGeom {
    int tw_log2, th_log2, W, H, D, nbatch;
};
template <int CIN>
struct Im2colImage {
    static constexpr int K = 27 * CIN, KPAD = (K + 15) / 16 * 16, KC = (KPAD + 63) / 64, NTRIPLE = 9 * CIN;
    static constexpr int kBuilders = 256;                                 // threads that build the image
    static constexpr int kTasks = (NTRIPLE * 16 + kBuilders - 1) / kBuilders;   // tasks per thread and brick
    // One task = 8 consecutive voxels (chunk cj) of one (channel, kd, kh): f = x[w - 1 .. w + 8] of the source row, zero
    // outside the volume (the convolution's padding); nvalid = voxels of the chunk inside the volume (partial bricks
    // get zero rows: the weight-gradient GEMM sums over voxels).  The loads of a whole brick are issued together and
    // consumed one brick later (see the builders): a load is an L2 / HBM round trip of ~1 us, as long as a brick takes.
    struct Regs {
        float f[kTasks][10];
        int nvalid[kTasks];
    };
    static DEV void load(Regs& r, const float* __restrict__ x, int pt, int w0, int h0, int d0, int nb,
                         const BrickGeom& g) {
        const int hw = g.H * g.W;
        const long long plane = (long long)g.D * hw;
        const bool vec_ok = (g.W & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
        const int wmask = (1 << g.tw_log2) - 1, hmask = (1 << g.th_log2) - 1;
#pragma unroll
        for (int i = 0; i < kTasks; ++i) {
            const int t = pt + i * kBuilders;
            float (&f)[10] = r.f[i];
#pragma unroll
            for (int e = 0; e < 10; ++e) f[e] = 0.f;
            r.nvalid[i] = 0;
            if (t >= NTRIPLE * 16) continue;
            const int q = t >> 4, cj = t & 15;
            const int c = q / 9, r9 = q - 9 * c, kd = r9 / 3, kh = r9 - 3 * kd;
            const int m0 = cj * 8;
            const int w = w0 + (m0 & wmask), h = h0 + ((m0 >> g.tw_log2) & hmask);
            const int d = d0 + (m0 >> (g.tw_log2 + g.th_log2));
            const int hs = h + kh - 1, ds = d + kd - 1;
            const bool row_ok = nb < g.nbatch && w < g.W && h < g.H && d < g.D && (unsigned)hs < (unsigned)g.H &&
                                (unsigned)ds < (unsigned)g.D;
            if (!row_ok) continue;
            const float* src = x + ((long long)nb * CIN + c) * plane + ((long long)ds * hw + hs * g.W + w);
            const int nv = min(8, g.W - w);
            r.nvalid[i] = nv;
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
            if (w > 0) f[0] = __ldg(src - 1);
            if (w + 8 < g.W) f[9] = __ldg(src + 8);
        }
    }
    // addr(k, cj): shared-memory address of the 16-byte chunk cj (0..15: voxels 8*cj .. 8*cj+7) of column k's row
    template <class Addr>
    static DEV void store(const Regs& r, int pt, Addr&& addr) {
#pragma unroll
        for (int i = 0; i < kTasks; ++i) {
            const int t = pt + i * kBuilders;
            if (t >= NTRIPLE * 16) continue;
            const int q = t >> 4, cj = t & 15;
            const float (&f)[10] = r.f[i];
            const int nvalid = r.nvalid[i];
            if (nvalid == 8 || nvalid == 0) {   // (nvalid == 0: f is all zero)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
                    st_shared_v4(addr(3 * q + kw, cj), pack_bf16x2(f[kw], f[kw + 1]), pack_bf16x2(f[kw + 2], f[kw + 3]),
                                 pack_bf16x2(f[kw + 4], f[kw + 5]), pack_bf16x2(f[kw + 6], f[kw + 7]));
            } else {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    uint32_t wd[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        wd[e] = pack_bf16x2(2 * e < nvalid ? f[2 * e + kw] : 0.f,
                                            2 * e + 1 < nvalid ? f[2 * e + 1 + kw] : 0.f);
                    st_shared_v4(addr(3 * q + kw, cj), wd[0], wd[1], wd[2], wd[3]);
                }
            }
        }
        // columns K .. KPAD-1 (zero rows of the packed weights): the MMAs read them, so they must be finite
        for (int t = pt; t < (KPAD - K) * 16; t += kBuilders) st_shared_v4(addr(K + (t >> 4), t & 15), 0u, 0u, 0u, 0u);
    }
};

// ------------------------------------------------------------------------------------------------
// igemm_kernel
// ------------------------------------------------------------------------------------------------
// smem: [stages x A box 16 KB][stages x B tile block_n x 128 B][barriers][tmem ptr][stats scratch][column sums]
// kIm2colC > 0: direct first-layer form — launched with kIm2colThreads threads; warps 8..15 build the A boxes from the
// fp32 network input (kIm2colC modalities), the TMA producer only fetches B (see IgemmParams::x_src)
template <bool kPair, int kIm2colC = 0>
DEV void igemm_body(const IgemmParams& p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform role index
    const int lane = threadIdx.x & 31;
    const uint32_t nst = p.stages;
    const uint32_t a_bytes = p.a_stage_bytes;
    // CTA-pair mode (kPair, launched as clusters of 2): the pair computes two adjacent M tiles against one N tile with
    // tcgen05.mma.cta_group::2 (M = 256).  Each CTA stages its own A box and HALF of the B tile (columns
    // [rank * block_n / 2, +block_n / 2)), which halves the B fill and the B operand reads per SM; the even CTA
    // issues every MMA and commit, both run their own producer and epilogue.
    const uint32_t mmul = kPair ? 2u : 1u;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const uint32_t bn_cta = p.block_n / mmul;          // B columns staged by this CTA
    const uint32_t b_atoms = (bn_cta + 63) >> 6;       // MN-major B: 64-column atoms per stage
    const uint32_t bt_bytes = p.b_mn ? 8192u : bn_cta * 128;      // one tap of B (MN-major: of one atom)
    const uint32_t b_atom_bytes = 8192u * p.group;
    const uint32_t b_bytes = p.b_mn ? b_atoms * b_atom_bytes : bt_bytes * p.group;  // B bytes per stage
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + nst * a_bytes;
    const uint32_t smem_c = smem_b + nst * b_bytes;    // epilogue v2 staging: (block_n / 64) boxes of 128 rows x 128 B
    const uint32_t c_one = p.epi_v2 ? ((p.block_n + 63) >> 6) * kBoxBytes : 0;   // one staging tile
    const uint32_t c_bytes = c_one * (p.c_bufs > 1 ? 2u : 1u);
    const uint32_t bar_base = smem_c + c_bytes;        // 8-byte aligned (multiple of 1024)
    // barrier layout: full[nst], empty[nst], tmem_full[2], tmem_empty[2]
    auto full_bar = [&](uint32_t s) { return bar_base + 8 * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8 * (nst + s); };
    auto tfull_bar = [&](uint32_t s) { return bar_base + 8 * (2 * nst + s); };
    auto tempty_bar = [&](uint32_t s) { return bar_base + 8 * (2 * nst + 2 + s); };
    const uint32_t tmem_ptr_smem = bar_base + 8 * (2 * nst + 4);
    auto sink_bar = [&](uint32_t s) { return tmem_ptr_smem + 16 + 8 * s; };   // dev ablation 3 only
    const uint32_t scratch_off = (tmem_ptr_smem + 16 + 64 - smem_base + 15u) & ~15u;
    float* scratch = reinterpret_cast<float*>(smem_gen + scratch_off);  // [4 warps][256 cols][2]
    float* colacc = scratch + 4 * 256 * 2;                              // [ncols <= kMaxStatCols][2], per-CTA running sums
    float* colvec = colacc + 2 * kMaxStatCols;                          // epilogue v2: [2][256] per-tile bias | scale, shift

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.b_map);
        prefetch_tmap(&p.a_map[0]);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < nst; ++s) {
            mbar_init(full_bar(s), kIm2colC ? 1 + 8 : 1);   // + one arrival per A-building warp
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 128 * mmul);   // pair: the epilogue threads of both CTAs arrive on the leader's
        }
        for (uint32_t s = 0; s < nst; ++s) mbar_init(sink_bar(s), 1);   // dev ablation 3: loads the MMAs do not wait for
        fence_mbar_init();
    }
    if (warp == 2) {
        if (kPair) {
            tmem_alloc_pair(tmem_ptr_smem, 512);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr_smem, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    pdl_wait();   // launch.cuh: the predecessor's results are complete and visible from here on
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

    // work units: (M tile, N tile), or in pair mode (pair of adjacent M tiles, N tile); a pair walks the same units,
    // CTA `rank` owning M tile 2 * m_unit + rank (past the last tile: all-out-of-range coordinates, nothing stored)
    const int m_tiles = p.nbw * p.nbh * p.nbd * p.nbatch;
    const int m_units = (m_tiles + (int)mmul - 1) / (int)mmul;
    const int total_tiles = m_units * p.n_tiles;
    const int unit0 = blockIdx.x / mmul, unit_stride = gridDim.x / mmul;
    // split-K: unit = tile * splits + split; split s takes tap groups [s * G / S, (s + 1) * G / S)
    const int nsplit = p.splits > 1 ? p.splits : 1;
    const int total_units = total_tiles * nsplit;
    const int tap_groups = p.ntaps / p.group;

    if (warp == 0) {
        // ===================================================================== TMA producer
        // The whole warp walks the loop (uniform control flow, barrier polls by all lanes); one elected lane issues.
        PipeState ps;
        uint32_t sink_uses = 0;   // dev ablation 3: stages issued so far
        const int kc_blocks = p.kc_blocks, n_tiles = p.n_tiles, group = p.group;
        for (int unit = unit0; unit < total_units; unit += unit_stride) {
            const int tile = unit / nsplit, sp = unit - tile * nsplit;
            const int tap_lo = sp * tap_groups / nsplit, tap_hi = (sp + 1) * tap_groups / nsplit;
            int mt = tile / n_tiles;
            const int n_tile = tile - mt * n_tiles;
            mt = mt * (int)mmul + (int)rank;
            const int bw = mt % p.nbw; mt /= p.nbw;
            const int bh = mt % p.nbh; mt /= p.nbh;
            const int bd = mt % p.nbd; mt /= p.nbd;
            const int nb = mt;
            const int w0 = bw << p.tw_log2, h0 = bh << p.th_log2, d0 = bd << p.td_log2;
            const int n0 = n_tile * p.block_n;
            for (int tap = tap_lo; tap < tap_hi; ++tap) {
                const void* amap = &p.a_map[p.a_map_of_tap[tap]];
                const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap], cd = d0 + p.tap_dd[tap];
                for (int kc = 0; kc < kc_blocks; ++kc) {
                    mbar_wait(empty_bar(ps.stage), ps.phase ^ 1);
                    if (elect_one()) {
                        uint32_t fb = full_bar(ps.stage);
                        if (B200_ABLATE(p) == 3) {   // loads issued, the MMAs do not wait for them: contention without latency
                            if (rank == 0) mbar_arrive(fb);
                            fb = sink_bar(ps.stage);
                            if (rank == 0 && sink_uses >= nst)   // this slot's previous loads have landed
                                while (!mbar_try_wait(fb, ((sink_uses / nst) - 1) & 1)) {}
                        }
                        if (B200_ABLATE(p) == 1) {
                            if (rank == 0) mbar_arrive(fb);
                        } else if (!kPair) {
                            mbar_arrive_expect_tx(fb, kIm2colC ? b_bytes : a_bytes + b_bytes);
                            if (!kIm2colC) tma_load_5d(smem_a + ps.stage * a_bytes, amap, fb, kc * 64, cw, ch, cd, nb);
                            if (p.b_mn) {
                                for (uint32_t a = 0; a < b_atoms; ++a)
                                    tma_load_3d(smem_b + ps.stage * b_bytes + a * b_atom_bytes, &p.b_map, fb,
                                                n0 + a * 64, kc * 64, tap * group);
                            } else {
                                tma_load_3d(smem_b + ps.stage * b_bytes, &p.b_map, fb, kc * 64, n0, tap * group);
                            }
                        } else {
                            // both CTAs' bytes are counted on the leader's barrier (the pair form of the TMA load)
                            if (rank == 0) mbar_arrive_expect_tx(fb, 2 * (a_bytes + b_bytes));
                            tma_load_5d_pair(smem_a + ps.stage * a_bytes, amap, fb, kc * 64, cw, ch, cd, nb);
                            const int nc = n0 + (int)(rank * bn_cta);
                            if (p.b_mn) {
                                for (uint32_t a = 0; a < b_atoms; ++a)
                                    tma_load_3d_pair(smem_b + ps.stage * b_bytes + a * b_atom_bytes, &p.b_map, fb,
                                                     nc + a * 64, kc * 64, tap * group);
                            } else {
                                tma_load_3d_pair(smem_b + ps.stage * b_bytes, &p.b_map, fb, kc * 64, nc, tap * group);
                            }
                        }
                    }
                    __syncwarp();
                    ++sink_uses;
                    ps.advance(nst);
                }
            }
        }
        if (B200_ABLATE(p) == 3 && rank == 0 && elect_one()) {   // no load may be in flight when the CTAs exit
            for (uint32_t s = 0; s < nst; ++s) {
                const uint32_t uses = sink_uses / nst + (s < sink_uses % nst ? 1u : 0u);
                if (uses > 0) while (!mbar_try_wait(sink_bar(s), (uses - 1) & 1)) {}
            }
        }
        __syncwarp();
    } else if (warp == 1 && rank == 0) {
        // ===================================================================== MMA issuer (pair mode: leader CTA only)
        // Lean loop: descriptors are base + stage offset (low word only), no divisions, one elected lane issues.
        // (direct first-layer form: A is the voxel-contiguous image above — MN-major, voxel halves 8 KB apart, one K step
        //  = 16 rows of 128 B)
        const uint32_t idesc = make_idesc_bf16(128 * mmul, p.block_n, kIm2colC ? 1u : 0u, p.b_mn ? 1u : 0u);
        const bool pair = kPair;
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_a, kIm2colC ? 8192u : 0u, 1024);
        constexpr uint32_t kAk = kIm2colC ? 128u : 2u;
        // K-major B: rows = N, 128 B = 64 K.  MN-major B: rows = K, 128 B = 64 N, atoms of 64 N are LBO apart.
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_b, p.b_mn ? b_atom_bytes : 0, 1024);
        const uint32_t kinc = p.b_mn ? 128u : 2u;  // one K step (16): 16 rows x 128 B, or 32 B inside the swizzle row
        const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4, bt_step = bt_bytes >> 4;
        const int kc_blocks = p.kc_blocks, group = p.group;
        const uint32_t goff1 = p.a_goff[1], goff2 = p.a_goff[2], goff0 = p.a_goff[0];
        const int nk_last = ((p.cin - (kc_blocks - 1) * 64) + 15) >> 4;  // K steps of the last channel block (1..4)
        PipeState ps;
        int iter = 0;
        for (int unit = unit0; unit < total_units; unit += unit_stride, ++iter) {
            const int sp = unit % nsplit;
            const int tap_lo = sp * tap_groups / nsplit, tap_hi = (sp + 1) * tap_groups / nsplit;
            const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * p.block_n;
            uint32_t accum = 0;
            for (int tap = tap_lo; tap < tap_hi; ++tap) {
                for (int kc = 0; kc < kc_blocks; ++kc) {
                    const int nk = (kc == kc_blocks - 1) ? nk_last : 4;
                    mbar_wait(full_bar(ps.stage), ps.phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_st = a_desc0 + ps.stage * a_step;
                        const uint64_t b_st = b_desc0 + ps.stage * b_step;
                        for (int g = 0; g < group && B200_ABLATE(p) != 2; ++g) {
                            const uint64_t a_desc = a_st + (g == 0 ? goff0 : (g == 1 ? goff1 : goff2));
                            const uint64_t b_desc = b_st + g * bt_step;
                            // 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (>>4) address field
                            if (!pair) {
                                umma_f16(d_tmem, a_desc, b_desc, idesc, g == 0 ? accum : 1u);
                                if (nk > 1) umma_f16(d_tmem, a_desc + kAk, b_desc + kinc, idesc, 1u);
                                if (nk > 2) umma_f16(d_tmem, a_desc + 2 * kAk, b_desc + 2 * kinc, idesc, 1u);
                                if (nk > 3) umma_f16(d_tmem, a_desc + 3 * kAk, b_desc + 3 * kinc, idesc, 1u);
                            } else {
                                umma_f16_pair(d_tmem, a_desc, b_desc, idesc, g == 0 ? accum : 1u);
                                if (nk > 1) umma_f16_pair(d_tmem, a_desc + 2, b_desc + kinc, idesc, 1u);
                                if (nk > 2) umma_f16_pair(d_tmem, a_desc + 4, b_desc + 2 * kinc, idesc, 1u);
                                if (nk > 3) umma_f16_pair(d_tmem, a_desc + 6, b_desc + 3 * kinc, idesc, 1u);
                            }
                        }
                        if (pair) umma_commit_pair(empty_bar(ps.stage), 3u);   // frees the slot in both CTAs
                        else umma_commit(empty_bar(ps.stage));
                    }
                    __syncwarp();
                    accum = 1u;
                    ps.advance(nst);
                }
            }
            if (elect_one()) {
                if (pair) umma_commit_pair(tfull_bar(acc), 3u);
                else umma_commit(tfull_bar(acc));
            }
            __syncwarp();
        }
    } else if (kIm2colC > 0 && warp >= 8) {
        // ===================================================================== A builders (direct first-layer form)
        // One brick = Img::KC consecutive ring slots (64 im2col columns each): the builders take all of them, write the
        // whole image, and every warp arrives once per slot after its rows are visible to the tensor core's
        // (async-proxy) reads.
        PipeState ps;
        const int pt = threadIdx.x - 256;
        const int n_tiles = p.n_tiles;
        const BrickGeom geom{p.tw_log2, p.th_log2, p.W, p.H, p.D, p.nbatch};
        using Img = Im2colImage<kIm2colC ? kIm2colC : 1>;
        auto load_tile = [&](typename Img::Regs& r, int tile) {
            if (B200_ABLATE(p) == 5) {   // dev: no input loads (zero image)
#pragma unroll
                for (int i = 0; i < Img::kTasks; ++i) {
                    r.nvalid[i] = 0;
#pragma unroll
                    for (int e = 0; e < 10; ++e) r.f[i][e] = 0.f;
                }
                return;
            }
            int mt = tile / n_tiles;
            const int bw = mt % p.nbw; mt /= p.nbw;
            const int bh = mt % p.nbh; mt /= p.nbh;
            const int bd = mt % p.nbd; mt /= p.nbd;
            Img::load(r, p.x_src, pt, bw << p.tw_log2, bh << p.th_log2, bd << p.td_log2, mt, geom);
        };
        auto store_tile = [&](const typename Img::Regs& r) {
            uint32_t slot[Img::KC], bars[Img::KC];
#pragma unroll
            for (int i = 0; i < Img::KC; ++i) {
                mbar_wait(empty_bar(ps.stage), ps.phase ^ 1);
                slot[i] = smem_a + ps.stage * a_bytes;
                bars[i] = full_bar(ps.stage);
                ps.advance(nst);
            }
            Img::store(r, pt, [&](int k, int cj) -> uint32_t {
                static_assert(Img::KC <= 3, "at most 192 im2col columns");
                const uint32_t base = k < 64 ? slot[0] : (k < 128 ? slot[Img::KC > 1 ? 1 : 0] : slot[Img::KC > 2 ? 2 : 0]);
                return base + (uint32_t)(cj >> 3) * 8192u + (uint32_t)(k & 63) * 128u +
                       ((uint32_t)((cj & 7) ^ (k & 7)) << 4);
            });
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < Img::KC; ++i) mbar_arrive(bars[i]);
            }
        };
        // two register sets, ping-pong: the loads of brick i + 1 are in flight while brick i is written
        typename Img::Regs ra, rb;
        int tile = unit0;
        if (tile < total_tiles) load_tile(ra, tile);
        while (tile < total_tiles) {
            const int t1 = tile + unit_stride;
            if (t1 < total_tiles) load_tile(rb, t1);
            store_tile(ra);
            if (t1 >= total_tiles) break;
            const int t2 = t1 + unit_stride;
            if (t2 < total_tiles) load_tile(ra, t2);
            store_tile(rb);
            tile = t2;
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue
        const int q = warp - 4;  // == warp % 4: TMEM lane quarter
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 128;  // 0..127
        const int rw = row & ((1 << p.tw_log2) - 1);
        const int rh = (row >> p.tw_log2) & ((1 << p.th_log2) - 1);
        const int rd = row >> (p.tw_log2 + p.th_log2);
        if (p.mode == EPI_BIAS_STATS) {
            for (int i = et; i < 2 * p.ncols; i += 128) colacc[i] = 0.f;
            named_bar_sync(1, 128);
        }
        int iter = 0;
        if (p.epi_v2) {
            // ------------------------------------------------------------- epilogue v2
            // TMEM -> registers (32 columns per load) -> bias / affine+ReLU -> bf16 -> swizzled shared-memory tile ->
            // one TMA store per 64-channel box (clips rows/columns outside the tensor); BatchNorm partial sums are
            // read back from the staged tile (conflict-free word reads) instead of 30 shuffles per 16 columns.
            const int nbox = (p.block_n + 63) >> 6, nslab = p.block_n >> 5;
            const int mode = p.mode;
            const uint32_t sw = row & 7;
            const bool two_bufs = p.c_bufs > 1;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride, ++iter) {   // (never split: host)
                const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
                const int m_unit = tile / p.n_tiles;
                const int n_tile = tile - m_unit * p.n_tiles;
                int mt = m_unit * (int)mmul + (int)rank;
                const int bw = mt % p.nbw; mt /= p.nbw;
                const int bh = mt % p.nbh; mt /= p.nbh;
                const int bd = mt % p.nbd; mt /= p.nbd;
                const int nb = mt;
                const int w0 = bw << p.tw_log2, h0 = bh << p.th_log2, d0 = bd << p.td_log2;
                const bool row_ok = (w0 + rw) < p.W && (h0 + rh) < p.H && (d0 + rd) < p.D && nb < p.nbatch;
                const int n0 = n_tile * p.block_n;
                if (mode != EPI_PLAIN) {
                    for (int c = et; c < p.block_n; c += 128) {
                        const bool ok = n0 + c < p.ncols;
                        colvec[c] = ok ? __ldg(p.vec0 + n0 + c) : 0.f;
                        if (mode == EPI_AFFINE_RELU) colvec[256 + c] = ok ? __ldg(p.vec1 + n0 + c) : 0.f;
                    }
                }
                // the staging tile about to be written is free: the TMA store that last read it (of the previous tile,
                // or with two tiles of the one before) has finished reading
                const uint32_t cbuf = smem_c + ((two_bufs && (iter & 1)) ? c_one : 0u);
                const uint32_t row_smem = cbuf + row * 128;
                if (et == 0) {
                    if (two_bufs) bulk_wait_read1(); else bulk_wait_read0();
                }
                named_bar_sync(1, 128);
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                if (B200_ABLATE(p) == 4) {   // dev: the epilogue only hands the accumulator back
                    tc_fence_before();
                    if (kPair) mbar_arrive_leader(tempty_bar(acc));
                    else mbar_arrive(tempty_bar(acc));
                    continue;
                }
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
                for (int j = 0; j < nslab; ++j) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + j * 32, v);
                    tmem_ld_wait();
                    const float* cv = colvec + j * 32;
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
                        if (mode == EPI_AFFINE_RELU) {
                            a = fmaxf(fmaf(a, cv[2 * i], cv[256 + 2 * i]), 0.f);
                            b = fmaxf(fmaf(b, cv[2 * i + 1], cv[256 + 2 * i + 1]), 0.f);
                        } else if (mode != EPI_PLAIN) {
                            a += cv[2 * i];
                            b += cv[2 * i + 1];
                        }
                        pk[i] = row_ok ? pack_bf16x2(a, b) : 0u;
                    }
                    const uint32_t box_addr = row_smem + (j >> 1) * kBoxBytes;
                    const uint32_t c0 = (j & 1) * 4;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        st_shared_v4(box_addr + (((c0 + c) ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2],
                                     pk[4 * c + 3]);
                }
                tc_fence_before();
                if (kPair) mbar_arrive_leader(tempty_bar(acc));
                else mbar_arrive(tempty_bar(acc));  // TMEM stage drained
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the TMA engine
                named_bar_sync(1, 128);
                if (et == 0) {
                    for (int b = 0; b < nbox; ++b) {
                        const int col0 = n0 + b * 64;
                        if (col0 < p.ncols) {
                            const int g = col0 / p.cols_per_group;
                            tma_store_5d(&p.c_map[g], cbuf + b * kBoxBytes, col0 - g * p.cols_per_group, w0, h0, d0,
                                         nb);
                        }
                    }
                    bulk_commit();
                }
                if (mode == EPI_BIAS_STATS) {
                    // warp q sums rows 32q..32q+31 of column pair `lane` of every box
                    for (int b = 0; b < nbox; ++b) {
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                        const uint32_t base = cbuf + b * kBoxBytes + (q * 32) * 128 + (lane & 3) * 4;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const uint32_t wv = ld_shared_b32(base + r * 128 + ((((uint32_t)lane >> 2) ^ (r & 7)) << 4));
                            const float lo = __uint_as_float(wv << 16), hi = __uint_as_float(wv & 0xffff0000u);
                            s0 += lo; q0 = fmaf(lo, lo, q0);
                            s1 += hi; q1 = fmaf(hi, hi, q1);
                        }
                        float4* dst = reinterpret_cast<float4*>(scratch + ((q * 256 + b * 64 + 2 * lane) * 2));
                        *dst = make_float4(s0, q0, s1, q1);
                    }
                    named_bar_sync(1, 128);
                    for (int cl = et; cl < p.block_n; cl += 128) {
                        const int col = n0 + cl;
                        if (col < p.ncols) {
                            float a = 0.f, b2 = 0.f;
#pragma unroll
                            for (int w4 = 0; w4 < 4; ++w4) {
                                a += scratch[(w4 * 256 + cl) * 2 + 0];
                                b2 += scratch[(w4 * 256 + cl) * 2 + 1];
                            }
                            colacc[2 * col] += a;
                            colacc[2 * col + 1] += b2;
                        }
                    }
                }
            }
            if (et == 0) bulk_wait0();  // all output tiles are in global memory before the CTA exits
        } else
        for (int unit = unit0; unit < total_units; unit += unit_stride, ++iter) {
            const int tile = unit / nsplit;
            const uint32_t acc = iter & 1, acc_phase = (iter >> 1) & 1;
            const int m_unit = tile / p.n_tiles;
            const int n_tile = tile - m_unit * p.n_tiles;
            int mt = m_unit * (int)mmul + (int)rank;
            const int bw = mt % p.nbw; mt /= p.nbw;
            const int bh = mt % p.nbh; mt /= p.nbh;
            const int bd = mt % p.nbd; mt /= p.nbd;
            const int nb = mt;
            const int gw = (bw << p.tw_log2) + rw, gh = (bh << p.th_log2) + rh, gd = (bd << p.td_log2) + rd;
            const bool row_ok = gw < p.W && gh < p.H && gd < p.D && nb < p.nbatch;
            const int n0 = n_tile * p.block_n;

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
            const int nchunks = p.block_n >> 4;
            if (p.mode == EPI_SPLITK) {
                // fp32 partial tile -> this split's slice of the workspace, row of this voxel (plain 16-byte stores)
                const long long vox = (((long long)nb * p.D + gd) * p.H + gh) * p.W + gw;
                float* wrow = p.ws + ((long long)(unit - tile * nsplit) * p.ws_slice_vox + vox) * p.ncols;
                for (int c = 0; c < nchunks; ++c) {
                    uint32_t v[16];
                    tmem_ld16(t_addr + c * 16, v);
                    tmem_ld_wait();
                    const int col0 = n0 + c * 16;
                    if (row_ok && col0 < p.ncols) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4*>(wrow + col0 + 4 * j) =
                                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    }
                }
                tc_fence_before();
                if (kPair) mbar_arrive_leader(tempty_bar(acc));
                else mbar_arrive(tempty_bar(acc));
                continue;
            }
            for (int c = 0; c < nchunks; ++c) {
                uint32_t v[16];
                tmem_ld16(t_addr + c * 16, v);
                tmem_ld_wait();
                const int col0 = n0 + c * 16;
                const bool col_ok = col0 < p.ncols;
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
                if (p.mode == EPI_BIAS_STATS || p.mode == EPI_BIAS) {
                    if (col_ok) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] += __ldg(p.vec0 + col0 + j);
                    }
                } else if (p.mode == EPI_AFFINE_RELU) {
                    if (col_ok) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            f[j] = fmaxf(fmaf(f[j], __ldg(p.vec0 + col0 + j), __ldg(p.vec1 + col0 + j)), 0.f);
                    }
                }
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                if (row_ok && col_ok) {
                    const int g = col0 / p.cols_per_group;
                    const int cc = col0 - g * p.cols_per_group;
                    const long long off = nb * p.out_sn + (long long)(gd * p.out_mul + p.out_od[g]) * p.out_sd +
                                          (long long)(gh * p.out_mul + p.out_oh[g]) * p.out_sh +
                                          (long long)(gw * p.out_mul + p.out_ow[g]) * p.out_sw + cc;
                    uint4* dst = reinterpret_cast<uint4*>(p.out + off);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
                if (p.mode == EPI_BIAS_STATS) {
                    // statistics of the *stored* (bf16-rounded) values, rows outside the volume excluded
                    float s[16], ss[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        __nv_bfloat162 b2 = *reinterpret_cast<__nv_bfloat162*>(&pk[j]);
                        const float lo = row_ok ? __low2float(b2) : 0.f, hi = row_ok ? __high2float(b2) : 0.f;
                        s[2 * j] = lo; s[2 * j + 1] = hi;
                        ss[2 * j] = lo * lo; ss[2 * j + 1] = hi * hi;
                    }
                    // transpose-reduce over the 32 lanes: 16 -> 8 -> 4 -> 2 -> 1 values per lane
#pragma unroll
                    for (int width = 8, bit = 16; width >= 1; width >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0;
#pragma unroll
                        for (int j = 0; j < width; ++j) {
                            const float keep_s = up ? s[j + width] : s[j], send_s = up ? s[j] : s[j + width];
                            const float keep_q = up ? ss[j + width] : ss[j], send_q = up ? ss[j] : ss[j + width];
                            s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
                            ss[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
                        }
                    }
                    s[0] += __shfl_xor_sync(0xffffffffu, s[0], 1);
                    ss[0] += __shfl_xor_sync(0xffffffffu, ss[0], 1);
                    // lanes 2c and 2c+1 now hold column c of this 16-column chunk
                    const int cl = c * 16 + (lane >> 1);
                    scratch[(q * 256 + cl) * 2 + (lane & 1)] = (lane & 1) ? ss[0] : s[0];
                }
            }
            // TMEM stage drained: hand it back to the MMA warp (pair mode: of the leader CTA)
            tc_fence_before();
            if (kPair) mbar_arrive_leader(tempty_bar(acc));
            else mbar_arrive(tempty_bar(acc));

            if (p.mode == EPI_BIAS_STATS) {
                named_bar_sync(1, 128);
                for (int cl = et; cl < p.block_n; cl += 128) {
                    const int col = n0 + cl;
                    if (col < p.ncols) {
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4) {
                            a += scratch[(w4 * 256 + cl) * 2 + 0];
                            b += scratch[(w4 * 256 + cl) * 2 + 1];
                        }
                        colacc[2 * col] += a;  // column `col` of a tile is always reduced by thread cl % 128
                        colacc[2 * col + 1] += b;
                    }
                }
                named_bar_sync(1, 128);
            }
        }
        if (p.mode == EPI_BIAS_STATS) {
            // one partial row per CTA: stats[blockIdx.x][ncols][2]
            named_bar_sync(1, 128);
            float* dst = p.stats + (long long)blockIdx.x * p.ncols * 2;
            for (int i = et; i < 2 * p.ncols; i += 128) dst[i] = colacc[i];
        }
    }

    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();   // no CTA of a pair leaves while its peer may still signal it
    if (warp == 2) {
        tc_fence_after();
        if (kPair) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// The single-CTA form contains no cluster instruction at all (a kernel that does cannot be launched without a
// cluster configuration); the pair form is compiled for clusters of two.
extern "C" __global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
    igemm_body<false>(p);
}
extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    igemm_pair_kernel(const __grid_constant__ IgemmParams p) {
    igemm_body<true>(p);
}
// first layer of the 5-modality network (models/unet3d.py:194 inc = DoubleConv3D(5, 64)), fp32 input read directly
extern "C" __global__ void __launch_bounds__(kIm2colThreads, 1)
    igemm_im2col5_kernel(const __grid_constant__ IgemmParams p) {
    igemm_body<false, 5>(p);
}

// ------------------------------------------------------------------------------------------------
// wgrad_kernel
// ------------------------------------------------------------------------------------------------
// smem: [2 x P slot 32 KB (two 64-channel boxes)][2 x Q slot 64 KB (up to four boxes)][barriers][tmem ptr]
// One CTA = (p tile of 128 channels, group of <= 8 column blocks, voxel split); accumulators for the whole
// group stay in TMEM (<= 512 columns) across all bricks of the split, then are added atomically to G.
// kIm2colC > 0: first-layer form — the Q operand (im2col rows of the fp32 network input, kIm2colC modalities) is built
// in shared memory by warps 8..15 instead of being loaded by TMA (kernel launched with kIm2colThreads threads)
template <int kIm2colC>
DEV void wgrad_body(const WgradParams& p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    // (first-layer form: three column blocks per slot and shallower P slots leave ~70 KB of the SM to L1, which the
    //  27-fold re-read of the fp32 input by the Q builders lives on)
    constexpr uint32_t kPSlot = 2 * kBoxBytes, kQSlot = (kIm2colC ? 3 : 4) * kBoxBytes, kNP = 2, kNQ = 2;
    // first-layer form: the Q slot holds the voxel-contiguous image (Im2colImage): two halves of KPAD rows x 128 B
    using Img = Im2colImage<kIm2colC ? kIm2colC : 1>;
    constexpr uint32_t kQHalf = Img::KPAD * 128;
    static_assert(kIm2colC == 0 || 2 * kQHalf <= kQSlot, "im2col image does not fit the Q slot");
    const uint32_t smem_p = smem_base;
    const uint32_t smem_q = smem_p + kNP * kPSlot;
    const uint32_t bar_base = smem_q + kNQ * kQSlot;
    auto pfull = [&](uint32_t s) { return bar_base + 8 * s; };
    auto pempty = [&](uint32_t s) { return bar_base + 8 * (kNP + s); };
    auto qfull = [&](uint32_t s) { return bar_base + 8 * (2 * kNP + s); };
    auto qempty = [&](uint32_t s) { return bar_base + 8 * (2 * kNP + kNQ + s); };
    const uint32_t tfull = bar_base + 8 * (2 * kNP + 2 * kNQ);
    const uint32_t tmem_ptr_smem = tfull + 8;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&p.p_map);
        prefetch_tmap(&p.q_map[0]);
    }
    if (warp == 1 && lane == 0) {
        for (uint32_t s = 0; s < kNP; ++s) { mbar_init(pfull(s), 1); mbar_init(pempty(s), 1); }
        for (uint32_t s = 0; s < kNQ; ++s) { mbar_init(qfull(s), kIm2colC ? 8 : 1); mbar_init(qempty(s), 1); }
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

    // work decode: blockIdx = (split * n_groups + group) * p_tiles + ptile
    int bid = blockIdx.x;
    const int ptile = bid % p.p_tiles; bid /= p.p_tiles;
    const int group = bid % p.n_groups; bid /= p.n_groups;
    const int split = bid;
    const int p0 = ptile * 128;
    const int cb0 = group * p.cb_per_group;
    const int ncb = min(p.cb_per_group, p.n_colblocks - cb0);
    const int nsg = (ncb + 3) >> 2;  // slot groups of up to 4 column blocks (UMMA N up to 256)
    const int nbricks = p.nbw * p.nbh * p.nbd * p.nbatch;

    if (warp == 0) {
        // ===================================================================== TMA producer (whole warp, elected issue)
        PipeState pp, qp;
        for (int b = split; b < nbricks; b += p.splits) {
            int mt = b;
            const int bw = mt % p.nbw; mt /= p.nbw;
            const int bh = mt % p.nbh; mt /= p.nbh;
            const int bd = mt % p.nbd; mt /= p.nbd;
            const int nb = mt;
            const int w0 = bw * p.tw, h0 = bh * p.th, d0 = bd * p.td;
            mbar_wait(pempty(pp.stage), pp.phase ^ 1);
            if (elect_one()) {
                const uint32_t fb = pfull(pp.stage);
                mbar_arrive_expect_tx(fb, 2 * kBoxBytes);
                tma_load_5d(smem_p + pp.stage * kPSlot, &p.p_map, fb, p0, w0, h0, d0, nb);
                tma_load_5d(smem_p + pp.stage * kPSlot + kBoxBytes, &p.p_map, fb, p0 + 64, w0, h0, d0, nb);
            }
            __syncwarp();
            pp.advance(kNP);
            for (int sg = 0; sg < nsg && kIm2colC == 0; ++sg) {
                const int nb4 = min(4, ncb - sg * 4);
                mbar_wait(qempty(qp.stage), qp.phase ^ 1);
                if (elect_one()) {
                    const uint32_t fb = qfull(qp.stage);
                    mbar_arrive_expect_tx(fb, nb4 * kBoxBytes);
                    for (int i = 0; i < nb4; ++i) {
                        const int cb = cb0 + sg * 4 + i;
                        const int tap = p.tap_minor ? cb % p.ntaps : cb / p.q_chunks;
                        const int qc = p.tap_minor ? cb / p.ntaps : cb - tap * p.q_chunks;
                        tma_load_5d(smem_q + qp.stage * kQSlot + i * kBoxBytes, &p.q_map[p.q_map_of_tap[tap]], fb,
                                    qc * 64, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], d0 + p.tap_dd[tap], nb);
                    }
                }
                __syncwarp();
                qp.advance(kNQ);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (whole warp, elected issue)
        PipeState pp, qp;
        // MN-major SWIZZLE_128B operands: 64-channel atoms (16 KB boxes) LBO apart, 8-voxel groups SBO apart
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_p, kBoxBytes, 1024);
        // (first-layer form: Q is K-major — rows = im2col columns, 128 B = 64 voxels, the second 64 voxels kQHalf on)
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_q, kIm2colC ? 0u : (uint32_t)kBoxBytes, 1024);
        uint32_t accum = 0;
        for (int b = split; b < nbricks; b += p.splits) {
            mbar_wait(pfull(pp.stage), pp.phase);
            tc_fence_after();
            const uint64_t a_desc = a_desc0 + pp.stage * (kPSlot >> 4);
            for (int sg = 0; sg < nsg; ++sg) {
                const int nb4 = min(4, ncb - sg * 4);
                const uint32_t idesc = kIm2colC ? make_idesc_bf16(128, Img::KPAD, 1, 0) : make_idesc_bf16(128, 64 * nb4, 1, 1);
                mbar_wait(qfull(qp.stage), qp.phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t b_desc = b_desc0 + qp.stage * (kQSlot >> 4);
                    const uint32_t d_tmem = tmem_base + sg * 256;
                    // 16 voxels = 16 rows x 128 B = 2048 B along K: +128 in the (>>4) address field
                    umma_f16(d_tmem, a_desc, b_desc, idesc, accum);
#pragma unroll
                    for (int k = 1; k < 8; ++k)
                        umma_f16(d_tmem, a_desc + 128 * k,
                                 kIm2colC ? b_desc + (k >> 2) * (kQHalf >> 4) + (k & 3) * 2 : b_desc + 128 * k, idesc, 1u);
                    umma_commit(qempty(qp.stage));
                }
                __syncwarp();
                qp.advance(kNQ);
            }
            if (elect_one()) umma_commit(pempty(pp.stage));
            __syncwarp();
            accum = 1u;
            pp.advance(kNP);
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
    } else if (kIm2colC > 0 && warp >= 8) {
        // ===================================================================== Q builders (first-layer form)
        PipeState qp;
        const int pt = threadIdx.x - 256;
        const BrickGeom geom{p.tw_log2, p.th_log2, p.W, p.H, p.D, p.nbatch};
        auto load_brick = [&](typename Img::Regs& r, int b) {
            int mt = b;
            const int bw = mt % p.nbw; mt /= p.nbw;
            const int bh = mt % p.nbh; mt /= p.nbh;
            const int bd = mt % p.nbd; mt /= p.nbd;
            Img::load(r, p.x_src, pt, bw * p.tw, bh * p.th, bd * p.td, mt, geom);
        };
        auto store_brick = [&](const typename Img::Regs& r) {
            mbar_wait(qempty(qp.stage), qp.phase ^ 1);
            const uint32_t slot = smem_q + qp.stage * kQSlot;
            Img::store(r, pt, [&](int k, int cj) -> uint32_t {
                return slot + (uint32_t)(cj >> 3) * kQHalf + (uint32_t)k * 128u + ((uint32_t)((cj & 7) ^ (k & 7)) << 4);
            });
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(qfull(qp.stage));
            qp.advance(kNQ);
        };
        // two register sets, ping-pong: the loads of brick i + 1 are in flight while brick i is written
        typename Img::Regs ra, rb;
        int b0 = split;
        if (b0 < nbricks) load_brick(ra, b0);
        while (b0 < nbricks) {
            const int b1 = b0 + p.splits;
            if (b1 < nbricks) load_brick(rb, b1);
            store_brick(ra);
            if (b1 >= nbricks) break;
            const int b2 = b1 + p.splits;
            if (b2 < nbricks) load_brick(ra, b2);
            store_brick(rb);
            b0 = b2;
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================================== epilogue
        const int q = warp - 4;
        const int row = q * 32 + lane;
        const int pidx = p0 + row;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (p.sq == 1) {
            // q is the contiguous output axis: transpose each 32 x 32 block through shared memory so that one RED
            // instruction covers 128 contiguous bytes of one output row (1 request instead of 32 scattered sectors)
            float* tile = reinterpret_cast<float*>(smem_gen + (tmem_ptr_smem + 16 - smem_base)) + q * (32 * 33);
            for (int i = 0; i < ncb; ++i) {
                const int cb = cb0 + i;
                const int tap = p.tap_minor ? cb % p.ntaps : cb / p.q_chunks;
                const int qc = p.tap_minor ? cb / p.ntaps : cb - tap * p.q_chunks;
                float* obase = p.out + (long long)p.tap_out[tap] * p.st + (long long)(p0 + q * 32) * p.sp + qc * 64;
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + i * 64 + half * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int qi = qc * 64 + half * 32 + lane;
                    if (qi < p.q_extent) {
                        const int nrows = min(32, p.p_extent - (p0 + q * 32));
                        for (int rr = 0; rr < nrows; ++rr)
                            atomicAdd(obase + (long long)rr * p.sp + half * 32 + lane, tile[rr * 33 + lane]);
                    }
                    __syncwarp();
                }
            }
        } else if (p.tap_minor && ncb == 8 && p.ntaps == 8 && p.sq == 8 && p.st == 1 &&
                   (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (p.sp & 3) == 0) {
            // transposed conv, G[p][q][8 taps]: this CTA holds all 8 taps of one 64-channel chunk, so the taps of one
            // (p, q) are 32 contiguous bytes: two 16-byte vector REDs instead of eight scattered 4-byte ones (the
            // scattered form made every convT weight gradient cost ~0.17 ms whatever its size: ~19 M REDs per launch)
            const int qc = cb0 >> 3;
            for (int c8 = 0; c8 < 8; ++c8) {
                uint32_t v[8][8];
#pragma unroll
                for (int t = 0; t < 8; ++t) tmem_ld8(t_addr + t * 64 + c8 * 8, v[t]);
                tmem_ld_wait();
                if (pidx < p.p_extent) {
                    float* dst = p.out + (long long)pidx * p.sp + (long long)(qc * 64 + c8 * 8) * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (qc * 64 + c8 * 8 + j < p.q_extent) {
                            atomicAdd(reinterpret_cast<float4*>(dst + j * 8),
                                      make_float4(__uint_as_float(v[0][j]), __uint_as_float(v[1][j]),
                                                  __uint_as_float(v[2][j]), __uint_as_float(v[3][j])));
                            atomicAdd(reinterpret_cast<float4*>(dst + j * 8 + 4),
                                      make_float4(__uint_as_float(v[4][j]), __uint_as_float(v[5][j]),
                                                  __uint_as_float(v[6][j]), __uint_as_float(v[7][j])));
                        }
                    }
                }
            }
        } else {
            for (int i = 0; i < ncb; ++i) {
                const int cb = cb0 + i;
                const int tap = p.tap_minor ? cb % p.ntaps : cb / p.q_chunks;
                const int qc = p.tap_minor ? cb / p.ntaps : cb - tap * p.q_chunks;
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[16];
                    tmem_ld16(t_addr + i * 64 + c * 16, v);
                    tmem_ld_wait();
                    if (pidx < p.p_extent) {
                        float* dst = p.out + (long long)p.tap_out[tap] * p.st + (long long)pidx * p.sp;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int qi = qc * 64 + c * 16 + j;
                            if (qi < p.q_extent) atomicAdd(dst + (long long)qi * p.sq, __uint_as_float(v[j]));
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

extern "C" __global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
    wgrad_body<0>(p);
}
extern "C" __global__ void __launch_bounds__(kIm2colThreads, 1)
    wgrad_im2col5_kernel(const __grid_constant__ WgradParams p) {
    wgrad_body<5>(p);
}

}  // namespace b200
