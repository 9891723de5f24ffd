// Parameter blocks shared between the host-side launchers (api.cu) and the tcgen05 kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200 {

constexpr int kMaxTaps = 27;
constexpr int kMaxMaps = 8;
constexpr int kBrickRows = 128;          // output voxels per M tile (one TMA box of TW x TH x TD)
constexpr int kBoxBytes = 128 * 128;     // one TMA box: 128 rows x 64 bf16
constexpr int kMaxStatCols = 1024;       // widest Cout whose BatchNorm partial sums fit the CTA's smem
constexpr int kThreads = 256;            // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4..7 epilogue

// Ablation switches (loads without MMAs, MMAs without loads, ...) exist only in the development library
// (-DB200_DEV, build.py dev=True): in the product library B200_ABLATE(p) is the constant 0 and every branch on it
// is removed by the compiler.
#ifdef B200_DEV
#define B200_ABLATE(p) ((p).ablate)
#else
#define B200_ABLATE(p) 0
#endif

enum EpilogueMode : int {
    EPI_PLAIN = 0,       // out = acc                         (dgrad)
    EPI_BIAS_STATS = 1,  // out = bf16(acc + bias); per-tile sum / sum-of-squares of out (train fprop)
    EPI_AFFINE_RELU = 2, // out = relu(acc * scale + shift)   (eval fprop, BatchNorm folded)
    EPI_BIAS = 3,        // out = acc + bias                  (transposed conv forward)
    EPI_SPLITK = 4,      // split-K partial: fp32 acc stored to this split's workspace slice; finalised by a second pass
};

// Implicit GEMM  D[m, n] = sum_{tap, c} A_tap[m, c] * B[tap][n][c]
//   m : output voxel (row of a TW x TH x TD brick, w fastest), n : output column, c : input channel.
// A_tap is read by TMA from an NDHWC bf16 tensor through a_map[a_map_of_tap[tap]] at the brick origin
// shifted by (tap_dw, tap_dh, tap_dd); out-of-range coordinates are zero-filled by the TMA unit, which is
// the convolution's zero padding.  B is the packed weight [taps][n][c] (c contiguous).
// h-halo mode (group == 3): bricks are tw x th x 1 (w fastest), the A box holds rows h0-1 .. h0+th of the brick, and
// the three kh taps of a (kd, kw) group are the same box read at row offsets 0, tw, 2*tw (multiples of the 8-row
// swizzle atom, so the plain SWIZZLE_128B descriptor stays valid): A traffic per tap drops ~2.7x.
struct alignas(64) IgemmParams {
    CUtensorMap a_map[kMaxMaps];
    CUtensorMap b_map;
    CUtensorMap c_map[kMaxMaps];  // epilogue v2: TMA store maps of the output (one per output offset group)
    int epi_v2;         // 1: stage the bf16 tile in shared memory, TMA store, statistics from the staged tile
    int c_bufs;         // staging tiles (epilogue v2): 2 = tile i+1 is staged while the TMA store of tile i still reads
    int pair;           // 1: launched as clusters of two CTAs; M = 256 tcgen05.mma.cta_group::2, B tile split between them
    int ablate;         // development library only (B200_ABLATE): 1 = barriers armed without TMA loads, 2 = no MMAs issued
    int ntaps;
    int a_map_of_tap[kMaxTaps];
    int tap_dw[kMaxTaps], tap_dh[kMaxTaps], tap_dd[kMaxTaps];
    int cin;        // channels per tap (K extent of one tap)
    int kc_blocks;  // ceil(cin / 64)
    int block_n;    // UMMA N (multiple of 16, <= 256)
    int n_tiles;    // ceil(ncols / block_n)
    int ncols;      // valid output columns
    int nbw, nbh, nbd, nbatch;  // bricks per axis, batch
    int tw_log2, th_log2, td_log2;
    int W, H, D;    // extent of the M grid
    int stages;
    int group;          // taps per pipeline stage: 1, or 3 (h-halo mode: one A box of th+2 rows serves kh = 0,1,2)
    int a_stage_bytes;  // bytes of one A box (16 KB, or (th+2)*tw*128 B in h-halo mode)
    int a_goff[3];      // start offset (bytes >> 4) of tap g of a group inside the A box
    int b_mn;           // 1: B is read MN-major from the fprop-packed weights [tap][K rows][N contiguous] (dgrad):
                        //    per stage ceil(block_n/64) atoms of [group taps][64 K rows][64 N] (8 KB per tap)
    int mode;
    const float* vec0;  // bias (modes 1,3) or scale (mode 2)
    const float* vec1;  // shift (mode 2)
    float* stats;       // mode 1: [gridDim.x][ncols][2]  (sum, sum of squares) per CTA
    __nv_bfloat16* out;
    long long out_sn, out_sd, out_sh, out_sw;  // element strides of the output tensor
    int out_mul;        // output coordinate = m coordinate * out_mul + offset[group]
    int cols_per_group; // columns sharing one output offset (transposed conv: Cout per tap)
    int out_od[kMaxMaps], out_oh[kMaxMaps], out_ow[kMaxMaps];
    // split-K over the taps (deep levels: a handful of M tiles, K = 27 * Cin up to 27648): every (M, N) tile is computed
    // by `splits` work units that each take a contiguous range of the tap groups and store their fp32 partial tile to
    // ws[split][voxel][ncols]; splitk_finalize_kernel adds the slices in split order and writes the bf16 output (+ bias,
    // BatchNorm partial sums)
    int splits;
    long long ws_slice_vox;   // voxels per slice
    float* ws;
    // direct first-layer form (igemm_im2col_kernel): the A operand is not loaded by TMA but built in shared memory by
    // eight producer warps straight from the fp32 (N, C, D, H, W) network input — column k = c*27 + kd*9 + kh*3 + kw of
    // row m is x[n, c, d+kd-1, h+kh-1, w+kw-1] (zero outside the volume), i.e. the rows of the im2col matrix, which
    // therefore never exists in HBM (it cost 288 B per voxel to write and again to read, against 20 B of input)
    const float* x_src;
};
constexpr int kIm2colThreads = 512;     // igemm_im2col_kernel: the 256 threads above + warps 8..15 building A

// Depth-marching implicit GEMM for 3x3x3 convolutions with 64 output columns (fprop with Cout = 64, dgrad with Cin = 64).
// A tcgen05.mma in SS mode costs max(N/2, ~42 + 0.18 N) cycles (tools/probe_mma.py): N = 64 cannot exceed 60 % of the
// tensor pipe, N = 192 reaches 99.9 %.  So the three kd taps are concatenated along N: for input slice dz the MMA
// D[128 voxels x 192] = A(dz, kh, kw) . [W(kd=a) | W(kd=b) | W(kd=c)] adds into the accumulators of the three output
// slices dz-1, dz, dz+1, which live side by side in a ring of eight 64-column TMEM slots.  A CTA owns a (w,h) brick
// column and marches along d; an output slice is complete one input slice later and is drained by the epilogue warps
// while the MMAs go on.
constexpr int kDmSlots = 8;      // TMEM ring: 8 x 64 columns
constexpr int kDmG = 2;          // input slices that share one pass over the tap weights (B stream / kDmG)
constexpr int kDmAStages = 3;    // A ring: kDmG h-halo boxes of 18 KB per stage
constexpr int kDmBStages = 4;    // B ring: [3 slabs][64 rows][128 B] = 24 KB
constexpr int kDmABytes = 18 * 8 * 128;
constexpr int kDmAStageBytes = kDmG * kDmABytes;
constexpr int kDmBBytes = 3 * 8192;
struct alignas(64) DmarchParams {
    CUtensorMap a_map;   // box (64 ch, 8 w, 18 h, 1, 1)
    CUtensorMap b_map;   // K-major: (k, rows, taps) box (64, 64, 1);  MN-major: (n, k rows, taps) box (64, 64, 1)
    CUtensorMap c_map;   // output store, box (64 ch, 8 w, 16 h, 1, 1)
    int b_mn;            // 1: dgrad (B read MN-major from the fprop-packed weights)
    int sign;            // +1 fprop, -1 dgrad (tap shifts negated)
    int cin;             // K extent per tap
    int kc_blocks;
    int ncols;           // valid output columns: 64, or 32 (computed as 64: the weight rows / columns past 32 are
                         // zero-filled by the TMA unit, the store is clipped by the output map)
    int W, H, D, nbatch, nbw, nbh;
    int seg_len, nseg;   // output slices per work unit, segments per column
    int mode;
    const float* vec0;
    const float* vec1;
    float* stats;        // [gridDim.x][ncols][2]
    int ablate;          // development library only (B200_ABLATE): 1 = no TMA loads after arming the barriers (MMAs on stale
                         // shared memory), 2 = no MMAs (loads + epilogue only); results are garbage, timings are not
};

// CTA-pair form (dmarch2.cu, tcgen05.mma.cta_group::2): each CTA stages HALF of the B tile (96 of the 192 rows, three
// 32-row boxes — b_map box (64, 32, 1), K-major only); six logical accumulator slots + two mirror slots
constexpr int kDm2BStages = 6;
constexpr int kDm2BBytes = 96 * 128;
constexpr int kDm2Slots = 6;
constexpr int kDm2Smem = 1024 + kDmAStages * kDmAStageBytes + kDm2BStages * kDm2BBytes + kBoxBytes +
                         8 * (2 * kDmAStages + 2 * kDm2BStages + 2 * kDm2Slots) + 64 + (4 * 64 * 2 + 128 + 128) * 4;

// Weight-gradient GEMM  G[tap][p][q] += sum_{voxel} P[voxel][p] * Q_tap[voxel][q]
// Both operands are voxel-major in memory (channel contiguous), i.e. MN-major UMMA operands.
struct alignas(64) WgradParams {
    CUtensorMap p_map;
    CUtensorMap q_map[kMaxMaps];
    int ntaps;
    int q_map_of_tap[kMaxTaps];
    int tap_dw[kMaxTaps], tap_dh[kMaxTaps], tap_dd[kMaxTaps];
    int tap_out[kMaxTaps];  // index of tap t along the output's tap axis (native or packed tap order)
    int p_extent, q_extent;
    int q_chunks;      // ceil(q_extent / 64)
    int n_colblocks;   // ntaps * q_chunks
    int tap_minor;     // column block = chunk * ntaps + tap (else tap * q_chunks + chunk): with 8 taps a CTA then owns
                       // every tap of one 64-channel chunk, which the transposed-conv epilogue stores as 32-byte runs
    int cb_per_group;  // column blocks (64 wide) accumulated by one CTA (<= 8 -> 512 TMEM columns)
    int n_groups, p_tiles, splits;
    int nbw, nbh, nbd, nbatch;
    int tw, th, td;
    float* out;
    long long st, sp, sq;  // element strides of G for (tap, p, q)
    // first-layer form (wgrad_im2col5_kernel): Q = im2col rows built in shared memory from the fp32 network input
    const float* x_src;
    int tw_log2, th_log2, W, H, D;
};

// h-halo variant of the weight-gradient GEMM (wgrad_halo.cu): bricks are 8 w x 16 h x 1 d; the shifted operand Q is
// loaded as one 18-row halo box per (kd, kw, 64-channel chunk) and its three kh taps are three N-atoms of ONE MN-major
// descriptor (LBO = 1 KB = one h line), i.e. one N = 192 MMA per k-step instead of three boxes / 64 columns each.
constexpr int kWhQStages = 5;
constexpr int kWhPBoxes = 6;     // P ring capacity in 16 KB boxes (wgrad_halo.cu)
constexpr int kWhQBytes = 18 * 8 * 128;
struct alignas(64) WgradHaloParams {
    CUtensorMap p_map;   // box (64, 8, 16, 1, 1)
    CUtensorMap q_map;   // box (64, 8, 18, 1, 1)
    int sgn;             // +1: Q taps shifted by +off(t); -1: swapped roles, shifted by -off(t)
    int tap_out[kMaxTaps];
    int p_extent, q_extent, q_chunks;
    int n_units;         // 9 * q_chunks column units of 192 TMEM columns: unit = (kd*3 + kw) * q_chunks + chunk
    int pair;            // depth-pair mode for P tiles of <= 64 channels: the M side is [P(brick); P(brick + 1 slice)],
                         // so one MMA yields two kd taps.  6 * q_chunks units: unit = ((kw * q_chunks + chunk) * 2 + kind,
                         // kind 0: Q one slice ahead (rows 0-63 -> kd 2, rows 64-127 -> kd 1); kind 1: Q one slice
                         // behind (rows 0-63 -> kd 0, upper rows unused).  Bricks then start at depth -1 (nbd = D + 1).
    int units_per_group; // <= 2 (384 of 512 TMEM columns)
    int n_groups, p_tiles, splits;
    int last_splits;     // voxel splits of the last group (fewer when it holds a single unit: equal work per CTA)
    int nbw, nbh, nbd, nbatch;
    float* out;
    long long st, sp, sq;
};

// Depth-marching forward of the 5-modality first layer (conv1_march.cu): slice images of 48 rows (45 real: k = c*9 +
// kh*3 + kw) x 128 voxels, MN-major, two 64-voxel halves kC1HalfBytes apart
constexpr int kC1EpiWGs = 2;                         // epilogue warpgroups (warps 4 ..); a third one changes nothing (0.217 vs 0.212 ms)
constexpr int kC1BuilderWarp0 = 4 + 4 * kC1EpiWGs;   // first of the four builder warps
constexpr int kC1Threads = 32 * (kC1BuilderWarp0 + 4);
constexpr int kC1Cin = 5;
constexpr int kC1Rows = 48;
constexpr int kC1HalfBytes = kC1Rows * 128;
constexpr int kC1ImgBytes = 2 * kC1HalfBytes;
constexpr int kC1Imgs = 6;       // slice-image ring
constexpr int kC1Slots = 8;      // TMEM ring: 64-column accumulators (one output slice each)
constexpr int kC1Smem = 1024 + kC1Imgs * kC1ImgBytes + 3 * 8192 + kC1EpiWGs * kBoxBytes +
                        8 * (2 * kC1Imgs + 2 * kC1Slots + 1) + 64 + (4 * kC1EpiWGs * 64 * 2 + 2 * 64) * 4;
struct alignas(64) Conv1MarchParams {
    CUtensorMap b_map;   // packed weights [3 kd][Cout][64] (k = c*9 + kh*3 + kw, 45 real), box (64, 64, 1)
    CUtensorMap c_map;   // output store, box (64 ch, 8 w, 16 h, 1, 1)
    const float* x;      // (N, 5, D, H, W) fp32
    const float* vec0;
    const float* vec1;
    float* stats;        // [gridDim.x][ncols][2]
    int ncols;           // valid output columns (<= 64; computed as 64 against zero-filled weight rows)
    int W, H, D, nbatch, nbw, nbh;
    int seg_len, nseg;
    int mode;
    int ablate;          // development library only (B200_ABLATE): 1 = no output stores, 2 = builders skip the input loads, 3 = no MMAs
};

// ... and its weight gradient (conv1_march_wgrad_kernel): image ring laid out [half][slot][48 rows], dy bricks by TMA
constexpr int kC1WgGroups = 4;    // builder groups of two warps (an image takes ~1.5 us from loads to arrive)
constexpr int kC1WgImgs = 8;      // slice-image ring
constexpr int kC1WgThreads = 256 + 64 * kC1WgGroups;
constexpr int kC1WgPSlots = 5;    // dy bricks in flight (16 KB each) + one shared zero box
constexpr int kC1WgSmem = 1024 + 2 * kC1WgImgs * kC1HalfBytes + (kC1WgPSlots + 1) * kBoxBytes +
                          8 * (2 * kC1WgImgs + 2 * kC1WgPSlots + 1) + 64;
struct alignas(64) Conv1MarchWgradParams {
    CUtensorMap p_map;   // dy, box (64 ch, 8 w, 16 h, 1, 1)
    const float* x;      // (N, 5, D, H, W) fp32
    float* dw;           // fp32 [Cout][5][3][3][3] (+=)
    int ncols;           // Cout <= 64
    int W, H, D, nbatch, nbw, nbh;
    int seg_len, nseg;
    int ablate;          // development library only: 1 = dy slots armed without loads, 2 = builders skip the input loads, 3 = no MMAs
};

}  // namespace b200
