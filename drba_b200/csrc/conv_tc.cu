// tcgen05 implicit-GEMM convolution engine for sm_100a (fp16 operands, fp32 accumulation in TMEM).
//
// Serves every conv of IFNet 4.26-heavy (models/rife_426_heavy/IFNet_HDv3.py:28-96: Head, conv0
// (two 3x3 stride-2), 8 x ResConv (3x3), lastconv ConvTranspose2d(c, 52, 4, 2, 1) + PixelShuffle(2))
// in the reference's own GPU precision (fp16 under torch.autocast, models/rife.py:26,78).
//
// GEMM view: M = output pixels, N = output channels, K = taps x Cin.
//  * activations are NHWC fp16; an M tile is a 128-pixel rectangle (16x8, 32x4, ...), the N tile is
//    <= 128 channels; the accumulator is a [128 lanes x ntile columns] fp32 block of tensor memory;
//  * per K step ONE tiled TMA load brings the shifted input window into shared memory as K-major,
//    hardware-swizzled [128 x Kc] A tiles -- zero padding is TMA out-of-bounds fill, so there is no
//    im2col buffer.  Stride-2 convs view the input as [H/2][2][W/2][2][C] (rank-5 tensor map) so a
//    tap stays a dense box;
//  * HALO MODES (3x3 layers whose weights are resident): the TMA unit moves ~one box line per 3-5
//    cycles whatever the line length, so instead of one box per tap the tile's whole neighbourhood is
//    loaded once per K chunk and the nine taps are UMMA descriptor start offsets into it (the swizzle
//    phase comes from the absolute shared-memory address, so any start offset and any stride between
//    8-row groups works): stride 1 = one (16 MT + 2) x 16 pixel box for 8-pixel-wide tiles; thin
//    inputs (16 / 32 channels) PACKED 4 / 2 pixels per 128-byte line; stride 2 = the four parity
//    planes of the 17 x 9 cell neighbourhood as four boxes of one stage;
//  * thin layers (Kc = 16 / 32 channels per K step) work on SUPER TILES of 4 / 2 vertically stacked
//    M tiles so that every TMA box still carries 16 KB (measured: a TMA instruction costs its issuing
//    thread ~130 cycles regardless of size, which is what bounds layers with 4 KB boxes);
//  * WEIGHTS: when all weight tiles of a layer fit in 96 KB they are loaded ONCE per layer per CTA
//    and stay resident (every IFNet layer at the fine levels; the ring then carries activations only);
//    otherwise they stream through the ring next to the activations, issued by a second producer warp;
//  * ConvTranspose(4,2,1) is four 2x2-tap phase convs; the lastconv epilogue writes the
//    PixelShuffle(2)'d NHWC fp32 [4h][4w][16] directly.
//
// Execution model: ONE PERSISTENT LAUNCH RUNS A PROGRAM OF LAYERS.  Measured on B200, a 128-pixel
// tile costs more in CTA start-up (barrier init, TMEM allocation, descriptor fetch, drain) than in
// MMA or TMA time, and the coarse IFNet levels (510 ... 8160 pixels) are pure launch latency.  So:
//  * grid = min(tiles, 148) CTAs, one per SM; each CTA walks its tiles of the current layer with a
//    shared-memory ring that never drains between tiles;
//  * warp 0 = activation (A) producer, warp 10 = weight (B) producer, warp 1 = tcgen05.mma issuer --
//    each runs its loop warp-converged with ONE ELECTED lane issuing, which keeps TMA / MMA operands in
//    uniform registers; warps 2-5 / 6-9 = two epilogue groups that drain EVERY tile together (one M
//    tile of the super tile each, or half of the columns) from a double-buffered TMEM accumulator
//    (the epilogue of tile i overlaps the main loop of tile i+1).  The plain epilogue exists in two
//    template variants: one lane per pixel row, or rows moved four lanes per pixel through a
//    swizzled per-warp staging tile (coalesced residual / result traffic for wide rows);
//  * measured in isolation (scripts/mma_shapes.cu, profiles/r2_mma_shapes.json) a tcgen05.mma with M = 128, K = 16 costs
//    max((32 M + 32 N) / 128, M N / 256) cycles: 40 / 48 / 64 for N = 32 / 64 / 128 -- the operand bytes over 128 B/clk
//    of shared-memory read bandwidth, or the math floor.  What bounds the layers in practice is elsewhere: an SM pulls
//    ~29 B per clock from L2 whatever the box shape (block4.conv0a: 78 KB per tile in 2700 cycles; block0.res: 20 KB
//    per K iteration in 700), and every layer boundary (drain, fence, grid barrier, refill) costs ~15k cycles;
//  * layers in the latency regime whose weights do not fit the resident region take a narrower N tile and keep ONE
//    N split resident per CTA (tile index with the split fastest, LayerDev::wsplit): they stay in halo mode and read
//    their input once per K chunk instead of once per tap (block0.res: 540 KB -> 218 KB per CTA, 12.4 -> 6.9 us);
//  * consecutive layers (a whole IFBlock: conv0a, conv0b, 8 x ResConv, lastconv) are chained inside
//    the launch with a grid-wide barrier (release/acquire counter in global memory) instead of a
//    kernel boundary; up to two independent images (the two interpolated frames of a DRBA window)
//    share the launch and the weights.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace drba {

constexpr int kTcThreads = 352;        // warp 0 A producer, warp 1 MMA, warps 2..9 epilogue, warp 10 B producer
constexpr int kTileM = 128;
constexpr int kStages = 6;
constexpr int kABytesMax = 16384;      // one A stage: MT x 128 rows x Kc x 2 B
constexpr int kStageBytes = 32768;     // streaming mode: A (<= 16 KB) | B (<= 16 KB)
constexpr int kBRegion = 98304;        // resident mode: all weight tiles of the layer, then 6 A stages
constexpr int kRingBytes = kStages * kStageBytes + 8192;   // whole ring; the extra 8 KB give block3.conv0a (36 KB of weights + 80 KB
                                                           // stride-2 halo stages) a second stage: single-buffered it ran at 6600 cycles per tile
constexpr int kMaxNTile = 128;
constexpr int kMaxTapsTc = 10;          // 3x3 + the identity tap of a ResConv (residual added by the tensor core)
constexpr int kMaxGroups = 4;
constexpr int kAccBufs = 4;            // accumulator ring in tensor memory
constexpr int kTmemCols = kAccBufs * kMaxNTile;   // four accumulators of <= 128 columns = all 512 columns (one CTA per SM)
constexpr int kMaxLayers = DRBA_CONV_MAX_LAYERS;
constexpr int kMaxImages = 2;

struct alignas(64) LayerDev {
    CUtensorMap ta[kMaxImages];
    CUtensorMap tb;
    const float* bias;      // [G][cout_pad]
    const float* slope;     // PReLU slopes [cout_pad] (act 2) or NULL
    const __half* res[kMaxImages];
    const __half* res2[kMaxImages];   // second residual input (GridNet three-way sums)
    void* out[kMaxImages];
    void* out1[kMaxImages];           // optional extra outputs of the same values with their own activation
    void* out2[kMaxImages];
    float slope0, slope1, slope2;     // scalar PReLU slopes of out / out1 / out2 (act 4)
    int act1, act2;
    int OH, OW;
    int Kc, kchunks;        // channels per K step, Cin / Kc
    int S, T, G;
    int MT;                 // M tiles per super tile (stacked vertically)
    int KI;                 // pipeline iterations per super tile = T * kchunks
    int resident;           // 1: weights resident in smem for the whole layer
    int bgemm;              // 1: batched GEMM: the weight rows of a tile are those of batch element oy0 (H = batch)
    int has_bias;
    int halo;               // 1: 3x3 stride-1 layer whose taps are descriptor offsets into ONE halo tile per K chunk
    int nst;                // ring stages used by this layer
    int a_base, a_stride;   // byte offset of the activation ring and bytes per stage
    int pack;               // packed halo mode: pixels per 128-byte line (2 or 4), else 1
    int staged;             // epilogue moves residual / result through the per-warp staging tile
    int alt;                // epilogue groups take alternate tiles (whole accumulator each) instead of splitting every tile
    int wsplit;             // resident region holds the weights of ONE N split only: tile index has the split fastest, and
                            // every tile of a CTA belongs to split blockIdx.x % nsplits (host checks grid % nsplits == 0 or <= 1 tile per CTA)
    int ntile, nsplits, cout_pad, cout;
    int swz_bytes;          // 32 / 64 / 128
    int b_sub;              // bytes of one weight tile (padded to the swizzle period)
    int epilogue;           // 0: NHWC fp16 (+res, act); 1: lastconv pixel shuffle -> fp32 [4*OH][4*OW][16]
    int act;                // 0 none, 1 LeakyReLU(0.2), 2 PReLU(slope), 3 ReLU
    int out_cstride, os;
    int tiles_x, mtiles, tile_w, tile_h;   // super-tile grid; tile_w x tile_h = one 128-pixel M tile
    int total_tiles;        // nimg * G * nsplits * mtiles
    // per (group, tap): cell offset y | cell offset x << 8 | parity y << 16 | parity x << 17 (offsets biased by +64)
    int tapc[kMaxGroups][kMaxTapsTc + 3];
};

struct Program {
    int nlayers, nimg;
    unsigned* sync;         // [2] grid-barrier arrival counter, exit counter (zero between launches)
    unsigned* err;          // sticky error word in mapped pinned host memory (bit 0: grid barrier timed out)
    int dbg;
    int pad_;
    long long* trace;       // debug: clock64 stamps of CTA 0 (NULL in production)
    LayerDev L[kMaxLayers];
};

// MMA issue loops, K steps unrolled.  Descriptor start addresses are (addr >> 4) in the low 14 bits: advancing
// 16 fp16 = 32 bytes along K inside the swizzle span is +2, and shared memory (< 256 KB) never carries out of
// the field, so offsets are plain 64-bit adds.
template <int KS>
__device__ __forceinline__ void issue_tile(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t a_mt16, int MT, int ntile,
                                           uint32_t idesc, uint32_t accumulate)
{
    for (int mt = 0; mt < MT; ++mt) {
        const uint64_t da = da0 + (uint64_t)(mt * a_mt16);
#pragma unroll
        for (int k = 0; k < KS; ++k)
            tc_mma_f16(tmem_d + mt * ntile, da + (uint64_t)(k * 2), db0 + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)k);
    }
}

// halo mode (canonical 3x3 taps, checked on the host): tap (ty, tx) of M tile mt starts at halo pixel
// ((mt * 16 + ty) * 16 + tx); the tap's weight tile is b_tap16 further on
// (tap 9, present when ntaps == 10, is the IDENTITY tap of a ResConv: the centre pixel against a unit weight tile, so the
// residual is accumulated exactly in fp32 by the tensor core and the epilogue has no residual to fetch)
template <int KS>
__device__ __forceinline__ void issue_halo(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t rowb16, uint32_t b_tap16,
                                           int MT, int ntile, uint32_t idesc, uint32_t accumulate, int ntaps)
{
    for (int mt = 0; mt < MT; ++mt) {
        const uint64_t da_mt = da0 + (uint64_t)((uint32_t)(mt * 256) * rowb16);
        const uint32_t d = tmem_d + mt * ntile;
#pragma unroll
        for (int tap = 0; tap < 10; ++tap) {
            if (tap == 9 && ntaps < 10) break;
            const int ty = tap == 9 ? 1 : tap / 3, tx = tap == 9 ? 1 : tap % 3;
            const uint64_t da = da_mt + (uint64_t)((uint32_t)(ty * 16 + tx) * rowb16);
            const uint64_t db = db0 + (uint64_t)((uint32_t)tap * b_tap16);
#pragma unroll
            for (int k = 0; k < KS; ++k)
                tc_mma_f16(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)(tap | k));
        }
    }
}

// stride-2 halo mode: the input is viewed as [H/2][2][W/2][2][C]; the four (row, column) parity planes of the
// 17 x 9 cell neighbourhood of an 8 x 16 output tile are four boxes of one stage (sub16 apart).  Tap (ky, kx) reads
// parity (ky != 1, kx != 1) at cell offset (ky == 0 ? -1 : 0, kx == 0 ? -1 : 0), i.e. box row / column
// (ky != 0, kx != 0) since every box starts one cell up and left.  Rows of 8 cells are contiguous, image rows are
// 9 cells apart (SBO).
template <int KS>
__device__ __forceinline__ void issue_halo_s2(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t rowb16, uint32_t sub16,
                                              uint32_t b_tap16, uint32_t idesc, uint32_t accumulate)
{
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        const int par = (ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0);
        const int offr = ky != 0 ? 1 : 0, offc = kx != 0 ? 1 : 0;
        const uint64_t da = da0 + (uint64_t)((uint32_t)par * sub16 + (uint32_t)(offr * 9 + offc) * rowb16);
        const uint64_t db = db0 + (uint64_t)((uint32_t)tap * b_tap16);
#pragma unroll
        for (int k = 0; k < KS; ++k)
            tc_mma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)(tap | k));
    }
}

// packed halo mode (Cin = 32 or 16: P = 2 or 4 pixels share one 128-byte line; the halo box is 10 lines x 18 rows
// for an 8P x 16 pixel tile).  M tile p holds the output pixels x = x0 + P * j + p (j = 0..7, 16 rows): for tap
// (ty, tx) its operand rows are the input pixels x0 - P + P * j + (p + tx - 1 + P), i.e. line j + q / P of halo row
// iy + ty at byte q % P * Kc * 2 inside the line -- a K offset inside the 128-byte swizzle span, like the k * 2
// advance.  Eight consecutive lines are one core group; the next image row is 10 lines further (SBO = 1280 B).
template <int P>
__device__ __forceinline__ void issue_halo_packed(uint32_t tmem_d, uint64_t da0, uint64_t db0, uint32_t b_tap16, int ntile,
                                                  uint32_t idesc, uint32_t accumulate, int ntaps)
{
    constexpr int KS = 4 / P;            // Kc = 64 / P channels = KS steps of 16
    constexpr int SUB16 = 8 / P;         // one pixel inside the line, in 16-byte units
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const uint32_t d = tmem_d + p * ntile;
#pragma unroll
        for (int tap = 0; tap < 10; ++tap) {
            if (tap == 9 && ntaps < 10) break;
            const int ty = tap == 9 ? 1 : tap / 3, tx = tap == 9 ? 1 : tap % 3;
            const int q = p + tx - 1 + P;
            const uint64_t da = da0 + (uint64_t)((ty * 10 + q / P) * 8 + (q % P) * SUB16);
            const uint64_t db = db0 + (uint64_t)((uint32_t)tap * b_tap16);
#pragma unroll
            for (int k = 0; k < KS; ++k)
                tc_mma_f16(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)(tap | k));
        }
    }
}

// exact GELU (nn.GELU default); out of line: erff inlined 16x per site bloats every instantiation
__device__ __noinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// epilogue staging tile of one warp: 32 pixel rows x 64 bytes; piece k (16 bytes) of row r sits at k ^ (r / 2 % 4),
// conflict-free both for "lane = row" and for "four lanes per row" accesses
constexpr int kStagingBytes = 2048;
__device__ __forceinline__ int stg_slot(int r, int k) { return r * 4 + (k ^ ((r >> 1) & 3)); }

// grid-wide barrier between two layers of a program: every CTA arrives once per layer
// A CTA that never arrives (the grid was not co-resident: MPS / MIG / a foreign kernel holding SMs) would hang the
// GPU, so the wait is bounded -- but a time-out is an ERROR, not a fall-through: it sets the sticky word in mapped
// host memory that drba_conv_tc_program_f16 / drba_conv_tc_status return as DRBA_E_BARRIER from then on.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target, unsigned* err) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    unsigned spins = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target && ++spins < (1u << 24));
    if (v < target && err) {
        atomicOr_system(err, 1u);
        __threadfence_system();
    }
}


// clock64 pipeline trace of CTA 0 (scripts/trace_conv.py): compiled in only with -DDRBA_TC_TRACE=1 -- even the
// never-taken branches cost the producer / MMA warps 10-15 % on TMA-issue-bound layers
#if DRBA_TC_TRACE
#define TC_TRACE(slot) do { if (trace && blockIdx.x == 0) trace[(slot)] = clock64(); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#endif

// FULL = false: the IFNet / GMFlow feature set (one output, one residual, act none / LeakyReLU / ReLU / GELU);
// FULL = true adds the GMFSS epilogue (second residual, up to three outputs, PReLU variants) at a higher register cost
// STAGED selects the plain epilogue's data path (one instantiation carries only one of them: the kernel's code size
// is felt by the issue-bound producer / MMA warps): true = through the per-warp staging tile (wide rows, many tiles
// per SM), false = one lane per pixel row (latency-bound streaming layers)
template <bool FULL, bool STAGED>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ Program prog)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], acc_full[kAccBufs], acc_empty[kAccBufs], b_full;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[512];
    __shared__ __align__(16) float s_slope[512];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* const trace = prog.trace;
    const int dbg = prog.dbg;
    const int nlayers = prog.nlayers;

    // bias / PReLU slopes of the current layer live in shared memory (the epilogue reads them per tile)
    // Executed by warps 1.. only: thread 0 goes straight from the layer boundary into the grid barrier (its share of
    // the table loads used to sit on every CTA's critical path after the barrier).  The same threads touch the next
    // layer's descriptor in the constant bank (kernel parameters), so that the ~25 scalar reads at the top of the
    // layer loop hit the constant cache instead of missing one after the other behind the barrier.
    auto stage_tables = [&](int li) {
        const LayerDev& L = prog.L[li];
        const int nb = L.G * L.cout_pad;
        const int t = (int)threadIdx.x - 32;
        if (t < 0) return;
        if (L.has_bias)
            for (int i = t; i < nb; i += kTcThreads - 32) s_bias[i] = L.bias[i];
        if (L.act == 2)
            for (int i = t; i < L.cout_pad; i += kTcThreads - 32) s_slope[i] = L.slope[i];
        if (t < (int)(sizeof(LayerDev) / 32)) {
            const int v = reinterpret_cast<const int*>(&L)[t * 8];
            asm volatile("" ::"r"(v));                       // keeps the (constant-bank) read alive
        }
    };
    uint32_t bprod_uses = 0;         // weight producer warp: resident layers issued so far
    // resident mode: every weight tile of the layer, once (B producer warp, warp-converged)
    auto load_resident_weights = [&](int li) {
        const LayerDev& L = prog.L[li];
        if (!L.resident) return;
        // the previous resident layer's weight loads must have landed (nobody else waits for them when this
        // CTA had no tile in that layer) before the region and the barrier are reused
        if (bprod_uses > 0) mbar_wait(&b_full, (bprod_uses - 1u) & 1u);
        ++bprod_uses;
        if (elect_one()) {
            const int ntiles_b = L.G * L.nsplits * L.T * L.kchunks;
            const int b_bytes = L.ntile * L.Kc * 2, b_sub = L.b_sub, Kc = L.Kc, kchunks = L.kchunks, T = L.T;
            const int nsplits = L.nsplits, ntile = L.ntile, cout_pad = L.cout_pad;
            if (L.wsplit) {
                // only this CTA's N split: tile (g, tap, kc) at index (g * T + tap) * kchunks + kc
                const int nsplit = (int)blockIdx.x % nsplits;
                mbar_expect_tx(&b_full, (uint32_t)(L.G * T * kchunks * b_bytes));
                int i = 0;
                for (int g = 0; g < L.G; ++g)
                    for (int tap = 0; tap < T; ++tap)
                        for (int kc = 0; kc < kchunks; ++kc, ++i)
                            tma_load_2d(smem + (uint32_t)(i * b_sub), &L.tb, &b_full, kc * Kc, (g * T + tap) * cout_pad + nsplit * ntile);
            } else {
            mbar_expect_tx(&b_full, (uint32_t)(ntiles_b * b_bytes));
            int i = 0;
            for (int gn = 0; gn < L.G * nsplits; ++gn) {
                const int g = gn / nsplits, nsplit = gn - g * nsplits;
                for (int tap = 0; tap < T; ++tap)
                    for (int kc = 0; kc < kchunks; ++kc, ++i)
                        tma_load_2d(smem + (uint32_t)(i * b_sub), &L.tb, &b_full, kc * Kc, (g * T + tap) * cout_pad + nsplit * ntile);
            }
            }
        }
        __syncwarp();
    };

    // let the next kernel in the stream start its prologue while this one runs (PDL)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) TC_TRACE(4090);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[0].ta[0]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[0].tb) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < kAccBufs; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
        mbar_init(&b_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    stage_tables(0);           // weights / bias are never written by a preceding kernel: safe before griddepcontrol.wait
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 10) load_resident_weights(0);
    // everything above overlapped the previous kernel's tail; its results are visible after this
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) TC_TRACE(4091);

    // ring bookkeeping (both producers and the MMA warp keep identical copies): every layer starts at stage 0 and
    // walks its own number of stages; bit s of `pbits` is the parity of the next `full` phase of stage s
    uint32_t stage = 0, pbits = 0;
    uint32_t tcount = 0;             // tiles this CTA has started so far (MMA warp and epilogue warps keep their own copy)
    uint32_t bres_phase = 0;         // MMA warp: parity of the next resident-weights barrier phase

    for (int li = 0; li < nlayers; ++li) {
        if (threadIdx.x == 0) TC_TRACE(li * 256 + 208);
        const LayerDev& L = prog.L[li];
        // layer constants -> registers (once per layer)
        const int total_tiles = L.total_tiles, mtiles = L.mtiles, nsplits = L.nsplits, G = L.G;
        const int tiles_x = L.tiles_x, tile_w = L.tile_w, tile_h = L.tile_h, MT = L.MT;
        const int KI = L.KI, T = L.T, ntile = L.ntile, kchunks = L.kchunks, Kc = L.Kc, resident = L.resident;
        const uint32_t a_base = smem + (uint32_t)L.a_base;
        const uint32_t a_stride = (uint32_t)L.a_stride;
        const int nst = L.nst, halo = L.halo, pack = L.pack, wsplit = L.wsplit;
        const int super_h = pack > 1 ? tile_h : tile_h * MT;     // output rows of one super tile
        stage = 0;

        if (warp == 0) {
            if (lane == 0) TC_TRACE(li * 256 + 205);
            // ===== activation producer: the whole warp walks the loop, one elected lane issues =====
            const uint32_t a_bytes = (dbg & 2) ? 0u : (uint32_t)(halo == 2 ? 4 * 17 * 9 * Kc * 2 : (pack > 1 ? 10 * 18 * 128 : (halo ? (16 * MT + 2) * 16 * Kc * 2 : MT * kTileM * Kc * 2)));
            const uint32_t s2_sub = ((uint32_t)(17 * 9 * Kc * 2) + 1023u) & ~1023u;      // stride-2 halo: one parity box
            int tl = 0;
            for (int idx = blockIdx.x; idx < total_tiles; idx += gridDim.x, ++tl) {
                int m, r;
                if (wsplit) { r = idx / nsplits; m = r % mtiles; r /= mtiles; }
                else { m = idx % mtiles; r = idx / mtiles; r /= nsplits; }
                const int g = r % G, img = r / G;
                const int ty = m / tiles_x;
                const int oy0 = ty * super_h, ox0 = (m - ty * tiles_x) * tile_w;
                const CUtensorMap* ta = &L.ta[img];
                int tap = 0, kc = 0;
                for (int it = 0; it < KI; ++it) {
                    if (tl < 10 && it < 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + it * 4 + 0);
                    mbar_wait(&empty_bar[stage], ((pbits >> stage) & 1u) ^ 1u);
                    if (tl == 0 && it == 0 && lane == 0) TC_TRACE(li * 256 + 207);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], a_bytes);
                        if (resident) mbar_arrive(&full_bar[stage]);      // stands in for the weight producer
                        if (halo == 2) {
                            if (!(dbg & 2)) {
#pragma unroll
                                for (int par = 0; par < 4; ++par)
                                    tma_load_5d(a_base + stage * a_stride + (uint32_t)par * s2_sub, ta, &full_bar[stage], it * Kc, par & 1, ox0 - 1, par >> 1, oy0 - 1);
                            }
                        } else if (halo) {
                            // one box per K chunk: the (16*MT+2) x 16 pixel neighbourhood of the 8 x 16*MT tile
                            if (!(dbg & 2)) tma_load_5d(a_base + stage * a_stride, ta, &full_bar[stage], it * Kc, 0, pack > 1 ? ox0 / pack - 1 : ox0 - 1, 0, oy0 - 1);
                        } else {
                            const int e = L.tapc[g][tap];
                            const int cy = (e & 0xff) - 64, cx = ((e >> 8) & 0xff) - 64, pary = (e >> 16) & 1, parx = (e >> 17) & 1;
                            if (!(dbg & 2)) tma_load_5d(a_base + stage * a_stride, ta, &full_bar[stage], kc * Kc, parx, ox0 + cx, pary, oy0 + cy);
                        }
                    }
                    __syncwarp();
                    if (tl < 10 && it < 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + it * 4 + 1);
                    if (++kc == kchunks) { kc = 0; ++tap; }
                    pbits ^= 1u << stage;
                    if (++stage == (uint32_t)nst) stage = 0;
                }
            }
        } else if (warp == 10) {
            // the next layer's tensor maps are prefetched HERE, not by the activation producer: there the two prefetches
            // sat in front of the first TMA of the layer (trace: 1250 cycles between the producer's entry and its first tile)
            if (li + 1 < nlayers && elect_one()) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[li + 1].ta[0]) : "memory");
                if (prog.nimg > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[li + 1].ta[1]) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[li + 1].tb) : "memory");
            }
            __syncwarp();
            // ===== weight producer (streaming mode only; resident weights were issued at the layer boundary) =====
            if (!resident) {
                const uint32_t b_bytes = (dbg & 4) ? 0u : (uint32_t)(ntile * Kc * 2);
                const int cout_pad = L.cout_pad;
                const CUtensorMap* tb = &L.tb;
                for (int idx = blockIdx.x; idx < total_tiles; idx += gridDim.x) {
                    int r = idx / mtiles;
                    const int nsplit = r % nsplits; r /= nsplits;
                    const int g = r % G;
                    int brow0 = g * T * cout_pad + nsplit * ntile;
                    if (L.bgemm) brow0 += ((idx % mtiles) / tiles_x) * tile_h * MT * cout_pad;    // batch element = output row
                    int tap = 0, kc = 0;
                    for (int it = 0; it < KI; ++it) {
                        mbar_wait(&empty_bar[stage], ((pbits >> stage) & 1u) ^ 1u);
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], b_bytes);
                            if (!(dbg & 4)) tma_load_2d(smem + stage * kStageBytes + kABytesMax, tb, &full_bar[stage], kc * Kc, brow0 + tap * cout_pad);
                        }
                        __syncwarp();
                        if (++kc == kchunks) { kc = 0; ++tap; }
                        pbits ^= 1u << stage;
                        if (++stage == (uint32_t)nst) stage = 0;
                    }
                }
            } else {
                // keep the ring position in step with the other warps
                const int mine = total_tiles > (int)blockIdx.x ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
                const uint32_t uses = (uint32_t)mine * (uint32_t)KI;     // stage s is used ceil((uses - s) / nst) times
                for (uint32_t s2 = 0; s2 < (uint32_t)nst; ++s2) {
                    const uint32_t cnt = uses > s2 ? (uses - s2 + nst - 1) / nst : 0u;
                    pbits ^= (cnt & 1u) << s2;
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer: whole warp in the loop, one elected lane issues =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, K-major both,
            // N >> 3 at [17,23), M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(ntile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const int ksteps = (dbg & 1) ? 0 : (Kc >> 4);
            const uint32_t a_mt16 = (uint32_t)(kTileM * Kc * 2) >> 4;     // one M tile inside the stage
            const uint32_t b_sub16 = (uint32_t)L.b_sub >> 4;
            const uint64_t desc_hi = make_desc(0, L.swz_bytes);
            // halo mode: rows of an 8-pixel tile row are contiguous, tile rows are 16 pixels apart in the halo tile
            const uint32_t rowb16 = (uint32_t)(Kc * 2) >> 4;
            const uint64_t desc_hi_halo = (desc_hi & ~(0x3FFFull << 32)) | ((uint64_t)(16u * rowb16) << 32);
            const uint64_t desc_hi_s2 = (desc_hi & ~(0x3FFFull << 32)) | ((uint64_t)(9u * rowb16) << 32);
            const uint32_t s2_sub16 = (((uint32_t)(17 * 9 * Kc * 2) + 1023u) & ~1023u) >> 4;
            const uint32_t b_tap16 = (uint32_t)kchunks * b_sub16;
            const uint64_t desc_a_packed = (make_desc(0, 128) & ~(0x3FFFull << 32)) | ((uint64_t)(1280u >> 4) << 32);
            const uint32_t a_base16 = (a_base & 0x3FFFFu) >> 4, a_stride16 = a_stride >> 4;
            const uint32_t b_res16 = (smem & 0x3FFFFu) >> 4;
            if (resident && total_tiles > (int)blockIdx.x) {
                mbar_wait(&b_full, bres_phase);
                tc_fence_after();
            }
            if (resident) bres_phase ^= 1u;
            int tl = 0;
            for (int idx = blockIdx.x; idx < total_tiles; idx += gridDim.x, ++tcount, ++tl) {
                const uint32_t buf = tcount & (uint32_t)(kAccBufs - 1), use = tcount / (uint32_t)kAccBufs;
                int r, nsplit;
                if (wsplit) { nsplit = idx % nsplits; r = idx / nsplits / mtiles; }
                else { r = idx / mtiles; nsplit = r % nsplits; r /= nsplits; }
                const int g = r % G;
                const uint32_t b_tile16 = b_res16 + (uint32_t)((wsplit ? g : g * nsplits + nsplit) * T * kchunks) * b_sub16;   // resident: tiles of (g, nsplit)
                mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);     // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * (uint32_t)kMaxNTile;
                uint32_t accumulate = 0;
                for (int it = 0; it < KI; ++it) {
                    mbar_wait(&full_bar[stage], (pbits >> stage) & 1u);
                    tc_fence_after();
                    if (tl < 10 && it < 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + it * 4 + 2);
                    const uint32_t sa16 = a_base16 + stage * a_stride16;
                    if (elect_one()) {
                        if (pack > 1) {
                            const uint64_t da0 = desc_a_packed | (uint64_t)sa16;
                            const uint64_t db0 = desc_hi | (uint64_t)b_tile16;
                            if (ksteps == 0) {}
                            else if (pack == 2) issue_halo_packed<2>(tmem_d, da0, db0, b_tap16, ntile, idesc, accumulate, T);
                            else issue_halo_packed<4>(tmem_d, da0, db0, b_tap16, ntile, idesc, accumulate, T);
                        } else if (halo == 2) {
                            const uint64_t da0 = desc_hi_s2 | (uint64_t)sa16;
                            const uint64_t db0 = desc_hi | (uint64_t)(b_tile16 + (uint32_t)it * b_sub16);
                            switch (ksteps) {
                                case 1: issue_halo_s2<1>(tmem_d, da0, db0, rowb16, s2_sub16, b_tap16, idesc, accumulate); break;
                                case 2: issue_halo_s2<2>(tmem_d, da0, db0, rowb16, s2_sub16, b_tap16, idesc, accumulate); break;
                                case 4: issue_halo_s2<4>(tmem_d, da0, db0, rowb16, s2_sub16, b_tap16, idesc, accumulate); break;
                                default: break;
                            }
                        } else if (halo) {
                            const uint64_t da0 = desc_hi_halo | (uint64_t)sa16;
                            const uint64_t db0 = desc_hi | (uint64_t)(b_tile16 + (uint32_t)it * b_sub16);
                            switch (ksteps) {
                                case 1: issue_halo<1>(tmem_d, da0, db0, rowb16, b_tap16, MT, ntile, idesc, accumulate, T); break;
                                case 2: issue_halo<2>(tmem_d, da0, db0, rowb16, b_tap16, MT, ntile, idesc, accumulate, T); break;
                                case 4: issue_halo<4>(tmem_d, da0, db0, rowb16, b_tap16, MT, ntile, idesc, accumulate, T); break;
                                default: break;
                            }
                        } else {
                            const uint32_t sb16 = resident ? b_tile16 + (uint32_t)it * b_sub16 : sa16 + (kABytesMax >> 4);
                            const uint64_t da0 = desc_hi | (uint64_t)sa16, db0 = desc_hi | (uint64_t)sb16;
                            switch (ksteps) {
                                case 1: issue_tile<1>(tmem_d, da0, db0, a_mt16, MT, ntile, idesc, accumulate); break;
                                case 2: issue_tile<2>(tmem_d, da0, db0, a_mt16, MT, ntile, idesc, accumulate); break;
                                case 4: issue_tile<4>(tmem_d, da0, db0, a_mt16, MT, ntile, idesc, accumulate); break;
                                default: break;
                            }
                        }
                        tc_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (tl < 10 && it < 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + it * 4 + 3);
                    pbits ^= 1u << stage;
                    if (++stage == (uint32_t)nst) stage = 0;
                }
                if (elect_one()) tc_commit(&acc_full[buf]);
                __syncwarp();
            }
        } else {
            // ===== epilogue: group 0 = warps 2..5, group 1 = warps 6..9; a warp may only touch TMEM
            // lanes 32*(warp%4) .. +31 =====
            const uint32_t group = (uint32_t)(warp - 2) >> 2;
            const int q = warp & 3;
            const int row = q * 32 + lane;
            uint4* stg = reinterpret_cast<uint4*>(smem_raw + (smem - smem_u32(smem_raw)) + kRingBytes + (warp - 2) * kStagingBytes);
            const int row_w = pack > 1 ? 8 : tile_w;
            const int ry = row / row_w, rx = row - ry * row_w;
            const int OH = L.OH, OW = L.OW, cout = L.cout, cout_pad = L.cout_pad, act = L.act;
            const int epilogue = L.epilogue, os = L.os, cstride = L.out_cstride, act1 = L.act1, act2 = L.act2;
            const float slope0 = L.slope0, slope1 = L.slope1, slope2 = L.slope2;
            // Both groups work on EVERY tile (half of the accumulator each), so a tile's drain latency is halved and
            // layers with one or two tiles per SM do not leave a group idle: M tiles split when the super tile
            // stacks several, columns split otherwise.
            // ALTERNATE mode (layers with many tiles per CTA): group g drains the WHOLE accumulator of the tiles with
            // tcount % 2 == g, so the drains of consecutive tiles overlap each other (a drain is a latency chain:
            // TMEM load -> residual -> activation -> staging -> store) and, with four accumulators, the MMA warp runs
            // up to three tiles ahead; measured tile period of block4.res before: 3700 cycles, of which MMA issue 1900.
            const bool alt = L.alt != 0;
            const int mt_lo = alt ? 0 : (MT >= 2 ? (int)group * (MT >> 1) : 0), mt_hi = alt ? MT : (MT >= 2 ? mt_lo + (MT >> 1) : 1);
            const int chalf = ((ntile >> 1) + 15) & ~15;
            const int c_lo = (alt || MT >= 2) ? 0 : (group ? chalf : 0), c_hi = (alt || MT >= 2) ? ntile : (group ? ntile : chalf);
            int tl = 0;
            for (int idx = blockIdx.x; idx < total_tiles; idx += gridDim.x, ++tcount, ++tl) {
                const uint32_t buf = tcount & (uint32_t)(kAccBufs - 1), use = tcount / (uint32_t)kAccBufs;
                // alternate mode: this tile belongs to the other group (which signs off for both, see below)
                if (alt && (tcount & 1u) != group) continue;
                int m, r, nsplit;
                if (wsplit) { nsplit = idx % nsplits; r = idx / nsplits; m = r % mtiles; r /= mtiles; }
                else { m = idx % mtiles; r = idx / mtiles; nsplit = r % nsplits; r /= nsplits; }
                const int g = r % G, img = r / G;
                const int ty = m / tiles_x;
                // packed halo mode: M tile mt holds the pixels x = x0 + pack * rx + mt of the same 16 rows
                const int oy_s = ty * super_h + ry, ox_s = (m - ty * tiles_x) * tile_w + rx * pack;
                const int mt_dy = pack > 1 ? 0 : tile_h, mt_dx = pack > 1 ? 1 : 0;
                const int oy_f = oy_s + mt_lo * mt_dy, ox_f = ox_s + mt_lo * mt_dx;      // this group's first M tile
                const int nbase = nsplit * ntile;
                const bool has_bias = L.has_bias != 0;
                const float* bias = s_bias + (has_bias ? g * cout_pad + nbase : 0);
                const uint32_t taddr0 = tmem_base + buf * (uint32_t)kMaxNTile + ((uint32_t)(q * 32) << 16);
                const __half* resb = (epilogue == 0) ? L.res[img] : nullptr;
                // ---- plain epilogue (epilogue == 0): residual in and result out go through a per-warp staging tile ----
                // One lane per pixel would touch 32 different lines per LDG.128 / STG.128 (the L1 data pipe it shares
                // with the tensor core's operand reads became the bound); instead four lanes move one pixel's 64 bytes
                // of a 32-channel chunk, each lane swaps through shared memory (16-byte pieces XOR-swizzled by the row
                // pair) and reads / writes its own pixel row there.  The next chunk's residual is in flight while
                // the current one is computed; the first one is fetched before the accumulator is waited for.
                if (STAGED && epilogue == 0) {
                    const int kq = lane & 3, sb = lane >> 2;
                    const int ncc = c_hi > c_lo ? (c_hi - c_lo + 31) >> 5 : 0;
                    const int total_chunks = (mt_hi - mt_lo) * ncc;
                    const __half* res2b = FULL ? L.res2[img] : nullptr;
                    __half* outb = reinterpret_cast<__half*>(L.out[img]);
                    __half* out1b = FULL ? reinterpret_cast<__half*>(L.out1[img]) : nullptr;
                    __half* out2b = FULL ? reinterpret_cast<__half*>(L.out2[img]) : nullptr;
                    // element offset of this lane's pixel of M tile mt (channel nbase), -1 outside the image
                    auto pix_off = [&](int mt) -> int {
                        const int oy = oy_s + mt * mt_dy, ox = ox_s + mt * mt_dx;
                        if (oy >= OH || ox >= OW) return -1;
                        const size_t pix = os == 1 ? (size_t)oy * OW + ox
                                                   : (size_t)(oy * os + (g >> 1)) * (OW * os) + (ox * os + (g & 1));
                        return (int)(pix * cstride) + nbase;
                    };
                    uint4 rnext[4];
                    if (resb && total_chunks > 0 && !(dbg & 16)) {
                        const int mo = pix_off(mt_lo);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int o = __shfl_sync(0xffffffffu, mo, sb + 8 * i);
                            if (o >= 0 && c_lo + kq * 8 < c_hi) rnext[i] = *reinterpret_cast<const uint4*>(resb + o + c_lo + kq * 8);
                        }
                    }
                    mbar_wait(&acc_full[buf], use & 1u);
                    tc_fence_after();
                    if (tl < 10 && q == 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + 8 + group * 2);
                    int mt = mt_lo, c0 = c_lo;
                    for (int n = 0; n < total_chunks; ++n) {
                        const int myoff = pix_off(mt);
                        int offs[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) offs[i] = __shfl_sync(0xffffffffu, myoff, sb + 8 * i);
                        const uint32_t taddr = taddr0 + (uint32_t)(mt * ntile);
                        uint32_t rr[32];
                        tc_ld16_nowait(taddr + c0, rr);
                        if (c0 + 16 < c_hi) tc_ld16_nowait(taddr + c0 + 16, rr + 16);
                        // next chunk of this group
                        int mt2 = mt, c02 = c0 + 32;
                        if (c02 >= c_hi) { c02 = c_lo; ++mt2; }
                        const bool use_res = resb && !(dbg & 16);
                        if (use_res) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (c0 + kq * 8 < c_hi) stg[stg_slot(sb + 8 * i, kq)] = rnext[i];
                            __syncwarp();
                            if (n + 1 < total_chunks) {
                                const int mo = mt2 == mt ? myoff : pix_off(mt2);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int o = __shfl_sync(0xffffffffu, mo, sb + 8 * i);
                                    if (o >= 0 && c02 + kq * 8 < c_hi) rnext[i] = *reinterpret_cast<const uint4*>(resb + o + c02 + kq * 8);
                                }
                            }
                        }
                        tc_ld_wait();
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int c = c0 + hh * 16;
                            if (c >= c_hi) continue;
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 b4 = has_bias ? *reinterpret_cast<const float4*>(bias + c + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                                v[i] = __uint_as_float(rr[hh * 16 + i]) + b4.x;
                                v[i + 1] = __uint_as_float(rr[hh * 16 + i + 1]) + b4.y;
                                v[i + 2] = __uint_as_float(rr[hh * 16 + i + 2]) + b4.z;
                                v[i + 3] = __uint_as_float(rr[hh * 16 + i + 3]) + b4.w;
                            }
                            if (use_res) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const uint4 rv = stg[stg_slot(lane, hh * 2 + h)];
                                    const __half2* hp = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const float2 f = __half22float2(hp[k]);
                                        v[h * 8 + k * 2] += f.x; v[h * 8 + k * 2 + 1] += f.y;
                                    }
                                }
                            }
                            if (FULL && res2b && myoff >= 0) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const uint4 rv = reinterpret_cast<const uint4*>(res2b + myoff + c)[h];
                                    const __half2* hp = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const float2 f = __half22float2(hp[k]);
                                        v[h * 8 + k * 2] += f.x; v[h * 8 + k * 2 + 1] += f.y;
                                    }
                                }
                            }
                            // up to three outputs of the same pre-activation value, each with its own activation
#pragma unroll
                            for (int oi = 0; oi < (FULL ? 3 : 1); ++oi) {
                                __half* ob = oi == 0 ? outb : (oi == 1 ? out1b : out2b);
                                if (!ob) continue;
                                const int a = oi == 0 ? act : (oi == 1 ? act1 : act2);
                                const float sl = a == 1 ? 0.2f : (a == 3 ? 0.0f : (oi == 0 ? slope0 : (oi == 1 ? slope1 : slope2)));
                                float w[16];
                                if (a == 0) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) w[i] = v[i];
                                } else if (FULL && a == 2) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) w[i] = v[i] > 0.0f ? v[i] : s_slope[nbase + c + i] * v[i];
                                } else if (a == 5) {      // exact GELU (nn.GELU default)
#pragma unroll
                                    for (int i = 0; i < 16; ++i) w[i] = gelu_exact(v[i]);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) w[i] = v[i] > 0.0f ? v[i] : sl * v[i];
                                }
                                uint4 o[2];
                                __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
                                for (int k = 0; k < 8; ++k) oh[k] = __floats2half2_rn(w[2 * k], w[2 * k + 1]);
                                if (oi == 0) {
                                    stg[stg_slot(lane, hh * 2)] = o[0];
                                    stg[stg_slot(lane, hh * 2 + 1)] = o[1];
                                } else if (myoff >= 0 && nbase + c < cout) {
                                    uint4* o4 = reinterpret_cast<uint4*>(ob + myoff + c);
                                    o4[0] = o[0]; o4[1] = o[1];
                                }
                            }
                        }
                        __syncwarp();
                        {
                            const int cc = c0 + kq * 8;          // this lane's 8 columns; whole 16-column groups beyond cout stay unwritten
                            const bool wr = cc < c_hi && nbase + (cc & ~15) < cout && !(dbg & 8);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (wr && offs[i] >= 0) *reinterpret_cast<uint4*>(outb + offs[i] + cc) = stg[stg_slot(sb + 8 * i, kq)];
                        }
                        __syncwarp();
                        mt = mt2; c0 = c02;
                    }
                } else {
                // the residual does not depend on the accumulator: fetch (up to) 32 channels of the first M tile before waiting
                uint4 rpre[4];
                {
                    const bool v0 = oy_f < OH && ox_f < OW;
                    if (resb && v0) {
                        const __half* res = resb + ((size_t)oy_f * OW + ox_f) * cstride + nbase + c_lo;
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            if (c_lo + h * 8 < c_hi) rpre[h] = reinterpret_cast<const uint4*>(res)[h];
                    }
                }
                mbar_wait(&acc_full[buf], use & 1u);
                tc_fence_after();
                if (tl < 10 && q == 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + 8 + group * 2);
                for (int mt = mt_lo; mt < mt_hi; ++mt) {
                    const int oy = oy_s + mt * mt_dy, ox = ox_s + mt * mt_dx;
                    const bool valid = oy < OH && ox < OW;
                    const uint32_t taddr = taddr0 + (uint32_t)(mt * ntile);
                    if (!STAGED && epilogue == 0) {
                        const size_t pix = os == 1 ? (size_t)oy * OW + ox
                                                   : (size_t)(oy * os + (g >> 1)) * (OW * os) + (ox * os + (g & 1));
                        __half* out = reinterpret_cast<__half*>(L.out[img]) + pix * cstride + nbase;
                        __half* out1 = (FULL && L.out1[img]) ? reinterpret_cast<__half*>(L.out1[img]) + pix * cstride + nbase : nullptr;
                        __half* out2 = (FULL && L.out2[img]) ? reinterpret_cast<__half*>(L.out2[img]) + pix * cstride + nbase : nullptr;
                        const __half* res = (resb && valid && !(dbg & 16)) ? resb + pix * cstride + nbase : nullptr;
                        const __half* res2 = (FULL && L.res2[img] && valid) ? L.res2[img] + pix * cstride + nbase : nullptr;
                        for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
                            uint32_t rr[32];
                            tc_ld16_nowait(taddr + c0, rr);
                            if (c0 + 16 < c_hi) tc_ld16_nowait(taddr + c0 + 16, rr + 16);
                            tc_ld_wait();
                            if (!valid) continue;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int c = c0 + hh * 16;
                                if (c >= c_hi || nbase + c >= cout) continue;
                                float v[16];
#pragma unroll
                                for (int i = 0; i < 16; i += 4) {
                                    const float4 b4 = has_bias ? *reinterpret_cast<const float4*>(bias + c + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                                    v[i] = __uint_as_float(rr[hh * 16 + i]) + b4.x;
                                    v[i + 1] = __uint_as_float(rr[hh * 16 + i + 1]) + b4.y;
                                    v[i + 2] = __uint_as_float(rr[hh * 16 + i + 2]) + b4.z;
                                    v[i + 3] = __uint_as_float(rr[hh * 16 + i + 3]) + b4.w;
                                }
                                if (res) {
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {
                                        uint4 rv;
                                        if (mt == mt_lo && c0 == c_lo) rv = rpre[hh * 2 + h];
                                        else rv = reinterpret_cast<const uint4*>(res + c)[h];
                                        const __half2* hp = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            const float2 f = __half22float2(hp[k]);
                                            v[h * 8 + k * 2] += f.x; v[h * 8 + k * 2 + 1] += f.y;
                                        }
                                    }
                                }
                                if (res2) {
#pragma unroll
                                    for (int h = 0; h < 2; ++h) {
                                        const uint4 rv = reinterpret_cast<const uint4*>(res2 + c)[h];
                                        const __half2* hp = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            const float2 f = __half22float2(hp[k]);
                                            v[h * 8 + k * 2] += f.x; v[h * 8 + k * 2 + 1] += f.y;
                                        }
                                    }
                                }
                                // up to three outputs of the same pre-activation value, each with its own activation
#pragma unroll
                                for (int oi = 0; oi < (FULL ? 3 : 1); ++oi) {
                                    __half* op = oi == 0 ? out : (oi == 1 ? out1 : out2);
                                    if (!op) continue;
                                    const int a = oi == 0 ? act : (oi == 1 ? act1 : act2);
                                    const float sl = a == 1 ? 0.2f : (a == 3 ? 0.0f : (oi == 0 ? slope0 : (oi == 1 ? slope1 : slope2)));
                                    float w[16];
                                    if (a == 0) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i) w[i] = v[i];
                                    } else if (FULL && a == 2) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i) w[i] = v[i] > 0.0f ? v[i] : s_slope[nbase + c + i] * v[i];
                                    } else if (a == 5) {      // exact GELU (nn.GELU default)
#pragma unroll
                                        for (int i = 0; i < 16; ++i) w[i] = gelu_exact(v[i]);
                                    } else {
#pragma unroll
                                        for (int i = 0; i < 16; ++i) w[i] = v[i] > 0.0f ? v[i] : sl * v[i];
                                    }
                                    uint4 o[2];
                                    __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
                                    for (int k = 0; k < 8; ++k) oh[k] = __floats2half2_rn(w[2 * k], w[2 * k + 1]);
                                    uint4* o4 = reinterpret_cast<uint4*>(op + c);
                                    if (!(dbg & 8)) { o4[0] = o[0]; o4[1] = o[1]; }
                                }
                            }
                        }
                    } else {
                        // lastconv: group g = phase (py, px); channel co = c13*4 + i*2 + j lands at
                        // out[4*oy + 2*py + i][4*ox + 2*px + j][c13]  (ConvTranspose phase + PixelShuffle(2))
                        const int ph = G == 1 ? nsplit : g;       // 3x3 form: the N split is the phase
                        const int py = ph >> 1, px = ph & 1;
                        const int OW4 = OW * 4;
                        float* out = reinterpret_cast<float*>(L.out[img]);
                        // out_cstride = floats per pixel: 16, or 8 when only the channels 0..7 are wanted (the last IFBlock's
                        // consumer reads flow and mask only: half the store traffic of the largest lastconv)
                        const int c_end = cstride == 8 ? (c_hi < 32 ? c_hi : 32) : c_hi;
                        for (int c0 = c_lo; c0 < c_end; c0 += 32) {      // ntile == 64 (checked on the host)
                            uint32_t rr[32];
                            tc_ld16_nowait(taddr + c0, rr);
                            tc_ld16_nowait(taddr + c0 + 16, rr + 16);
                            tc_ld_wait();
                            if (L.staged) {
                                // Through the per-warp staging tile: a lane's four 16-byte pieces of one output row (two
                                // sub-pixels x two channel quads) are 64 contiguous bytes (32 + 32 with 16 floats per pixel);
                                // written lane-per-row they scatter into 32 lines per STG.128 (the layer was bound by LSU
                                // wavefronts: 4600 cycles per 256-pixel tile), written four lanes per row into 8.
                                float vv[2][16];
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                                    for (int i = 0; i < 16; ++i) vv[hh][i] = __uint_as_float(rr[hh * 16 + i]) + bias[c0 + hh * 16 + i];
                                const int kq = lane & 3, sb = lane >> 2;
                                const int doff = (kq >> 1) * cstride + ((c0 + (kq & 1) * 16) >> 2);
#pragma unroll
                                for (int i2 = 0; i2 < 2; ++i2) {
                                    const int myoff = valid ? (int)(((size_t)(4 * oy + 2 * py + i2) * OW4 + 4 * ox + 2 * px) * cstride) : -1;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const int ij = i2 * 2 + (k >> 1), hh = k & 1;
                                        const float4 o = make_float4(vv[hh][0 + ij], vv[hh][4 + ij], vv[hh][8 + ij], vv[hh][12 + ij]);
                                        reinterpret_cast<float4*>(stg)[stg_slot(lane, k)] = o;
                                    }
                                    __syncwarp();
#pragma unroll
                                    for (int t = 0; t < 4; ++t) {
                                        const int o = __shfl_sync(0xffffffffu, myoff, sb + 8 * t);
                                        if (o >= 0) *reinterpret_cast<float4*>(out + o + doff) = reinterpret_cast<const float4*>(stg)[stg_slot(sb + 8 * t, kq)];
                                    }
                                    __syncwarp();
                                }
                                continue;
                            }
                            if (!valid) continue;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                const int c = c0 + hh * 16;
                                float v[16];
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[hh * 16 + i]) + bias[c + i];
#pragma unroll
                                for (int ij = 0; ij < 4; ++ij) {
                                    const int yy = 4 * oy + 2 * py + (ij >> 1), xx = 4 * ox + 2 * px + (ij & 1);
                                    const float4 o = make_float4(v[0 + ij], v[4 + ij], v[8 + ij], v[12 + ij]);
                                    *reinterpret_cast<float4*>(out + ((size_t)yy * OW4 + xx) * cstride + (c >> 2)) = o;
                                }
                            }
                        }
                    }
                }
                }
                if (tl < 10 && q == 2 && lane == 0) TC_TRACE(li * 256 + tl * 16 + 9 + group * 2);
                // accumulator drained: hand the buffer back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&acc_empty[buf]); if (alt) mbar_arrive(&acc_empty[buf]); }   // 8 arrivals free a buffer
            }
        }

        if (li + 1 < nlayers) {
            // layer boundary: this layer's outputs (generic-proxy stores) must be visible to every CTA's
            // TMA loads (async proxy) of the next layer
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            // this CTA's MMAs are done: the next layer's resident weights may overwrite the weight region while
            // the grid barrier is pending (weights do not depend on other CTAs)
            if (warp == 10) load_resident_weights(li + 1);
            if (threadIdx.x == 0) { TC_TRACE(li * 256 + 202); grid_barrier(prog.sync, (unsigned)(li + 1) * gridDim.x, prog.err); TC_TRACE(li * 256 + 203); }
            stage_tables(li + 1);
            __syncthreads();
            asm volatile("fence.proxy.async;" ::: "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) TC_TRACE(4092);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
    if (nlayers > 1 && threadIdx.x == 0) {
        // the last CTA to leave re-zeroes the barrier state for the next launch
        const unsigned old = atomicAdd(prog.sync + 1, 1u);
        if (old == gridDim.x - 1) {
            prog.sync[0] = 0u;
            prog.sync[1] = 0u;
            __threadfence();
        }
    }
}

// ---- host side ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static CUtensorMapSwizzle swz_enum(int bytes)
{
    return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static thread_local int t_ntile_cap = 1 << 30;      // build_layer retries (see the halo fall-back)
static thread_local bool t_need_halo = false;
static thread_local bool t_no_narrow = false;     // set while a layer is rebuilt because its split-resident form does not fit the grid

static int build_layer(const drba_conv_layer& d, int nimg, LayerDev& L, EncodeTiledFn encode)
{
    const int H = d.H, W = d.W, Cin = d.Cin, G = d.G, T = d.T, S = d.S, OH = d.OH, OW = d.OW;
    if (!d.w) return DRBA_E_ARG;
    if (!d.bias && d.epilogue != 0) return DRBA_E_ARG;
    if (H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || Cin <= 0 || Cin % 16 != 0) return DRBA_E_ARG;
    if (G < 1 || G > kMaxGroups || T < 1 || T > kMaxTapsTc || G * T > 36) return DRBA_E_ARG;
    if (S != 1 && S != 2) return DRBA_E_ARG;
    if (S == 2 && (H % 2 != 0 || W % 2 != 0)) return DRBA_E_ARG;
    if (d.cout_pad <= 0 || d.cout_pad % 16 != 0 || d.cout <= 0 || d.cout > d.cout_pad) return DRBA_E_ARG;
    if (d.epilogue != 0 && d.epilogue != 1) return DRBA_E_ARG;
    // lastconv: four phase convs of four taps (G = 4, 64 padded channels each), or -- same arithmetic, zero weights on
    // the five taps a phase does not use -- ONE 3x3 conv with 4 x 64 output columns whose N split IS the phase
    const bool last3x3 = d.epilogue == 1 && G == 1 && T == 9 && d.cout_pad == 256 && S == 1;
    if (d.epilogue == 1 && !last3x3 && (d.cout_pad != 64 || d.cout != 52 || G != 4)) return DRBA_E_ARG;
    if (d.epilogue == 1 && (d.cout != 52 || (d.out_cstride != 16 && d.out_cstride != 8))) return DRBA_E_ARG;
    if (d.out_os != 1 && d.out_os != 2) return DRBA_E_ARG;
    if (d.epilogue == 0 && ((d.out_os == 1 && G != 1) || (d.out_os == 2 && G != 4))) return DRBA_E_ARG;
    if (d.epilogue == 0 && (d.out_cstride < d.cout_pad || d.out_cstride % 8 != 0)) return DRBA_E_ARG;
    if (d.act < 0 || d.act > 5 || (d.act == 2 && !d.slope)) return DRBA_E_ARG;
    if (d.act1 < 0 || d.act1 > 4 || d.act1 == 2 || d.act2 < 0 || d.act2 > 4 || d.act2 == 2) return DRBA_E_ARG;
    if (d.out_os != 1 && (d.res[0] || d.res[1])) return DRBA_E_ARG;   // residuals only for same-geometry layers
    if (d.bias && G * d.cout_pad > 512) return DRBA_E_UNSUPPORTED;   // bias / slope tables staged in shared memory
    if (d.bgemm && (G != 1 || T != 1 || S != 1 || d.epilogue != 0 || d.bias)) return DRBA_E_ARG;
    if (!aligned16(d.w)) return DRBA_E_ALIGN;
    for (int i = 0; i < nimg; ++i) {
        if (!d.in[i] || !d.out[i]) return DRBA_E_ARG;
        if (!aligned16(d.in[i]) || !aligned16(d.out[i]) || (d.res[i] && !aligned16(d.res[i]))) return DRBA_E_ALIGN;
    }

    memset(&L, 0, sizeof(L));
    L.OH = OH; L.OW = OW; L.S = S; L.T = T; L.G = G;
    L.bgemm = d.bgemm ? 1 : 0; L.has_bias = d.bias ? 1 : 0;
    L.Kc = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
    L.kchunks = Cin / L.Kc;
    L.swz_bytes = L.Kc * 2;
    for (int gi = 0; gi < G; ++gi)
        for (int t = 0; t < T; ++t) {
            const int dy = d.dy[gi * T + t], dx = d.dx[gi * T + t];
            if (dy < -32 || dy > 32 || dx < -32 || dx > 32) return DRBA_E_ARG;
            int cy = dy, cx = dx, pary = 0, parx = 0;
            // input coordinate = S * o + d  ->  (parity, cell) in the [H/S][S][W/S][S][C] view
            if (S == 2) { pary = dy & 1; parx = dx & 1; cy = (dy - pary) >> 1; cx = (dx - parx) >> 1; }
            L.tapc[gi][t] = (cy + 64) | ((cx + 64) << 8) | (pary << 16) | (parx << 17);
        }
    // N tile: <= 128 columns (<= t_ntile_cap when a retry asks for narrower tiles, see the halo fall-back below), a
    // multiple of 16 that divides cout_pad
    int ntile = 0;
    const int ncap = t_ntile_cap < kMaxNTile ? t_ntile_cap : kMaxNTile;
    for (int cand = d.cout_pad < ncap ? d.cout_pad : ncap; cand >= 16; cand -= 16)
        if (d.cout_pad % cand == 0) { ntile = cand; break; }
    if (!ntile) return DRBA_E_UNSUPPORTED;
    if (last3x3) ntile = 64;          // one phase per N split (the pixel-shuffle epilogue reads 64 columns)
    // halo mode: a 3x3 stride-1 layer loads ONE (16*MT+2) x 16 pixel neighbourhood per K chunk and runs the nine
    // taps as descriptor offsets into it (8-pixel-wide tiles keep every 8-row core group contiguous; the 16-pixel
    // pitch keeps the swizzle phase identical for all groups) -- 4x fewer activation bytes than one box per tap.
    // Needs the weights resident (checked below once ntile is final).
    static int env_halo = -1, env_pack = -1, env_s2 = 1;
    if (env_halo < 0) {
        const char* e = getenv("DRBA_TC_HALO"); env_halo = e ? atoi(e) : 1;
        e = getenv("DRBA_TC_PACK"); env_pack = e ? atoi(e) : 1;
        e = getenv("DRBA_TC_HALO_S2"); env_s2 = e ? atoi(e) : 1;
    }
    // T == 10: canonical 3x3 taps + an identity tap at (0, 0) (ResConv residual through the tensor core; stride 1 only)
    const bool restap = T == 10 && S == 1 && d.dy[9] == 0 && d.dx[9] == 0;
    bool halo = env_halo && (S == 1 || (S == 2 && env_s2)) && (T == 9 || restap) && G == 1 && (d.epilogue == 0 || last3x3) && !d.bgemm;
    if (halo)
        for (int t = 0; t < 9; ++t)
            if (d.dy[t] != t / 3 - 1 || d.dx[t] != t % 3 - 1) halo = false;
    const bool halo_s2 = halo && S == 2;      // parity-plane boxes (see issue_halo_s2)
    // M tiles per super tile: thin K steps stack M tiles so that one TMA box still carries ~16 KB
    int MT = kABytesMax / (kTileM * L.Kc * 2);
    while (MT > 1 && MT * ntile > kMaxNTile) MT >>= 1;
    long best = -1;
    int best_mt = 1;
    L.tile_w = 16; L.tile_h = 8;
    // packed halo mode: thin inputs (16 / 32 channels) put 4 / 2 pixels on one 128-byte line, so the halo box has
    // 10 x 18 lines for an 8P x 16 pixel tile instead of 16 x (16 MT + 2) short ones (TMA time goes by the line)
    int pack = 1;
    if (halo && !halo_s2 && env_pack && L.kchunks == 1 && L.Kc < 64 && W % (64 / L.Kc) == 0 && ntile * (64 / L.Kc) <= kMaxNTile)
        pack = 64 / L.Kc;
    if (pack > 1) {
        L.tile_w = 8 * pack; L.tile_h = 16;
        best_mt = pack;
    } else if (halo_s2) {
        L.tile_w = 8; L.tile_h = 16;
        best_mt = 1;
    } else if (halo) {
        L.tile_w = 8; L.tile_h = 16;
        while (MT > 1 && (OH + 16 * MT - 1) / (16 * MT) * MT * 100 > ((OH + 15) / 16) * 106) MT >>= 1;   // avoid > 6 % more rows
        best_mt = MT;
    } else {
        // pixel patch of an M tile: the rectangle whose super tiles cover the output with the least waste
        const int shapes[5][2] = {{16, 8}, {32, 4}, {8, 16}, {64, 2}, {128, 1}};
        for (int mt = MT; mt >= 1; mt >>= 1) {
            for (int i = 0; i < 5; ++i) {
                const int sh = shapes[i][1] * mt;
                if (sh > 256) continue;
                if (d.bgemm && sh != 1) continue;      // a tile must not straddle batch elements
                const long nt = (long)((OW + shapes[i][0] - 1) / shapes[i][0]) * ((OH + sh - 1) / sh);
                const long cost = nt * mt;       // covered M tiles (waste included)
                // prefer the larger super tile unless it wastes > 6 % more coverage
                if (best < 0 || cost * 100 < best * 94 || (mt == best_mt && cost < best)) {
                    best = cost; best_mt = mt; L.tile_w = shapes[i][0]; L.tile_h = shapes[i][1];
                }
            }
        }
    }
    MT = best_mt;
    L.MT = MT;
    L.tiles_x = (OW + L.tile_w - 1) / L.tile_w;
    L.mtiles = L.tiles_x * (pack > 1 ? (OH + 15) / 16 : (OH + L.tile_h * MT - 1) / (L.tile_h * MT));
    // small layers: split N further so that more SMs stream the K loop in parallel
    if (d.epilogue == 0) {
        while (ntile >= 64 && ntile % 32 == 0 && nimg * L.mtiles * G * (d.cout_pad / ntile) * 2 <= kNumSMs) ntile /= 2;
    }
    L.ntile = ntile; L.nsplits = d.cout_pad / ntile; L.cout_pad = d.cout_pad; L.cout = d.cout;
    L.total_tiles = nimg * G * L.nsplits * L.mtiles;
    L.epilogue = d.epilogue; L.act = d.act; L.bias = d.bias; L.slope = d.slope;
    L.out_cstride = d.out_cstride; L.os = d.out_os;
    for (int i = 0; i < nimg; ++i) {
        L.res[i] = (const __half*)d.res[i]; L.res2[i] = (const __half*)d.res2[i];
        L.out[i] = d.out[i]; L.out1[i] = d.out1[i]; L.out2[i] = d.out2[i];
        if ((d.res2[i] && !aligned16(d.res2[i])) || (d.out1[i] && !aligned16(d.out1[i])) || (d.out2[i] && !aligned16(d.out2[i]))) return DRBA_E_ALIGN;
        if (d.epilogue != 0 && (d.res2[i] || d.out1[i] || d.out2[i])) return DRBA_E_ARG;
    }
    L.act1 = d.act1; L.act2 = d.act2; L.slope0 = d.slope0; L.slope1 = d.slope1; L.slope2 = d.slope2;

    const int b_bytes = ntile * L.Kc * 2;
    const int swz_period = 8 * L.swz_bytes;
    L.b_sub = (b_bytes + swz_period - 1) / swz_period * swz_period;
    if (L.b_sub > kStageBytes - kABytesMax) return DRBA_E_UNSUPPORTED;
    // shared-memory plan.  Resident mode: [all weight tiles of the layer | activation ring]; streaming mode: six
    // stages of [A 16 KB | B 16 KB].  Halo mode needs resident weights.
    const long total_smem = (long)kRingBytes;
    // (inside a narrower-N retry the resident region holds ONE N split: see wsplit)
    const long w_bytes = ((long)G * (t_need_halo ? 1 : L.nsplits) * T * L.kchunks * L.b_sub + 1023) / 1024 * 1024;
    const long a_stage = halo_s2 ? 4 * (((long)17 * 9 * L.Kc * 2 + 1023) / 1024 * 1024) : pack > 1 ? (10 * 18 * 128 + 1023) / 1024 * 1024
                                  : (halo ? (((long)(16 * MT + 2) * 16 * L.Kc * 2 + 1023) / 1024 * 1024) : kABytesMax);
    static int env_res = -1;
    if (env_res < 0) { const char* e = getenv("DRBA_TC_RESIDENT"); env_res = e ? atoi(e) : 1; }
    bool resident = env_res && !d.bgemm && w_bytes + 2 * a_stage <= total_smem && w_bytes <= (long)kBRegion + 32768;
    if (halo && !resident) {
        // The coarse IFNet levels (96 / 128 / 192 channels at 1/16 ... 1/64 of the frame) are bound by the bytes a CTA
        // pulls from L2 (measured ~29 B per clock and SM whatever the box shape): with one box per tap a 128-pixel tile
        // re-reads its input nine times (block0.res: 27 K iterations x 20 KB = 540 KB per CTA, 700 cycles each).  A
        // NARROWER N tile whose weights fit the resident region keeps the halo mode -- the input is read once per K
        // chunk (block0.res with 32 columns: 110 KB of weights + 3 x 36 KB) -- and spreads the layer over more SMs.
        static int env_narrow = -1;
        if (env_narrow < 0) { const char* e = getenv("DRBA_TC_NARROW"); env_narrow = e ? atoi(e) : 1; }
        // Only for layers in the latency regime (<= 2 tiles per SM): large layers (GridNet's 128-channel level at 272 x 480)
        // are better off streaming with wide N tiles (measured 55 vs 77 us).
        if (env_narrow && !t_no_narrow && !t_need_halo && !halo_s2 && !d.bgemm && ntile >= 16 && (last3x3 || L.total_tiles <= 2 * kNumSMs)) {
            const int saved = t_ntile_cap;
            int rc = DRBA_E_UNSUPPORTED;
            // (first the same width with one split resident, then narrower ones; the 3x3 lastconv keeps its 64 columns)
            for (int cap = ntile; cap >= (last3x3 ? ntile : 16); cap -= 16) {
                if (d.cout_pad % cap != 0) continue;
                t_ntile_cap = cap;
                t_need_halo = true;
                rc = build_layer(d, nimg, L, encode);
                t_need_halo = false;
                if (rc == DRBA_OK) break;
            }
            t_ntile_cap = saved;
            if (rc == DRBA_OK) return rc;
        }
        if (t_need_halo) return DRBA_E_UNSUPPORTED;      // (inside such a retry: this cap does not fit either)
        // fall back to one box per tap: redo the tile choice without the halo constraint
        halo = false;
        static thread_local int depth = 0;
        if (depth == 0) {
            ++depth;
            const int saved = env_halo;
            env_halo = 0;
            const int rc = build_layer(d, nimg, L, encode);
            env_halo = saved;
            --depth;
            return rc;
        }
    }
    L.resident = resident ? 1 : 0;
    L.wsplit = (resident && t_need_halo) ? 1 : 0;
    L.halo = halo ? (halo_s2 ? 2 : 1) : 0;
    L.pack = halo ? pack : 1;
    {
        static int env_staged = -1;
        if (env_staged < 0) { const char* e = getenv("DRBA_TC_STAGED"); env_staged = e ? atoi(e) : 1; }
        // staging pays where rows are wide or the residual would otherwise be fetched chunk by chunk; latency-bound
        // streaming layers (one or two tiles per SM) keep the direct per-pixel path
        L.staged = (env_staged == 2 || (env_staged == 1 && resident && d.epilogue == 0)) ? 1 : 0;
        static int env_last = -1;
        if (env_last < 0) { const char* e = getenv("DRBA_TC_LASTSTG"); env_last = e ? atoi(e) : 1; }
        if (d.epilogue == 1) L.staged = env_last ? 1 : 0;      // lastconv: stores through the staging tile (64-byte runs)
    }
    L.KI = halo ? L.kchunks : T * L.kchunks;
    if (resident) {
        L.a_base = (int)w_bytes;
        L.a_stride = (int)a_stage;
        long nst = (total_smem - w_bytes) / a_stage;
        L.nst = (int)(nst > kStages ? kStages : nst);
    } else {
        L.a_base = 0; L.a_stride = kStageBytes; L.nst = kStages;
    }

    // A: input viewed as [H/S][S][W/S][S][C], innermost first
    for (int i = 0; i < nimg && L.pack > 1; ++i) {
        // packed halo: [H][W / P][P * Cin] with 128-byte lines
        const cuuint64_t dims[5] = {64, 1, (cuuint64_t)(W / L.pack), 1, (cuuint64_t)H};
        const cuuint64_t strides[4] = {128, 128, (cuuint64_t)W * Cin * 2, (cuuint64_t)W * Cin * 2};
        const cuuint32_t box[5] = {64, 1, 10, 1, 18};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUresult r = encode(&L.ta[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(d.in[i]), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return DRBA_E_UNSUPPORTED;
    }
    for (int i = 0; i < nimg && L.pack == 1; ++i) {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)S, (cuuint64_t)(W / S), (cuuint64_t)S, (cuuint64_t)(H / S)};
        const cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)S * Cin * 2, (cuuint64_t)W * Cin * 2,
                                       (cuuint64_t)S * W * Cin * 2};
        const cuuint32_t box[5] = {(cuuint32_t)L.Kc, 1, (cuuint32_t)(L.halo == 2 ? 9 : (L.halo ? 16 : L.tile_w)), 1,
                                   (cuuint32_t)(L.halo == 2 ? 17 : (L.halo ? 16 * L.MT + 2 : L.tile_h * L.MT))};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUresult r = encode(&L.ta[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(d.in[i]), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(L.swz_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return DRBA_E_UNSUPPORTED;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)Cin, d.bgemm ? (cuuint64_t)H * d.cout_pad : (cuuint64_t)G * T * d.cout_pad};
        const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        const cuuint32_t box[2] = {(cuuint32_t)L.Kc, (cuuint32_t)ntile};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&L.tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d.w), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(L.swz_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return DRBA_E_UNSUPPORTED;
    }
    return DRBA_OK;
}

static long long* g_trace = nullptr;

// sticky error word of the persistent programs: mapped pinned host memory, written by the kernel with system scope
// and read by the host WITHOUT synchronising (one-time allocation at the first multi-layer launch)
static unsigned* g_err_host = nullptr;
static unsigned* g_err_dev = nullptr;

static int ensure_err_word()
{
    if (g_err_host) return DRBA_OK;
    void* h = nullptr;
    cudaError_t e = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) return (int)e;
    memset(h, 0, 64);
    void* d = nullptr;
    e = cudaHostGetDevicePointer(&d, h, 0);
    if (e != cudaSuccess) { cudaFreeHost(h); return (int)e; }
    g_err_host = (unsigned*)h; g_err_dev = (unsigned*)d;
    return DRBA_OK;
}

// CTAs of conv_tc_kernel that can be resident at once on the current device (the grid barrier needs all of them)
template <typename K>
static int resident_ctas(K kernel, size_t smem)
{
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTcThreads, smem) != cudaSuccess) return 0;
    return sms * per_sm;
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_conv_tc_debug_trace(void* dev_buf_4096_i64)
{
    g_trace = (long long*)dev_buf_4096_i64;
    return DRBA_OK;
}

int drba_conv_tc_program_f16(const drba_conv_layer* layers, int nlayers, int nimg, void* sync_ws, void* stream)
{
    if (!layers || nlayers < 1 || nlayers > kMaxLayers || nimg < 1 || nimg > kMaxImages) return DRBA_E_ARG;
    if (nlayers > 1 && (!sync_ws || (reinterpret_cast<uintptr_t>(sync_ws) & 7u))) return DRBA_E_WORKSPACE;
    EncodeTiledFn encode = get_encode();
    if (!encode) return DRBA_E_UNSUPPORTED;

    static int env_pdl = -1, env_dbg = -1, env_grid = -1;
    if (env_pdl < 0) {
        const char* e = getenv("DRBA_TC_PDL"); env_pdl = e ? atoi(e) : 1;
        e = getenv("DRBA_TC_DBG"); env_dbg = e ? atoi(e) : 0;
        e = getenv("DRBA_TC_GRID"); env_grid = e ? atoi(e) : 0;
    }
    if (nlayers > 1) {
        const int rc = ensure_err_word();
        if (rc != DRBA_OK) return rc;
        if (*(volatile unsigned*)g_err_host) return DRBA_E_BARRIER;      // an earlier program lost a CTA: results are invalid
    }
    Program prog;
    prog.nlayers = nlayers; prog.nimg = nimg; prog.sync = (unsigned*)sync_ws; prog.dbg = env_dbg; prog.pad_ = 0;
    prog.err = g_err_dev;
    prog.trace = g_trace;
    int max_tiles = 0;
    for (int i = 0; i < nlayers; ++i) {
        const int rc = build_layer(layers[i], nimg, prog.L[i], encode);
        if (rc != DRBA_OK) return rc;
        if (prog.L[i].total_tiles > max_tiles) max_tiles = prog.L[i].total_tiles;
    }
    static bool attr_set = false;
    static int max_resident = 0;
    const size_t smem = (size_t)kRingBytes + 1024 + 8 * kStagingBytes;   // ring + alignment + epilogue staging
    if (!attr_set) {
        cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        // the grid barrier needs every CTA resident: never launch more than the device can hold at once
        max_resident = resident_ctas(conv_tc_kernel<true, true>, smem);
        attr_set = true;
    }
    if (max_resident < 1) return DRBA_E_UNSUPPORTED;
    int grid = max_tiles < max_resident ? max_tiles : max_resident;
    if (env_grid > 0 && env_grid < grid) grid = env_grid;
    for (int i = 0; i < nlayers; ++i) {
        // split-resident layers need every tile of a CTA in one N split
        if (prog.L[i].wsplit && !(prog.L[i].total_tiles <= grid || grid % prog.L[i].nsplits == 0)) {
            t_no_narrow = true;
            const int rc = build_layer(layers[i], nimg, prog.L[i], encode);
            t_no_narrow = false;
            if (rc != DRBA_OK) return rc;
        }
    }
    {
        // alternate-tile epilogue where a CTA walks at least four tiles of the layer (DRBA_TC_ALT: 0 never, 1 auto, 2 always)
        static int env_alt = -1;
        if (env_alt < 0) { const char* e = getenv("DRBA_TC_ALT"); env_alt = e ? atoi(e) : 1; }
        for (int i = 0; i < nlayers; ++i)
            prog.L[i].alt = env_alt == 2 || (env_alt == 1 && prog.L[i].total_tiles >= 4 * grid) ? 1 : 0;
    }
    bool full = false;
    for (int i = 0; i < nlayers; ++i) {
        const drba_conv_layer& d = layers[i];
        if (d.act == 2 || d.act == 4 || d.act1 || d.act2) full = true;
        for (int k = 0; k < nimg; ++k)
            if (d.res2[k] || d.out1[k] || d.out2[k]) full = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap our prologue with the previous kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = env_pdl ? 1 : 0;
    // staged epilogue when most plain layers of the program keep their weights resident (halo layers: many tiles per SM)
    int n_staged = 0, n_plain = 0;
    for (int i = 0; i < nlayers; ++i)
        if (layers[i].epilogue == 0) { ++n_plain; n_staged += prog.L[i].staged; }
    const bool staged = n_plain > 0 && 2 * n_staged >= n_plain;
    const cudaError_t le = full ? (staged ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true>, prog) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false>, prog))
                                : (staged ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, true>, prog) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false>, prog));
    if (le != cudaSuccess) return (int)le;
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_conv_tc_status(void)
{
    // no synchronisation: reports what the device has written so far (call after a stream / device sync for a
    // definite answer about the work enqueued before it)
    if (g_err_host && *(volatile unsigned*)g_err_host) return DRBA_E_BARRIER;
    return DRBA_OK;
}

int drba_conv_tc_f16(const void* in, int H, int W, int Cin,
                     const void* w, const float* bias, int G, int T, const int* dy, const int* dx,
                     int cout_pad, int cout, int S, int OH, int OW,
                     int epilogue, int act, const void* res, void* out, int out_cstride, int out_os, void* stream)
{
    if (!in || !w || !bias || !out || !dy || !dx) return DRBA_E_ARG;
    if (G < 1 || G > kMaxGroups || T < 1 || T > kMaxTapsTc || G * T > 36) return DRBA_E_ARG;
    drba_conv_layer d;
    memset(&d, 0, sizeof(d));
    d.in[0] = in; d.res[0] = res; d.out[0] = out;
    d.H = H; d.W = W; d.Cin = Cin; d.w = w; d.bias = bias; d.slope = nullptr;
    d.G = G; d.T = T;
    for (int i = 0; i < G * T; ++i) { d.dy[i] = dy[i]; d.dx[i] = dx[i]; }
    d.cout_pad = cout_pad; d.cout = cout; d.S = S; d.OH = OH; d.OW = OW;
    d.epilogue = epilogue; d.act = act; d.out_cstride = out_cstride; d.out_os = out_os;
    return drba_conv_tc_program_f16(&d, 1, 1, nullptr, stream);
}

}  // extern "C"
