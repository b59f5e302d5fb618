// Forward (summation) splat, OWNER-COMPUTES path for many-channel inputs (sm_100a).
//
// Same operator as csrc/splat.cu (models/softsplat/softsplat.py:248-293 + kernel :306-357,
// arithmetic per SURVEY.md appendix A.1), different algorithm.  The scatter formulation needs
// 4 x (C + 1) float atomics per source pixel; measured on B200 the L2 atomic units cap it at
// ~0.9 TB/s of payload, i.e. 10-13 % of the HBM roofline for the 64..192-channel feature splats
// of GMFSS.  The splat is a sparse matrix (<= 4 non-zeros per source pixel) applied to C
// columns, so here the matrix is built once per (flow, metric) -- cost independent of C -- and
// applied as a gather with NO float atomics:
//
//   1. count   : every source pixel bumps an int32 counter of each in-bounds target corner
//   2. scan    : exclusive prefix sum of the counters -> {offset, count} per target pixel
//   3. fill    : every source pixel writes (corner << 29 | source index) into a slot of each target's
//                list (slot = atomicSub on the counter, which thereby returns to zero)
//   4. gather  : one thread per TARGET pixel reads its list, sorts it by (corner, source index), recomputes
//                the bilinear corner weights from the flow, and accumulates all C channels in registers
//                in that order -- exactly the summation order of the reference's CPU path (corner-major:
//                softsplat_torch.py:146-174 adds all NW contributions in pixel order, then NE, SW, SE),
//                so results are run-to-run deterministic and (fp32, no FMA contraction) bit-identical
//                to the CPU restatement for sum / avg / linear; `soft` differs only through expf.  Normalisation (softsplat.py:273-290) happens in the
//                same pass and the output is written exactly once.
//
// HBM traffic: in read once + out written once + ~100 B/pixel of list traffic (C-independent).
// The workspace keeps the library-wide contract (all-zero on entry, all-zero on exit): counters
// return to zero in step 3, {offset, count} pairs and lists are cleared by a bulk memset after step 4.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace drba {

constexpr int kGThreads = 256;
constexpr int kScanThreads = 1024;
constexpr int kScanPerThread = 4;
constexpr int kScanBlock = kScanThreads * kScanPerThread;   // 4096 counters per scan block

struct GatherWs {
    int* count;        // [total]            zero on entry / exit
    int2* offcnt;      // [total]            {offset, count}
    int2* entries;     // [4 * total]   {key, corner weight}
    int* block_sums;   // [nblocks_scan + 4]
    int* done;         // [N * tiles of 64 x 16 targets]: 1 = served by the TMA-staged tile kernel
};

__device__ __forceinline__ void corner_targets(const Footprint& f, int H, int W, bool& nw, bool& ne, bool& sw, bool& se)
{
    const bool inx0 = f.x0 >= 0 && f.x0 < W, inx1 = f.x0 + 1 >= 0 && f.x0 + 1 < W;
    const bool iny0 = f.y0 >= 0 && f.y0 < H, iny1 = f.y0 + 1 >= 0 && f.y0 + 1 < H;
    nw = f.ok && inx0 && iny0; ne = f.ok && inx1 && iny0;
    sw = f.ok && inx0 && iny1; se = f.ok && inx1 && iny1;
}

// FILL = false: step 1 (count); FILL = true: step 3 (fill)
template <bool FILL>
__global__ void __launch_bounds__(kGThreads)
splat_list_kernel(const float* __restrict__ flow, int* __restrict__ count, const int2* __restrict__ offcnt,
                  int2* __restrict__ entries, int N, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * kGThreads + threadIdx.x;
    if (p >= (size_t)N * HW) return;
    const int n = (int)(p / HW);
    const size_t r = p - (size_t)n * HW;
    const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
    const float fx = flow[((size_t)n * 2) * HW + r];
    const float fy = flow[((size_t)n * 2 + 1) * HW + r];
    const Footprint f = footprint(x, y, fx, fy);
    bool c[4];
    corner_targets(f, H, W, c[0], c[1], c[2], c[3]);
    const long long q = (long long)n * (long long)HW + (long long)f.y0 * W + f.x0;
    const long long dq[4] = {0, 1, W, (long long)W + 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!c[k]) continue;
        const long long t = q + dq[k];
        if (!FILL) {
            atomicAdd(count + t, 1);
        } else {
            const int slot = atomicSub(count + t, 1) - 1;
            const float wk = k == 0 ? f.nw : (k == 1 ? f.ne : (k == 2 ? f.sw : f.se));
            entries[(size_t)offcnt[t].x + slot] = make_int2((int)(((unsigned)k << 29) | (unsigned)r), __float_as_int(wk));
        }
    }
}

// ---- exclusive scan of the counters (three small kernels) ------------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (kScanThreads / 32) ? smem_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        smem_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    const int warp_off = warp > 0 ? smem_warp[warp - 1] : 0;
    total = smem_warp[kScanThreads / 32 - 1];
    __syncthreads();
    return warp_off + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_block_sums_kernel(const int* __restrict__ count, int* __restrict__ block_sums, size_t total)
{
    __shared__ int sw[32];
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    int s = 0;
    if (base + 3 < total) {
        const int4 v = *reinterpret_cast<const int4*>(count + base);
        s = v.x + v.y + v.z + v.w;
    } else {
        for (int k = 0; k < kScanPerThread; ++k) if (base + k < total) s += count[base + k];
    }
    int tot;
    block_exclusive_scan(s, sw, tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(int* __restrict__ block_sums, int nblocks)
{
    __shared__ int sw[32];
    int carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += kScanThreads) {
        const int i = b0 + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, sw, tot);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += tot;
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_write_kernel(const int* __restrict__ count, int* __restrict__ block_sums, int2* __restrict__ offcnt, size_t total)
{
    __shared__ int sw[32];
    const size_t base = (size_t)blockIdx.x * kScanBlock + (size_t)threadIdx.x * kScanPerThread;
    int v[kScanPerThread];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) { v[k] = base + k < total ? count[base + k] : 0; s += v[k]; }
    int tot;
    int ex = block_exclusive_scan(s, sw, tot) + block_sums[blockIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = 0;      // scratch returns to zero
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        if (base + k < total) offcnt[base + k] = make_int2(ex, v[k]);
        ex += v[k];
    }
}

// ---- step 4: gather -----------------------------------------------------------------------------------
// List entry = {key = corner << 29 | source pixel, bilinear corner weight}: the fill pass has the footprint at
// hand, so the gather pass neither re-reads the flow nor recomputes it.

__device__ __forceinline__ void cswap(int& ka, float& wa, int& kb, float& wb) {
    const bool sw = ka > kb;
    const int tk = sw ? kb : ka; const float tw = sw ? wb : wa;
    kb = sw ? ka : kb; wb = sw ? wa : wb;
    ka = tk; wa = tw;
}

// Register path: up to 8 list entries per target, sorted with a fixed network; `nmax` is the warp-wide maximum
// list length, so the entry loop has a warp-uniform trip count (absent entries carry weight 0 and add +0).
// MODE soft is not bit-comparable with a CPU anyway (expf), so it merges the two weights, contracts to FMA and
// multiplies by the reciprocal of the denominator; sum / avg / linear keep the reference's operation order.
template <int MODE, int CCH>
__device__ __forceinline__ void gather_regs(const float* __restrict__ inb, const float* __restrict__ met, float* __restrict__ outb,
                                            int2* __restrict__ ent, int n, int nmax, int C, unsigned HW, int eps_mode, int dbg)
{
    constexpr int NE = 8;
    constexpr bool FAST = MODE == DRBA_SPLAT_SOFT;
    int key[NE]; float wc[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        key[e] = 0x7fffffff; wc[e] = 0.0f;
        if (e < n) { const int2 v = ent[e]; key[e] = v.x; wc[e] = __int_as_float(v.y); }
    }
    if (dbg & 2) {
    } else if (nmax <= 4) {
        cswap(key[0], wc[0], key[1], wc[1]); cswap(key[2], wc[2], key[3], wc[3]);
        cswap(key[0], wc[0], key[2], wc[2]); cswap(key[1], wc[1], key[3], wc[3]);
        cswap(key[1], wc[1], key[2], wc[2]);
    } else {
#pragma unroll
        for (int r = 0; r < NE; ++r)
#pragma unroll
            for (int i = (r & 1); i + 1 < NE; i += 2) cswap(key[i], wc[i], key[i + 1], wc[i + 1]);
    }
    unsigned src[NE]; float wg[NE];
    float den = 0.0f;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const bool on = e < n;
        src[e] = on ? (unsigned)(key[e] & 0x1fffffff) : 0u;
        wg[e] = 1.0f;
        if (MODE == DRBA_SPLAT_LINEAR) wg[e] = on ? met[src[e]] : 0.0f;
        if (MODE == DRBA_SPLAT_SOFT) wg[e] = on ? ((dbg & 4) ? 1.0f : expf(met[src[e]])) : 0.0f;
        if (MODE != DRBA_SPLAT_SUM) den += wg[e] * wc[e];
        if (FAST) wc[e] = wg[e] * wc[e];
    }
    float rden = 1.0f;
    if (MODE != DRBA_SPLAT_SUM) { den = splat_den(den, eps_mode); if (FAST) rden = 1.0f / den; }
    for (int c0 = 0; c0 < C; c0 += CCH) {
        float acc[CCH];
#pragma unroll
        for (int k = 0; k < CCH; ++k) acc[k] = 0.0f;
        const bool full = c0 + CCH <= C;
        // One plane pointer per channel of the chunk, no bounds predicate on the loads (channels past C re-read the last
        // plane: loaded, never stored).  The kernel used to execute 68 instructions per target and channel (ncu: 302 M
        // warp instructions, 57 % issue-bound): 17 per tap, most of them address arithmetic repeated under the
        // `c0 + k < C` predicate; this form needs 6.5 (SASS: 4 address + LDG + FFMA).
        const float* plk[CCH];
#pragma unroll
        for (int k = 0; k < CCH; ++k) plk[k] = inb + (size_t)(c0 + k < C ? c0 + k : C - 1) * HW;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (e < nmax) {        // warp-uniform
                const unsigned o = src[e];
                float t[CCH];
#pragma unroll
                for (int k = 0; k < CCH; ++k) t[k] = plk[k][o];
#pragma unroll
                for (int k = 0; k < CCH; ++k) {
                    if (FAST) {
                        acc[k] = fmaf(t[k], wc[e], acc[k]);
                    } else {
                        float v = t[k];
                        if (MODE == DRBA_SPLAT_LINEAR) v = v * wg[e];
                        acc[k] += v * wc[e];
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CCH; ++k)
            if (full || c0 + k < C)
                outb[(size_t)(c0 + k) * HW] = MODE == DRBA_SPLAT_SUM ? acc[k] : (FAST ? acc[k] * rden : acc[k] / den);
    }
}

template <int MODE, int CCH>
__global__ void __launch_bounds__(kGThreads)
splat_gather_kernel(const float* __restrict__ in, const float* __restrict__ metric,
                    float* __restrict__ out, int2* __restrict__ offcnt, int2* __restrict__ entries,
                    int N, int C, int H, int W, int eps_mode, int dbg, const int* __restrict__ done, int tiles_x, int tiles_y)
{
    // CTA = 32 x 8 targets (a warp = 32 x-adjacent targets of one row; the eight rows of a CTA share their source rows
    // in L1), grid = (tiles per image, N)
    const size_t HW = (size_t)H * W;
    const int img = blockIdx.y;
    const int gtx = (W + 31) >> 5;
    const int ty = blockIdx.x / gtx, tx = blockIdx.x - ty * gtx;
    const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
    bool live = x < W && y < H;
    const size_t rp = live ? (size_t)y * W + x : 0;
    const size_t p = (size_t)img * HW + rp;
    if (done && live) {
        // tiles already served by splat_gather_tile_kernel (64 x 16 targets, see below) -- except their targets with
        // more than 8 entries, which that kernel leaves to this one
        if (done[(img * tiles_y + (y >> 4)) * tiles_x + (x >> 6)] && offcnt[p].y <= 8) live = false;
    }
    int2 oc = make_int2(0, 0);
    if (live) oc = offcnt[p];
    const int n = oc.y;
    int2* ent = entries + (size_t)oc.x;
    const float* met = metric ? metric + (size_t)img * HW : nullptr;
    const float* inb = in + (size_t)img * C * HW;
    float* outb = out + (size_t)img * C * HW + rp;
    // warp-uniform bound of the register path (lanes with longer lists take the general path below)
    const int nmax = __reduce_max_sync(0xffffffffu, n <= 8 ? n : 0);
    if (!live) return;

    if (n <= 8) {
        gather_regs<MODE, CCH>(inb, met, outb, ent, n, nmax, C, (unsigned)HW, eps_mode, dbg);
    } else {
        // rare: many sources land on one target (folds / strong occlusion).  Sort the list in place in global
        // memory (it is private to this thread; heapsort keeps the worst case at n log n), then stream it once
        // per channel chunk.
        auto sift = [&](int root, int end) {
            const int2 v = ent[root];
            int i = root;
            for (;;) {
                int ch = 2 * i + 1;
                if (ch >= end) break;
                if (ch + 1 < end && ent[ch + 1].x > ent[ch].x) ++ch;
                if (ent[ch].x <= v.x) break;
                ent[i] = ent[ch];
                i = ch;
            }
            ent[i] = v;
        };
        for (int i = n / 2 - 1; i >= 0; --i) sift(i, n);
        for (int end = n - 1; end > 0; --end) {
            const int2 t = ent[0]; ent[0] = ent[end]; ent[end] = t;
            sift(0, end);
        }
        float den = 0.0f;
        if (MODE != DRBA_SPLAT_SUM) {
            for (int e = 0; e < n; ++e) {
                const int2 v = ent[e];
                const int src = v.x & 0x1fffffff;
                float wg = 1.0f;
                if (MODE == DRBA_SPLAT_LINEAR) wg = met[src];
                if (MODE == DRBA_SPLAT_SOFT) wg = expf(met[src]);
                den += wg * __int_as_float(v.y);
            }
            den = splat_den(den, eps_mode);
        }
        for (int c0 = 0; c0 < C; c0 += CCH) {
            float acc[CCH];
#pragma unroll
            for (int k = 0; k < CCH; ++k) acc[k] = 0.0f;
            for (int e = 0; e < n; ++e) {
                const int2 v = ent[e];
                const int src = v.x & 0x1fffffff;
                const float wc = __int_as_float(v.y);
                float wg = 1.0f;
                if (MODE == DRBA_SPLAT_LINEAR) wg = met[src];
                if (MODE == DRBA_SPLAT_SOFT) wg = expf(met[src]);
                const float* sp = inb + (size_t)c0 * HW + src;
#pragma unroll
                for (int k = 0; k < CCH; ++k) {
                    float t = (c0 + k < C) ? sp[(size_t)k * HW] : 0.0f;
                    if (MODE == DRBA_SPLAT_LINEAR || MODE == DRBA_SPLAT_SOFT) t = t * wg;
                    acc[k] += t * wc;
                }
            }
#pragma unroll
            for (int k = 0; k < CCH; ++k)
                if (c0 + k < C) outb[(size_t)(c0 + k) * HW] = MODE != DRBA_SPLAT_SUM ? acc[k] / den : acc[k];
        }
    }
}


// ---- step 4, tile flavour: sources staged in shared memory by TMA -------------------------------------------------
// The per-target gather above reads every source value through the L1 data stage: 4 entries x C channels scalar
// loads per target, two 128-byte wavefronts each because the sources of 32 adjacent targets are displaced by the
// flow -- ncu shows the L1 wavefront rate, not HBM, as its bound.  Here a CTA owns a 64 x 16 tile of
// targets.  The entries of a smooth flow all point into a compact source rectangle (tile + displacement spread), so:
//   * the CTA reduces the bounding box of its entries' sources; if it fits kBoxW x kBoxH,
//   * ONE TMA box load per 4-channel chunk ([4][kBoxH][kBoxW] fp32 out of the NCHW tensor viewed as a rank-3
//     [N*C][H][W] map, zero fill outside the image) lands in a 3-stage ring, the elected thread runs two chunks ahead,
//   * every thread (four targets) reads its <= 8 sorted entries' values from shared memory (adjacent lanes, adjacent
//     words: conflict-free), accumulates in the reference order and writes each output plane row-coalesced.
// Each source value crosses L2 -> SM once per tile instead of once per (entry, target); the arithmetic and its order
// are those of gather_regs, so results stay bit-identical to the CPU restatement for sum / avg / linear.
// Tiles whose sources do not fit the box (folds, large divergence) and targets with more than 8 entries take the
// per-target path above -- same results, old speed.
constexpr int kTileW = 64, kTileH = 16, kTileThreads = 256;     // four targets per thread: rows r, r + 4, r + 8, r + 12
constexpr int kTPT = kTileW * kTileH / kTileThreads;
// three box sizes (a tensor map each): a tile takes the smallest one its sources fit, so the L2 -> SM traffic follows
// the flow's actual spread (72 x 20: 1.4x the tile, 96 x 32: 3x); the ring holds as many stages as fit 96 KB
constexpr int kBoxC = 4, kNumBoxes = 3, kMaxBoxStages = 4;
constexpr int kRingBytes = 98304;
struct BoxMaps { CUtensorMap m[kNumBoxes]; };
__constant__ int c_box_w[kNumBoxes] = {72, 80, 96};
__constant__ int c_box_h[kNumBoxes] = {20, 24, 32};
static const int h_box_w[kNumBoxes] = {72, 80, 96}, h_box_h[kNumBoxes] = {20, 24, 32};

__device__ __forceinline__ uint32_t sg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sg_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sg_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = sg_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void sg_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(sg_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <int MODE, int BW, int BH>
__device__ __forceinline__ void staged_loop(const CUtensorMap* tmap, float* sg_box, uint64_t* full_bar,
                                            const unsigned (&src)[kTPT][8], const float (&wc)[kTPT][8], const float (&wg)[kTPT][8],
                                            const float (&den)[kTPT], const float (&rden)[kTPT], const int (&n)[kTPT],
                                            const bool (&live)[kTPT], const unsigned (&rp)[kTPT], float* outb, int nmax,
                                            int bx0, int by0, int plane0, int C, int W, size_t HW)
{
    constexpr int NE = 8;
    constexpr bool FAST = MODE == DRBA_SPLAT_SOFT;
    constexpr int plane = BH * BW, stage_floats = kBoxC * plane;
    constexpr int nst = kRingBytes / (stage_floats * 4) > kMaxBoxStages ? kMaxBoxStages : kRingBytes / (stage_floats * 4);
    const int tid = threadIdx.x;
    // two 16-bit byte offsets into the box per register (BH * BW * 4 < 65536)
    unsigned offp[kTPT][NE / 2];
#pragma unroll
    for (int j = 0; j < kTPT; ++j)
#pragma unroll
        for (int e = 0; e < NE; e += 2) {
            unsigned o2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int sy = (int)(src[j][e + h] / (unsigned)W), sx = (int)(src[j][e + h] - (unsigned)sy * (unsigned)W);
                o2[h] = (e + h < n[j] && n[j] <= NE) ? (unsigned)(((sy - by0) * BW + (sx - bx0)) * 4) : 0u;
            }
            offp[j][e / 2] = o2[0] | (o2[1] << 16);
        }
    const int nchunks = (C + kBoxC - 1) / kBoxC;
    if (tid == 0) {
        for (int s = 0; s < nst && s < nchunks; ++s) {
            sg_mbar_expect_tx(&full_bar[s], (uint32_t)(stage_floats * 4));
            sg_tma_load_3d(sg_smem_u32(sg_box + s * stage_floats), tmap, &full_bar[s], bx0, by0, plane0 + s * kBoxC);
        }
    }
    for (int it = 0; it < nchunks; ++it) {
        const int s = it % nst;
        sg_mbar_wait(&full_bar[s], (uint32_t)(it / nst) & 1u);
        const char* box = reinterpret_cast<const char*>(sg_box + s * stage_floats);
        const int c0 = it * kBoxC;
#pragma unroll
        for (int j = 0; j < kTPT; ++j) {
            if (!live[j] || n[j] > NE) continue;
            float acc[kBoxC];
#pragma unroll
            for (int k = 0; k < kBoxC; ++k) acc[k] = 0.0f;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                if (e < nmax) {          // warp-uniform; absent entries carry weight 0 and offset 0
                    const float* bp = reinterpret_cast<const float*>(box + ((e & 1) ? (offp[j][e >> 1] >> 16) : (offp[j][e >> 1] & 0xffffu)));
#pragma unroll
                    for (int k = 0; k < kBoxC; ++k) {
                        float v = bp[k * plane];
                        if (FAST) acc[k] = fmaf(v, wc[j][e], acc[k]);
                        else { if (MODE == DRBA_SPLAT_LINEAR) v = v * wg[j][e]; acc[k] += v * wc[j][e]; }
                    }
                }
            }
            float* o = outb + (size_t)c0 * HW + rp[j];
#pragma unroll
            for (int k = 0; k < kBoxC; ++k)
                if (c0 + k < C)
                    o[(size_t)k * HW] = MODE == DRBA_SPLAT_SUM ? acc[k] : (FAST ? acc[k] * rden[j] : acc[k] / den[j]);
        }
        __syncthreads();              // every thread is done with stage s
        if (tid == 0 && it + nst < nchunks) {
            sg_mbar_expect_tx(&full_bar[s], (uint32_t)(stage_floats * 4));
            sg_tma_load_3d(sg_smem_u32(sg_box + s * stage_floats), tmap, &full_bar[s], bx0, by0, plane0 + (it + nst) * kBoxC);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kTileThreads, 2)
splat_gather_tile_kernel(const __grid_constant__ BoxMaps maps, const float* __restrict__ in, const float* __restrict__ metric,
                         float* __restrict__ out, int2* __restrict__ offcnt, int2* __restrict__ entries, int* __restrict__ done,
                         int N, int C, int H, int W, int eps_mode, int tiles_x, int tiles_y)
{
    extern __shared__ __align__(128) float sg_box[];             // [stages][kBoxC][box h][box w]
    __shared__ uint64_t full_bar[kMaxBoxStages];
    __shared__ int s_min[2], s_max[2];
    constexpr int NE = 8;
    constexpr bool FAST = MODE == DRBA_SPLAT_SOFT;
    const int tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    int tile = blockIdx.x;
    const int img = tile / (tiles_x * tiles_y);
    tile -= img * tiles_x * tiles_y;
    const int ty0 = (tile / tiles_x) * kTileH, tx0 = (tile % tiles_x) * kTileW;
    const int x = tx0 + (tid & 63);
    const float* met = metric ? metric + (size_t)img * HW : nullptr;
    const float* inb = in + (size_t)img * C * HW;

    if (tid == 0) {
        for (int s = 0; s < kMaxBoxStages; ++s) sg_mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_min[0] = s_min[1] = 0x7fffffff; s_max[0] = s_max[1] = -0x7fffffff;
    }
    __syncthreads();

    // ---- my two targets: lists into registers, sorted into the reference's summation order -------------------------
    int n[kTPT]; bool live[kTPT]; unsigned rp[kTPT];
    unsigned src[kTPT][NE]; float wc[kTPT][NE], wg[kTPT][NE], den[kTPT], rden[kTPT];
    int2* ent[kTPT];
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -0x7fffffff, mxy = -0x7fffffff;
#pragma unroll
    for (int j = 0; j < kTPT; ++j) {
        const int y = ty0 + (tid >> 6) + (kTileH / kTPT) * j;
        live[j] = x < W && y < H;
        rp[j] = live[j] ? (unsigned)y * (unsigned)W + (unsigned)x : 0u;
        int2 oc = make_int2(0, 0);
        if (live[j]) oc = offcnt[(size_t)img * HW + rp[j]];
        n[j] = oc.y;
        ent[j] = entries + (size_t)oc.x;
        int key[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            key[e] = 0x7fffffff; wc[j][e] = 0.0f;
            if (e < n[j] && n[j] <= NE) { const int2 v = ent[j][e]; key[e] = v.x; wc[j][e] = __int_as_float(v.y); }
        }
#pragma unroll
        for (int r = 0; r < NE; ++r)
#pragma unroll
            for (int i = (r & 1); i + 1 < NE; i += 2) cswap(key[i], wc[j][i], key[i + 1], wc[j][i + 1]);
        den[j] = 0.0f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const bool on = e < n[j] && n[j] <= NE;
            src[j][e] = on ? (unsigned)(key[e] & 0x1fffffff) : 0u;
            wg[j][e] = 1.0f;
            if (MODE == DRBA_SPLAT_LINEAR) wg[j][e] = on ? met[src[j][e]] : 0.0f;
            if (MODE == DRBA_SPLAT_SOFT) wg[j][e] = on ? expf(met[src[j][e]]) : 0.0f;
            if (MODE != DRBA_SPLAT_SUM) den[j] += wg[j][e] * wc[j][e];
            if (FAST) wc[j][e] = wg[j][e] * wc[j][e];
            if (on) {
                const int sy = (int)(src[j][e] / (unsigned)W), sx = (int)(src[j][e] - (unsigned)sy * (unsigned)W);
                mnx = min(mnx, sx); mxx = max(mxx, sx); mny = min(mny, sy); mxy = max(mxy, sy);
            }
        }
        rden[j] = 1.0f;
        if (MODE != DRBA_SPLAT_SUM) { den[j] = splat_den(den[j], eps_mode); if (FAST) rden[j] = 1.0f / den[j]; }
    }
    // bounding box of the tile's sources
    int nloc = 0;
#pragma unroll
    for (int j = 0; j < kTPT; ++j) nloc = max(nloc, n[j] <= NE ? n[j] : 0);
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {      // shuffles, not redux.sync: ptxas turns the latter into uniform-datapath
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));      // CREDUX, which faulted (illegal instruction) here
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        nloc = max(nloc, __shfl_xor_sync(0xffffffffu, nloc, o));
    }
    const int nmax = nloc;                  // warp-wide maximum list length (<= NE)
    if ((tid & 31) == 0) {
        atomicMin(&s_min[0], mnx); atomicMin(&s_min[1], mny);
        atomicMax(&s_max[0], mxx); atomicMax(&s_max[1], mxy);
    }
    __syncthreads();
    // TMA moves 16-byte units: the box must start on a multiple of four floats in the innermost dimension (an
    // unaligned start coordinate raised "illegal instruction" on B200)
    const int bx0 = s_min[0] & ~3, by0 = s_min[1];
    const bool empty_tile = s_max[0] < s_min[0];
    int bsel = -1;
    if (!empty_tile)
        for (int b = kNumBoxes - 1; b >= 0; --b)
            if (s_max[0] - bx0 < c_box_w[b] && s_max[1] - by0 < c_box_h[b]) bsel = b;
    const bool fits = bsel >= 0;

    // sources too spread for the largest box (or none at all): the per-target kernel that runs next serves this tile
    if (!fits) return;
    if (tid == 0) done[blockIdx.x] = 1;

    // targets with more than 8 entries (folds) are left to the per-target kernel as well: served here they would
    // stall the whole tile behind one thread (measured: 39 % of the tiles of a gentle flow hold such targets)

    // ---- staged loop, instantiated per box size (compile-time plane pitch: a tap is LDS [reg + imm] + FFMA) ------------
    const int plane0 = img * C;
    float* outb = out + (size_t)img * C * HW;
    switch (bsel) {
        case 0: staged_loop<MODE, 72, 20>(&maps.m[0], sg_box, full_bar, src, wc, wg, den, rden, n, live, rp, outb, nmax, bx0, by0, plane0, C, W, HW); break;
        case 1: staged_loop<MODE, 80, 24>(&maps.m[1], sg_box, full_bar, src, wc, wg, den, rden, n, live, rp, outb, nmax, bx0, by0, plane0, C, W, HW); break;
        default: staged_loop<MODE, 96, 32>(&maps.m[2], sg_box, full_bar, src, wc, wg, den, rden, n, live, rp, outb, nmax, bx0, by0, plane0, C, W, HW); break;
    }
}

typedef CUresult (*SgEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static SgEncodeTiledFn sg_get_encode()
{
    static SgEncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<SgEncodeTiledFn>(sym);
    }
    return fn;
}

static GatherWs carve(void* ws, size_t total)
{
    GatherWs g;
    uint8_t* b = reinterpret_cast<uint8_t*>(ws);
    const size_t total4 = (total + 3) / 4 * 4;
    g.count = reinterpret_cast<int*>(b);                         b += total4 * sizeof(int);
    g.offcnt = reinterpret_cast<int2*>(b);                       b += total4 * sizeof(int2);
    g.entries = reinterpret_cast<int2*>(b);                      b += 4 * total4 * sizeof(int2);
    g.block_sums = reinterpret_cast<int*>(b);
    const size_t nblocks = (total + kScanBlock - 1) / kScanBlock;
    b += (nblocks + 4) * sizeof(int);
    g.done = reinterpret_cast<int*>(b);
    return g;
}

size_t splat_gather_workspace_bytes(int N, int H, int W)
{
    const size_t total = (size_t)N * H * W;
    const size_t total4 = (total + 3) / 4 * 4;
    const size_t nblocks = (total + kScanBlock - 1) / kScanBlock;
    const size_t ntiles = (size_t)N * ((W + 63) / 64) * ((H + 15) / 16);
    return total4 * 4 + total4 * 8 + 4 * total4 * 8 + (nblocks + 4) * 4 + (ntiles + 4) * 4;
}

int splat_gather_launch(const float* in, const float* flow, const float* metric, float* out,
                        int N, int C, int H, int W, int mode, int eps_mode, void* ws, cudaStream_t st, bool tile)
{
    const size_t total = (size_t)N * H * W;
    if (total >= (1ull << 31) || (size_t)H * W >= (1ull << 29) || (size_t)C * H * W >= (1ull << 32)) return DRBA_E_UNSUPPORTED;   // list keys: corner << 29 | pixel index
    const GatherWs g = carve(ws, total);
    static int env_dbg = -1;
    if (env_dbg < 0) { const char* e = getenv("DRBA_SPLAT_DBG"); env_dbg = e ? atoi(e) : 0; }
    const unsigned grid = cdiv(total, kGThreads);
    const unsigned nblocks = cdiv(total, kScanBlock);
    splat_list_kernel<false><<<grid, kGThreads, 0, st>>>(flow, g.count, g.offcnt, g.entries, N, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    scan_block_sums_kernel<<<nblocks, kScanThreads, 0, st>>>(g.count, g.block_sums, total);
    scan_sums_kernel<<<1, kScanThreads, 0, st>>>(g.block_sums, (int)nblocks);
    scan_write_kernel<<<nblocks, kScanThreads, 0, st>>>(g.count, g.block_sums, g.offcnt, total);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    splat_list_kernel<true><<<grid, kGThreads, 0, st>>>(flow, g.count, g.offcnt, g.entries, N, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    // tile flavour (sources staged by TMA) when the tensor can be described by a tiled map: W * 4 bytes a multiple of 16
    static int env_tile = -1;
    if (env_tile < 0) { const char* e = getenv("DRBA_SPLAT_TILE"); env_tile = e ? atoi(e) : 1; }
    const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
    bool tiled = false;
    if (env_tile && tile && W % 4 == 0 && (size_t)N * C < (1u << 30) && aligned16(in)) {
        SgEncodeTiledFn encode = sg_get_encode();
        BoxMaps maps;
        bool ok = encode != nullptr;
        for (int b = 0; ok && b < kNumBoxes; ++b) {
            const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * C};
            const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
            const cuuint32_t box[3] = {(cuuint32_t)h_box_w[b], (cuuint32_t)h_box_h[b], kBoxC};
            const cuuint32_t estr[3] = {1, 1, 1};
            ok = encode(&maps.m[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        if (ok) {
            static bool attr = false;
            const size_t smem = (size_t)kRingBytes;
            if (!attr) {
                cudaFuncSetAttribute(splat_gather_tile_kernel<DRBA_SPLAT_SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(splat_gather_tile_kernel<DRBA_SPLAT_AVG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(splat_gather_tile_kernel<DRBA_SPLAT_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(splat_gather_tile_kernel<DRBA_SPLAT_SOFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                attr = true;
            }
            const unsigned tgrid = (unsigned)(N * tiles_x * tiles_y);
#define GATHER_T(M) splat_gather_tile_kernel<M><<<tgrid, kTileThreads, smem, st>>>(maps, in, metric, out, g.offcnt, g.entries, g.done, N, C, H, W, eps_mode, tiles_x, tiles_y)
            switch (mode) {
                case DRBA_SPLAT_SUM: GATHER_T(DRBA_SPLAT_SUM); break;
                case DRBA_SPLAT_AVG: GATHER_T(DRBA_SPLAT_AVG); break;
                case DRBA_SPLAT_LINEAR: GATHER_T(DRBA_SPLAT_LINEAR); break;
                default: GATHER_T(DRBA_SPLAT_SOFT); break;
            }
#undef GATHER_T
            DRBA_RETURN_IF_LAUNCH_FAILED();
            tiled = true;
        }
    }
    // the per-target kernel serves whatever the tile kernel left (every tile when it did not run)
    const int* done = tiled ? g.done : nullptr;
#define GATHER(M) splat_gather_kernel<M, 8><<<dim3(((W + 31) / 32) * ((H + 7) / 8), N), kGThreads, 0, st>>>(in, metric, out, g.offcnt, g.entries, N, C, H, W, eps_mode, env_dbg, done, tiles_x, tiles_y)
    switch (mode) {
        case DRBA_SPLAT_SUM: GATHER(DRBA_SPLAT_SUM); break;
        case DRBA_SPLAT_AVG: GATHER(DRBA_SPLAT_AVG); break;
        case DRBA_SPLAT_LINEAR: GATHER(DRBA_SPLAT_LINEAR); break;
        default: GATHER(DRBA_SPLAT_SOFT); break;
    }
#undef GATHER
    DRBA_RETURN_IF_LAUNCH_FAILED();
    // workspace contract (all-zero on exit): the counters restored themselves in the fill pass; the
    // {offset, count} pairs and the lists are cleared by one bulk memset -- measured 13 us for 88 MB, while
    // zeroing from inside the gather kernel (a store chasing each load) tripled that kernel's time
    const size_t total4 = (total + 3) / 4 * 4;
    const cudaError_t me = cudaMemsetAsync(g.offcnt, 0, total4 * sizeof(int2) + 4 * total4 * sizeof(int2), st);
    if (me != cudaSuccess) return (int)me;
    if (tiled) {
        const cudaError_t me2 = cudaMemsetAsync(g.done, 0, (size_t)N * tiles_x * tiles_y * sizeof(int), st);
        if (me2 != cudaSuccess) return (int)me2;
    }
    return DRBA_OK;
}


// ======================================================================================================
// Pipeline flavour: lists are built ONCE per (flow, metric) and applied to several tensors -- GMFSS splats the
// half-resolution image (3 ch) and a 64 / 128 / 192-channel feature map with the same flow and metric
// (models/model_gmfss/GMFSS.py:96-115).  The sort pass merges the two weights (exp(metric) * bilinear corner)
// into one float per entry and stores the denominator per target, so the apply kernels are pure streams:
// no sort, no metric gather, no expf.  Features are NHWC fp16 (the conv engine's layout): one list entry is
// one contiguous channel vector, read 16 B per thread.
// ======================================================================================================

template <int MODE>
__global__ void __launch_bounds__(kGThreads)
splat_sortw_kernel(const float* __restrict__ metric, const int2* __restrict__ offcnt, int2* __restrict__ entries,
                   float* __restrict__ den_out, size_t HW)
{
    const size_t p = (size_t)blockIdx.x * kGThreads + threadIdx.x;
    if (p >= HW) return;
    const int2 oc = offcnt[p];
    const int n = oc.y;
    int2* ent = entries + (size_t)oc.x;
    auto weight = [&](int key) {
        const int src = key & 0x1fffffff;
        float wg = 1.0f;
        if (MODE == DRBA_SPLAT_LINEAR) wg = metric[src];
        if (MODE == DRBA_SPLAT_SOFT) wg = expf(metric[src]);
        return wg;
    };
    float den = 0.0f;
    if (n <= 8) {
        int key[8]; float wc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            key[e] = 0x7fffffff; wc[e] = 0.0f;
            if (e < n) { const int2 v = ent[e]; key[e] = v.x; wc[e] = __int_as_float(v.y); }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = (r & 1); i + 1 < 8; i += 2) cswap(key[i], wc[i], key[i + 1], wc[i + 1]);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e < n) {
                const float w = weight(key[e]) * wc[e];
                den += w;
                ent[e] = make_int2(key[e] & 0x1fffffff, __float_as_int(w));
            }
        }
    } else {
        auto sift = [&](int root, int end) {
            const int2 v = ent[root];
            int i = root;
            for (;;) {
                int ch = 2 * i + 1;
                if (ch >= end) break;
                if (ch + 1 < end && ent[ch + 1].x > ent[ch].x) ++ch;
                if (ent[ch].x <= v.x) break;
                ent[i] = ent[ch];
                i = ch;
            }
            ent[i] = v;
        };
        for (int i = n / 2 - 1; i >= 0; --i) sift(i, n);
        for (int end = n - 1; end > 0; --end) {
            const int2 t = ent[0]; ent[0] = ent[end]; ent[end] = t;
            sift(0, end);
        }
        for (int e = 0; e < n; ++e) {
            const int2 v = ent[e];
            const float w = weight(v.x) * __int_as_float(v.y);
            den += w;
            ent[e] = make_int2(v.x & 0x1fffffff, __float_as_int(w));
        }
    }
    den_out[p] = den;
}

// thread = (target pixel, group of 8 channels); consecutive lanes = consecutive channel groups of one target
__global__ void __launch_bounds__(kGThreads)
splat_apply_nhwc_f16_kernel(const int2* __restrict__ offcnt, const int2* __restrict__ entries, const float* __restrict__ den_in,
                            const __half* __restrict__ in, int in_cstride, __half* __restrict__ out, int out_cstride, int out_coffset,
                            int G8, size_t HW, int normalise, int eps_mode, int use_prelu, float slope)
{
    const size_t t = (size_t)blockIdx.x * kGThreads + threadIdx.x;
    if (t >= HW * (size_t)G8) return;
    const size_t p = t / G8;
    const int g = (int)(t - p * G8);
    const int2 oc = offcnt[p];
    const int2* ent = entries + (size_t)oc.x;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
    const __half* base = in + (size_t)g * 8;
    for (int e = 0; e < oc.y; ++e) {
        const int2 v = ent[e];
        const float w = __int_as_float(v.y);
        const uint4 q = *reinterpret_cast<const uint4*>(base + (size_t)v.x * in_cstride);
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            acc[2 * k] = fmaf(f.x, w, acc[2 * k]);
            acc[2 * k + 1] = fmaf(f.y, w, acc[2 * k + 1]);
        }
    }
    if (normalise) {
        const float r = 1.0f / splat_den(den_in[p], eps_mode);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] *= r;
    }
    if (use_prelu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = acc[k] > 0.0f ? acc[k] : slope * acc[k];
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
    *reinterpret_cast<uint4*>(out + p * out_cstride + out_coffset + (size_t)g * 8) = o;
}

__global__ void __launch_bounds__(kGThreads)
splat_apply_nchw_f32_kernel(const int2* __restrict__ offcnt, const int2* __restrict__ entries, const float* __restrict__ den_in,
                            const float* __restrict__ in, float* __restrict__ out, int C, size_t HW, int normalise, int eps_mode)
{
    const size_t p = (size_t)blockIdx.x * kGThreads + threadIdx.x;
    if (p >= HW) return;
    const int2 oc = offcnt[p];
    const int2* ent = entries + (size_t)oc.x;
    const float den = normalise ? splat_den(den_in[p], eps_mode) : 1.0f;
    for (int c = 0; c < C; ++c) {
        float acc = 0.0f;
        for (int e = 0; e < oc.y; ++e) {
            const int2 v = ent[e];
            acc += in[(size_t)c * HW + v.x] * __int_as_float(v.y);
        }
        out[(size_t)c * HW + p] = normalise ? acc / den : acc;
    }
}

}  // namespace drba

using namespace drba;

extern "C" {

size_t drba_splat_lists_workspace_bytes(int H, int W)
{
    if (H <= 0 || W <= 0) return 0;
    return splat_gather_workspace_bytes(1, H, W);
}

int drba_splat_lists_build(const float* flow, const float* metric, int mode, int H, int W, void* ws, size_t ws_bytes, void* stream)
{
    if (!flow || H <= 0 || W <= 0) return DRBA_E_ARG;
    if (mode < DRBA_SPLAT_SUM || mode > DRBA_SPLAT_SOFT) return DRBA_E_ARG;
    if ((mode == DRBA_SPLAT_LINEAR || mode == DRBA_SPLAT_SOFT) && !metric) return DRBA_E_ARG;
    const size_t total = (size_t)H * W;
    if (total >= (1ull << 29)) return DRBA_E_UNSUPPORTED;
    if (!ws || ws_bytes < splat_gather_workspace_bytes(1, H, W)) return DRBA_E_WORKSPACE;
    if (!aligned16(ws)) return DRBA_E_ALIGN;
    cudaStream_t st = as_stream(stream);
    const GatherWs g = carve(ws, total);
    const unsigned grid = cdiv(total, kGThreads);
    const unsigned nblocks = cdiv(total, kScanBlock);
    splat_list_kernel<false><<<grid, kGThreads, 0, st>>>(flow, g.count, g.offcnt, g.entries, 1, H, W);
    scan_block_sums_kernel<<<nblocks, kScanThreads, 0, st>>>(g.count, g.block_sums, total);
    scan_sums_kernel<<<1, kScanThreads, 0, st>>>(g.block_sums, (int)nblocks);
    scan_write_kernel<<<nblocks, kScanThreads, 0, st>>>(g.count, g.block_sums, g.offcnt, total);
    splat_list_kernel<true><<<grid, kGThreads, 0, st>>>(flow, g.count, g.offcnt, g.entries, 1, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    float* den = reinterpret_cast<float*>(g.count);     // the counters are back at zero: reuse them for the denominators
    switch (mode) {
        case DRBA_SPLAT_SUM: splat_sortw_kernel<DRBA_SPLAT_SUM><<<grid, kGThreads, 0, st>>>(metric, g.offcnt, g.entries, den, total); break;
        case DRBA_SPLAT_AVG: splat_sortw_kernel<DRBA_SPLAT_AVG><<<grid, kGThreads, 0, st>>>(metric, g.offcnt, g.entries, den, total); break;
        case DRBA_SPLAT_LINEAR: splat_sortw_kernel<DRBA_SPLAT_LINEAR><<<grid, kGThreads, 0, st>>>(metric, g.offcnt, g.entries, den, total); break;
        default: splat_sortw_kernel<DRBA_SPLAT_SOFT><<<grid, kGThreads, 0, st>>>(metric, g.offcnt, g.entries, den, total); break;
    }
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_splat_lists_apply_nhwc_f16(const void* ws, const void* in, int C, int in_cstride, void* out, int out_cstride, int out_coffset,
                                    int H, int W, int normalise, int eps_mode, int use_prelu, float slope, void* stream)
{
    if (!ws || !in || !out || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return DRBA_E_ARG;
    if (in_cstride < C || in_cstride % 8 != 0 || out_cstride < out_coffset + C || out_cstride % 8 != 0 || out_coffset % 8 != 0) return DRBA_E_ARG;
    if (!aligned16(in) || !aligned16(out)) return DRBA_E_ALIGN;
    const size_t total = (size_t)H * W;
    const GatherWs g = carve(const_cast<void*>(ws), total);
    const int G8 = C / 8;
    splat_apply_nhwc_f16_kernel<<<cdiv(total * G8, kGThreads), kGThreads, 0, as_stream(stream)>>>(
        g.offcnt, g.entries, reinterpret_cast<const float*>(g.count), (const __half*)in, in_cstride, (__half*)out, out_cstride, out_coffset,
        G8, total, normalise, eps_mode, use_prelu, slope);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_splat_lists_apply_nchw_f32(const void* ws, const float* in, float* out, int C, int H, int W, int normalise, int eps_mode, void* stream)
{
    if (!ws || !in || !out || H <= 0 || W <= 0 || C <= 0) return DRBA_E_ARG;
    const size_t total = (size_t)H * W;
    const GatherWs g = carve(const_cast<void*>(ws), total);
    splat_apply_nchw_f32_kernel<<<cdiv(total, kGThreads), kGThreads, 0, as_stream(stream)>>>(
        g.offcnt, g.entries, reinterpret_cast<const float*>(g.count), in, out, C, total, normalise, eps_mode);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_splat_lists_release(void* ws, int H, int W, void* stream)
{
    if (!ws || H <= 0 || W <= 0) return DRBA_E_ARG;
    const cudaError_t e = cudaMemsetAsync(ws, 0, splat_gather_workspace_bytes(1, H, W), as_stream(stream));
    return e == cudaSuccess ? DRBA_OK : (int)e;
}

}  // extern "C"
