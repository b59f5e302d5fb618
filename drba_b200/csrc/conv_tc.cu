// tcgen05 implicit-GEMM convolution for sm_100a (fp16 operands, fp32 accumulation in TMEM).
//
// Serves every conv of IFNet 4.26-heavy's refinement blocks
// (models/rife_426_heavy/IFNet_HDv3.py:62-96: conv0 (two 3x3 stride-2), 8 x ResConv (3x3),
// lastconv ConvTranspose2d(c, 52, 4, 2, 1) + PixelShuffle(2)) in the reference's own GPU
// precision (fp16 under torch.autocast, models/rife.py:26,78; accumulation fp32).
//
// GEMM view: M = output pixels, N = output channels, K = taps x Cin.
//  * activations are NHWC fp16; a CTA owns a 16 x 8 pixel patch (M tile = 128 rows) and an
//    N tile of <= 128 channels; its accumulator is a [128 lanes x ntile columns] fp32 block
//    of tensor memory;
//  * per K step (one tap, Kc input channels) ONE tiled TMA load brings the shifted
//    [8][16][Kc] window of the input into shared memory as a K-major, hardware-swizzled
//    [128 x Kc] A tile -- zero padding comes from TMA's out-of-bounds fill, so there is no
//    im2col buffer and no halo logic; a second TMA load brings the [ntile x Kc] weight tile.
//    Stride-2 convs view the input as [H/2][2][W/2][2][C] (rank-5 tensor map) so that a tap
//    is still a dense box;
//  * warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma issuer (+ TMEM alloc),
//    warps 2-5 = epilogue (tcgen05.ld -> bias / residual / LeakyReLU -> fp16 NHWC store, or
//    the lastconv pixel-shuffle scatter); smem stages are recycled through full/empty
//    mbarriers, tcgen05.commit releases a stage when the MMAs reading it retire;
//  * ConvTranspose(4,2,1) is four 2x2-tap phase convs (grid.z = phase), N = 52 padded to 64;
//    the epilogue writes PixelShuffle(2)'d NHWC fp32 [4h][4w][16] directly.
#include <cuda.h>
#include "common.cuh"

namespace drba {

constexpr int kTcThreads = 192;
constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kMaxNTile = 128;
constexpr int kMaxTapsTc = 9;
constexpr int kMaxGroups = 4;
constexpr int kTmemCols = 128;

struct TcParams {
    int OH, OW;             // output grid (per group)
    int Cin, Kc, kchunks;   // input channels (padded), channels per K step, Cin / Kc
    int S;                  // input stride (1 or 2)
    int T;                  // taps per group
    int dy[kMaxGroups][kMaxTapsTc], dx[kMaxGroups][kMaxTapsTc];
    int ntile, nsplits;     // N tile and number of N tiles per group
    int cout_pad;           // weight rows per (group, tap)
    int cout;               // real output channels per group
    int swz_bytes;          // 32 / 64 / 128
    int epilogue;           // 0: NHWC fp16 (+res, act); 1: lastconv pixel shuffle -> fp32 [4*OH][4*OW][16]
    int act;                // 0 none, 1 LeakyReLU(0.2)
    const float* bias;      // [G][cout_pad]
    const __half* res;      // NHWC fp16, same geometry as out (or NULL)
    void* out;
    int out_cstride;        // channels per pixel in `out` (epilogue 0)
    int os;                 // epilogue 0: output placement stride; group g = phase (py, px) lands at (os*oy + py, os*ox + px)
    int tiles_x;
    int tile_w, tile_h;     // pixel patch of one CTA: tile_w * tile_h = 128
    int stages;             // smem pipeline depth (<= kMaxStages)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 |
// [46,48) version = 1 | [61,64) layout type (SWIZZLE_128B = 2, 64B = 4, 32B = 6)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int swz_bytes) {
    const uint64_t layout = swz_bytes == 128 ? 2ull : (swz_bytes == 64 ? 4ull : 6ull);
    const uint64_t sbo = (uint64_t)(8 * swz_bytes) >> 4;   // 8 rows of one swizzle span
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ float lrelu02(float v) { return v > 0.0f ? v : 0.2f * v; }

__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: stages x (A tile | B tile), 1024-byte aligned, then barriers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_bytes = kTileM * p.Kc * 2;
    const int b_bytes = p.ntile * p.Kc * 2;
    const int b_off = a_bytes;                                   // a_bytes is a multiple of 1024 (Kc >= 16 -> 4096)
    const int stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], accum_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    const int oy0 = ty * p.tile_h, ox0 = tx * p.tile_w;
    const int kStages = p.stages;
    const int nsplit = blockIdx.y, g = blockIdx.z;
    const int KI = p.T * p.kchunks;

    // let the next kernel in the stream start its prologue while this one runs (PDL)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // everything above overlapped the previous kernel's tail; its results are visible after this
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int it = 0; it < KI; ++it) {
                const int s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + b_bytes));
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
                const int dy = p.dy[g][tap], dx = p.dx[g][tap];
                // input coordinate = S * o + d  ->  (parity, cell) in the [H/S][S][W/S][S][C] view
                int pary = 0, parx = 0, cy = oy0 + dy, cx = ox0 + dx;
                if (p.S == 2) {
                    pary = dy & 1; parx = dx & 1;
                    cy = oy0 + ((dy - pary) >> 1); cx = ox0 + ((dx - parx) >> 1);
                }
                uint8_t* sa = smem + (size_t)s * stage_bytes;
                tma_load_5d(sa, &tmap_a, &full_bar[s], kc * p.Kc, parx, cx, pary, cy);
                tma_load_2d(sa + b_off, &tmap_b, &full_bar[s], kc * p.Kc,
                            (g * p.T + tap) * p.cout_pad + nsplit * p.ntile);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, K-major both,
            // N >> 3 at [17,23), M >> 4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.ntile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const int ksteps = p.Kc >> 4;
            for (int it = 0; it < KI; ++it) {
                const int s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t da = make_desc(sa, p.swz_bytes), db = make_desc(sa + b_off, p.swz_bytes);
                for (int k = 0; k < ksteps; ++k) {
                    // advance 16 fp16 = 32 bytes along K inside the swizzle span: +2 in the (addr >> 4) field
                    tc_mma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
                }
                tc_commit(&empty_bar[s]);
            }
            tc_commit(&accum_bar);
        }
    } else {
        // ===== epilogue: warps 2..5; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int oy = oy0 + row / p.tile_w, ox = ox0 + row % p.tile_w;
        const bool valid = oy < p.OH && ox < p.OW;
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        const float* bias = p.bias + (size_t)g * p.cout_pad + nsplit * p.ntile;
        if (p.epilogue == 0) {
            const size_t pix = p.os == 1 ? (size_t)oy * p.OW + ox
                                         : (size_t)(oy * p.os + (g >> 1)) * (p.OW * p.os) + (ox * p.os + (g & 1));
            __half* out = reinterpret_cast<__half*>(p.out) + pix * p.out_cstride + nsplit * p.ntile;
            const __half* res = p.res ? p.res + pix * p.out_cstride + nsplit * p.ntile : nullptr;
            for (int c0 = 0; c0 < p.ntile; c0 += 16) {
                float v[16];
                tc_ld16(taddr + c0, v);
                if (!valid) continue;
                if (nsplit * p.ntile + c0 >= p.cout) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(bias + c0 + i);
                if (res) {
                    const uint4* r4 = reinterpret_cast<const uint4*>(res + c0);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint4 rr = r4[h];
                        const __half2* hh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 f = __half22float2(hh[k]);
                            v[h * 8 + k * 2] += f.x; v[h * 8 + k * 2 + 1] += f.y;
                        }
                    }
                }
                if (p.act == 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = lrelu02(v[i]);
                }
                uint4 o[2];
                __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
                for (int k = 0; k < 8; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
                uint4* o4 = reinterpret_cast<uint4*>(out + c0);
                o4[0] = o[0]; o4[1] = o[1];
            }
        } else {
            // lastconv: group g = phase (py, px); channel co = c13*4 + i*2 + j lands at
            // out[4*oy + 2*py + i][4*ox + 2*px + j][c13]  (ConvTranspose phase + PixelShuffle(2))
            const int py = g >> 1, px = g & 1;
            const int OW4 = p.OW * 4;
            float* out = reinterpret_cast<float*>(p.out);
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float v[16];
                tc_ld16(taddr + c0, v);
                if (!valid) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(bias + c0 + i);
#pragma unroll
                for (int ij = 0; ij < 4; ++ij) {
                    const int yy = 4 * oy + 2 * py + (ij >> 1), xx = 4 * ox + 2 * px + (ij & 1);
                    float4 o = make_float4(v[0 + ij], v[4 + ij], v[8 + ij], v[12 + ij]);
                    *reinterpret_cast<float4*>(out + ((size_t)yy * OW4 + xx) * 16 + (c0 >> 2)) = o;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
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

}  // namespace drba

using namespace drba;

extern "C" {

int drba_conv_tc_f16(const void* in, int H, int W, int Cin,
                     const void* w, const float* bias, int G, int T, const int* dy, const int* dx,
                     int cout_pad, int cout, int S, int OH, int OW,
                     int epilogue, int act, const void* res, void* out, int out_cstride, int out_os, void* stream)
{
    if (!in || !w || !bias || !out || !dy || !dx) return DRBA_E_ARG;
    if (H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || Cin <= 0 || Cin % 16 != 0) return DRBA_E_ARG;
    if (G < 1 || G > kMaxGroups || T < 1 || T > kMaxTapsTc) return DRBA_E_ARG;
    if (S != 1 && S != 2) return DRBA_E_ARG;
    if (S == 2 && (H % 2 != 0 || W % 2 != 0)) return DRBA_E_ARG;
    if (cout_pad <= 0 || cout_pad % 16 != 0 || cout <= 0 || cout > cout_pad) return DRBA_E_ARG;
    if (epilogue != 0 && epilogue != 1) return DRBA_E_ARG;
    if (epilogue == 1 && (cout_pad != 64 || cout != 52 || G != 4)) return DRBA_E_ARG;
    if (out_os != 1 && out_os != 2) return DRBA_E_ARG;
    if (epilogue == 0 && ((out_os == 1 && G != 1) || (out_os == 2 && G != 4))) return DRBA_E_ARG;
    if (epilogue == 0 && (out_cstride < cout_pad || out_cstride % 8 != 0)) return DRBA_E_ARG;
    if (!aligned16(in) || !aligned16(w) || !aligned16(out) || (res && !aligned16(res))) return DRBA_E_ALIGN;
    EncodeTiledFn encode = get_encode();
    if (!encode) return DRBA_E_UNSUPPORTED;

    TcParams p;
    p.OH = OH; p.OW = OW; p.Cin = Cin;
    p.Kc = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
    p.kchunks = Cin / p.Kc;
    p.swz_bytes = p.Kc * 2;
    p.S = S; p.T = T;
    for (int gi = 0; gi < G; ++gi)
        for (int t = 0; t < T; ++t) { p.dy[gi][t] = dy[gi * T + t]; p.dx[gi][t] = dx[gi * T + t]; }
    // pixel patch of a CTA: the 128-pixel rectangle that covers the output with the fewest tiles
    int best_tiles = 1 << 30;
    p.tile_w = 16; p.tile_h = 8;
    const int shapes[5][2] = {{16, 8}, {32, 4}, {8, 16}, {64, 2}, {128, 1}};
    for (int i = 0; i < 5; ++i) {
        const int nt = ((OW + shapes[i][0] - 1) / shapes[i][0]) * ((OH + shapes[i][1] - 1) / shapes[i][1]);
        if (nt < best_tiles) { best_tiles = nt; p.tile_w = shapes[i][0]; p.tile_h = shapes[i][1]; }
    }
    p.tiles_x = (OW + p.tile_w - 1) / p.tile_w;
    const int tiles_y = (OH + p.tile_h - 1) / p.tile_h;
    const int tiles = p.tiles_x * tiles_y;
    // N tile: <= 128 columns, a multiple of 16 that divides cout_pad; small layers are split further
    // so that more SMs stream the K loop in parallel
    int ntile = 0;
    for (int cand = cout_pad < kMaxNTile ? cout_pad : kMaxNTile; cand >= 16; cand -= 16)
        if (cout_pad % cand == 0) { ntile = cand; break; }
    if (!ntile) return DRBA_E_UNSUPPORTED;
    if (epilogue == 0) {
        while (ntile >= 64 && ntile % 32 == 0 && tiles * G * (cout_pad / ntile) * 2 <= kNumSMs) ntile /= 2;
    }
    p.ntile = ntile; p.nsplits = cout_pad / ntile; p.cout_pad = cout_pad; p.cout = cout;
    p.epilogue = epilogue; p.act = act; p.bias = bias; p.res = (const __half*)res; p.out = out;
    p.out_cstride = out_cstride; p.os = out_os;

    // A: input viewed as [H/S][S][W/S][S][C], innermost first
    CUtensorMap ta, tb;
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)S, (cuuint64_t)(W / S), (cuuint64_t)S, (cuuint64_t)(H / S)};
        const cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)S * Cin * 2, (cuuint64_t)W * Cin * 2,
                                       (cuuint64_t)S * W * Cin * 2};
        const cuuint32_t box[5] = {(cuuint32_t)p.Kc, 1, (cuuint32_t)p.tile_w, 1, (cuuint32_t)p.tile_h};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const CUresult r = encode(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(in), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(p.swz_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return DRBA_E_UNSUPPORTED;
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)G * T * cout_pad};
        const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        const cuuint32_t box[2] = {(cuuint32_t)p.Kc, (cuuint32_t)ntile};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(p.swz_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return DRBA_E_UNSUPPORTED;
    }
    const int a_bytes = kTileM * p.Kc * 2, b_bytes = ntile * p.Kc * 2;
    const int stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
    // pipeline depth: a TMA round trip is ~1-2 us, so a CTA needs many stages in flight.  Few CTAs
    // (latency-bound small layers): take the whole SM; many CTAs: leave room for 2-3 CTAs per SM.
    const int ctas = tiles * p.nsplits * G;
    const int budget = ctas <= kNumSMs ? 220 * 1024 : (ctas <= 2 * kNumSMs ? 108 * 1024 : 72 * 1024);
    int stages = budget / stage_bytes;
    const int KI = T * p.kchunks;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > KI) stages = KI;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles, p.nsplits, G);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap our prologue with the previous kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tc_kernel, ta, tb, p);
    if (le != cudaSuccess) return (int)le;
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
