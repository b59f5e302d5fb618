// Block-input assembly FUSED with the block's first convolution (conv0a: 3x3, stride 2, 64 -> N channels) for the
// two fine IFBlocks of IFNet 4.26-heavy (IFNet_HDv3.py:84-96 + :151-155; block3 at scale 2, block4 at scale 1).
//
// Why: at 1088 x 1920 the block-4 input (warp(img0), warp(img1), warp(f0), warp(f1), timestep, mask, feat, flow:
// 52 channels, packed to 64 fp16) is 267 MB per interpolated frame.  ifnet_assemble wrote it, conv0a read it back
// one kernel later: 1.07 GB of DRAM traffic per DRBA window that exists only because the two were separate kernels
// (ncu: 572 MB read by the block-4 program at 49 % L2 hit rate).  Here the assembled rows never
// leave the SM: producer warps compute them straight into the swizzled shared-memory operand tile that tcgen05.mma
// reads, and only conv0a's output (16 / 32 channels at half resolution, 17 / 33 MB) is written.
//
// Geometry (same as conv_tc.cu's stride-2 halo mode, so the MMA issue loop and its descriptors are the verified ones):
// an M tile is 16 x 8 output pixels; its 33 x 17 input pixels are stored as the four (row, column) parity planes of
// the [h/2][2][w/2][2] view, 17 rows x 9 cells x 128 B each, SWIZZLE_128B by absolute shared-memory address; tap
// (ky, kx) is a descriptor start offset into plane (ky != 1, kx != 1).  K = 64 channels = four K steps per tap,
// 36 MMAs (M = 128, N = 16 / 32) per tile.
//
// Roles (one persistent CTA per SM, 20 warps): warps 0..15 PRODUCE -- a lane pair per input pixel (NP = 1: the block
// works at full resolution) or a lane quad per input pixel (NP = 4: every input pixel is the 2 x 2 mean of
// full-resolution samples) evaluates all four parts of the block input with the arithmetic of ifnet_assemble_v2
// (ifnet_tc.cu) and stores four 16-byte pieces per lane; passes are dealt round-robin over the producer warps ACROSS
// tile boundaries (561 pixels per tile do not divide evenly), an mbarrier counts the passes of a tile.  Warps 16..19
// issue the MMAs (elected lane of warp 16) and drain the accumulator (bias, LeakyReLU, fp16 NHWC store); two operand
// stages and two accumulators keep producers, tensor core and epilogue overlapped.
#include <string.h>
#include "ifnet_common.cuh"
#include "tc_common.cuh"

namespace drba {

#ifndef DRBA_FU_PROD
#define DRBA_FU_PROD 16
#endif
constexpr int kFuProdWarps = DRBA_FU_PROD;
constexpr int kFuThreads = (kFuProdWarps + 4) * 32;
constexpr int kFuPlane = 20480;               // 17 * 9 * 128 B = 19584, padded to the 1024-byte swizzle period
constexpr int kFuStage = 4 * kFuPlane;
constexpr int kFuPixels = 561;                // 33 x 17 input pixels of a 16 x 8 output tile
constexpr int kFuMaxJobs = 2;

struct FusedParams {
    AssembleParams job[kFuMaxJobs];
    __half* out[kFuMaxJobs];                  // [oh][ow][N] fp16
    const __half* w;                          // [9][N][64] fp16, tap-major (ky * 3 + kx), K-major rows
    const float* bias;                        // [N]
    int njobs, oh, ow, tiles_x, tiles_per_img, total_tiles;
};

__device__ __forceinline__ void fu_cvt8(const uint4& a, float* v)
{
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        v[k * 2] = f.x; v[k * 2 + 1] = f.y;
    }
}

// the four taps of 8 channels of an NHWC-16 fp16 feature map: loads only (read-only path: the compiler may hoist them
// above the shared-memory stores of the previous pass), so that a lane has all its gathers in flight before it computes
__device__ __forceinline__ void fu_gather8(const __half* __restrict__ f, const WarpTap& t, uint4* q)
{
    q[0] = __ldg(reinterpret_cast<const uint4*>(f + (size_t)t.i00 * 16));
    q[1] = __ldg(reinterpret_cast<const uint4*>(f + (size_t)max(t.i01, 0) * 16));
    q[2] = __ldg(reinterpret_cast<const uint4*>(f + (size_t)max(t.i10, 0) * 16));
    q[3] = __ldg(reinterpret_cast<const uint4*>(f + (size_t)max(t.i11, 0) * 16));
}

// bilinear blend of the gathered taps (same expression order as ifnet_tc.cu: sample_feat8)
__device__ __forceinline__ void fu_blend8(const uint4* q, const WarpTap& t, float* out)
{
    float v0[8], v1[8], v2[8], v3[8];
    fu_cvt8(q[0], v0); fu_cvt8(q[1], v1); fu_cvt8(q[2], v2); fu_cvt8(q[3], v3);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float a = 0.0f + v0[c] * t.w00;
        if (t.i01 >= 0) a += v1[c] * t.w01;
        if (t.i10 >= 0) a += v2[c] * t.w10;
        if (t.i11 >= 0) a += v3[c] * t.w11;
        out[c] = a;
    }
}

__device__ __forceinline__ uint4 fu_pack8(const float* v)
{
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    return o;
}

// everything one lane contributes to one full-resolution sample position (x, y) whose flow `fl` the caller has already
// loaded (one pass ahead):
//   fa[8]  channels 8h..8h+7 of warp(f0), fb[8] of warp(f1)
//   im[3]  warp(img_h) (h = 0: img0 by flow[0:2], h = 1: img1 by flow[2:4])
//   ms[8]  h = 0: timestep, mask, feat 0..5;  h = 1: feat 6, 7, flow 0..3, 0, 0   (flow not yet scaled by 1 / s)
// Two load phases (features + image taps, then the previous lastconv output), each with every load issued before the
// first use: the kernel runs 16 producer warps per SM, so memory latency has to be covered by loads in flight per
// lane, not by occupancy.
__device__ __forceinline__ void fu_sample(const AssembleParams& p, int x, int y, int h, const float4& fl,
                                          float* fa, float* fb, float* im, float* ms)
{
    const size_t HW = (size_t)p.H * p.W;
    const WarpTap t0 = warp_tap(x, y, fl.x, fl.y, p.H, p.W);
    const WarpTap t1 = warp_tap(x, y, fl.z, fl.w, p.H, p.W);
    const WarpTap& ti = h ? t1 : t0;
    uint4 qa[4], qb[4];
    float iq[12];
    fu_gather8(reinterpret_cast<const __half*>(p.f0) + h * 8, t0, qa);
    fu_gather8(reinterpret_cast<const __half*>(p.f1) + h * 8, t1, qb);
    {
        const float* img = h ? p.img1 : p.img0;
        const int i01 = max(ti.i01, 0), i10 = max(ti.i10, 0), i11 = max(ti.i11, 0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* pl = img + (size_t)c * HW;
            iq[c * 4 + 0] = __ldg(pl + ti.i00); iq[c * 4 + 1] = __ldg(pl + i01);
            iq[c * 4 + 2] = __ldg(pl + i10); iq[c * 4 + 3] = __ldg(pl + i11);
        }
    }
    // x s_prev up-sampling of the previous lastconv output: lane h interpolates its channels 4 + 4h .. 11 + 4h
    // (two float4 per tap); the expression per channel is up_tmp's
    const Bilin b = bilin_up(y, x, p.prev);
    const float4* q00 = reinterpret_cast<const float4*>(p.prev.p + ((size_t)b.y0 * p.prev.w13 + b.x0) * p.prev.pitch + 4 + 4 * h);
    const float4* q01 = reinterpret_cast<const float4*>(p.prev.p + ((size_t)b.y0 * p.prev.w13 + b.x1) * p.prev.pitch + 4 + 4 * h);
    const float4* q10 = reinterpret_cast<const float4*>(p.prev.p + ((size_t)b.y1 * p.prev.w13 + b.x0) * p.prev.pitch + 4 + 4 * h);
    const float4* q11 = reinterpret_cast<const float4*>(p.prev.p + ((size_t)b.y1 * p.prev.w13 + b.x1) * p.prev.pitch + 4 + 4 * h);
    float4 ta[2], tb[2], tc[2], td[2];
#pragma unroll
    for (int k4 = 0; k4 < 2; ++k4) { ta[k4] = __ldg(q00 + k4); tb[k4] = __ldg(q01 + k4); tc[k4] = __ldg(q10 + k4); td[k4] = __ldg(q11 + k4); }
    const float ts = p.timestep ? __ldg(p.timestep + (size_t)y * p.W + x) : p.timestep_scalar;

    fu_blend8(qa, t0, fa);
    fu_blend8(qb, t1, fb);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // sample_plane's order
        float acc = 0.0f;
        acc += iq[c * 4 + 0] * ti.w00;
        if (ti.i01 >= 0) acc += iq[c * 4 + 1] * ti.w01;
        if (ti.i10 >= 0) acc += iq[c * 4 + 2] * ti.w10;
        if (ti.i11 >= 0) acc += iq[c * 4 + 3] * ti.w11;
        im[c] = acc;
    }
    float u[8];
#pragma unroll
    for (int k4 = 0; k4 < 2; ++k4) {
        const float4 a = ta[k4], bb = tb[k4], c = tc[k4], d = td[k4];
        u[k4 * 4 + 0] = b.hy * (b.hx * a.x + b.lx * bb.x) + b.ly * (b.hx * c.x + b.lx * d.x);
        u[k4 * 4 + 1] = b.hy * (b.hx * a.y + b.lx * bb.y) + b.ly * (b.hx * c.y + b.lx * d.y);
        u[k4 * 4 + 2] = b.hy * (b.hx * a.z + b.lx * bb.z) + b.ly * (b.hx * c.z + b.lx * d.z);
        u[k4 * 4 + 3] = b.hy * (b.hx * a.w + b.lx * bb.w) + b.ly * (b.hx * c.w + b.lx * d.w);
    }
    // h = 0: u = channels 4..11 -> [ts, mask(4), feat(5..10)];  h = 1: u = channels 8..15 -> [feat 11, 12, flow, 0, 0]
    ms[0] = h ? u[3] : ts;
    ms[1] = h ? u[4] : u[0];
    ms[2] = h ? fl.x : u[1];
    ms[3] = h ? fl.y : u[2];
    ms[4] = h ? fl.z : u[3];
    ms[5] = h ? fl.w : u[4];
    ms[6] = h ? 0.0f : u[5];
    ms[7] = h ? 0.0f : u[6];
}

// NP = 1: block scale 1 (input pixel = full-resolution pixel); NP = 4: block scale 2 (input pixel = 2 x 2 mean)
template <int NP, int N>
__global__ void __launch_bounds__(kFuThreads, 1)
ifnet_fused_conv0a_kernel(const __grid_constant__ FusedParams prm)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* const smem_g = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 1024-aligned (swizzle period)
    const uint32_t smem = smem_u32(smem_g);
    __shared__ uint64_t full_bar[2], done_bar[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_bias[N];

    constexpr int PXP = NP == 1 ? 16 : 8;                       // input pixels per producer pass
    constexpr int PASSES = (kFuPixels + PXP - 1) / PXP;         // per tile
    constexpr uint32_t kWBase = 2 * kFuStage;                   // weights behind the two stages
    constexpr uint32_t kWTap = N * 128;                         // one tap's [N][64] fp16 tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = prm.total_tiles > (int)blockIdx.x ? (prm.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], PASSES); mbar_init(&done_bar[s], 1); mbar_init(&tmem_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kFuProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x < N) s_bias[threadIdx.x] = prm.bias[threadIdx.x];
    // weights -> shared memory in the K-major SWIZZLE_128B image the B descriptors expect (row r of a tap tile at
    // r * 128, 16-byte piece c at (c ^ (r & 7)) * 16)
    for (int i = threadIdx.x; i < 9 * N * 8; i += kFuThreads) {
        const int c = i & 7, r = (i >> 3) % N, tap = (i >> 3) / N;
        const uint4 v = reinterpret_cast<const uint4*>(prm.w)[i];
        *reinterpret_cast<uint4*>(smem_g + kWBase + tap * kWTap + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < kFuProdWarps) {
        // ===== producers =====
        const int h = lane & 1;
        const int xpos = NP == 1 ? 0 : (lane >> 1) & 1;
        const int pl = NP == 1 ? lane >> 1 : lane >> 2;
        const int total_passes = my_tiles * PASSES;
        // one pass = PXP input pixels of one tile; everything a pass needs to know about its pixel
        struct Pass { int stage, use, img, par, row, cell, Xc, Yc; bool jvalid, inb; };
        auto decode = [&](int g, Pass& c) {
            const int i = g / PASSES, k = g - i * PASSES;
            c.stage = i & 1; c.use = i >> 1;
            const int idx = (int)blockIdx.x + i * (int)gridDim.x;
            c.img = idx / prm.tiles_per_img;
            const int m = idx - c.img * prm.tiles_per_img;
            const int ty = m / prm.tiles_x;
            const int oy0 = ty * 16, ox0 = (m - ty * prm.tiles_x) * 8;
            // pixel j of the tile's 33 x 17 input window, row-major: a pass covers x-adjacent pixels, so the gathers of
            // its lanes fall into the same lines (smooth flows).  Pixel (Y, X) lives in parity plane (Y & 1, X & 1) at
            // (row, cell) = (Y >> 1, X >> 1) relative to the plane origin one cell above / left of the tile.
            const int j = k * PXP + pl;
            c.jvalid = j < kFuPixels;
            const int jc = c.jvalid ? j : kFuPixels - 1;
            const int jy = jc / 17, jx = jc - jy * 17;
            const int Y = 2 * oy0 - 1 + jy, X = 2 * ox0 - 1 + jx;
            c.par = ((Y & 1) << 1) | (X & 1);
            c.row = ((Y - (Y & 1)) >> 1) - (oy0 - 1);
            c.cell = ((X - (X & 1)) >> 1) - (ox0 - 1);
            const AssembleParams& p = prm.job[c.img];
            c.inb = Y >= 0 && Y < p.h && X >= 0 && X < p.w;      // outside: the conv's zero padding
            c.Yc = min(max(Y, 0), p.h - 1); c.Xc = min(max(X, 0), p.w - 1);
        };
        // flow of the pass's sample positions, loaded one pass ahead (the gathers depend on it)
        auto load_flow = [&](const Pass& c, float4& f0, float4& f1) {
            const AssembleParams& p = prm.job[c.img];
            const float4* fl4 = reinterpret_cast<const float4*>(p.flow);
            if (NP == 1) { f0 = __ldg(fl4 + (size_t)c.Yc * p.W + c.Xc); f1 = f0; }
            else {
                f0 = __ldg(fl4 + (size_t)(2 * c.Yc) * p.W + 2 * c.Xc + xpos);
                f1 = __ldg(fl4 + (size_t)(2 * c.Yc + 1) * p.W + 2 * c.Xc + xpos);
            }
        };
        Pass cur;
        float4 fl0 = make_float4(0.f, 0.f, 0.f, 0.f), fl1 = fl0;
        if (warp < total_passes) { decode(warp, cur); load_flow(cur, fl0, fl1); }
        for (int g = warp; g < total_passes; g += kFuProdWarps) {
            Pass nxt = cur;
            float4 nf0 = fl0, nf1 = fl1;
            if (g + kFuProdWarps < total_passes) { decode(g + kFuProdWarps, nxt); load_flow(nxt, nf0, nf1); }
            const AssembleParams& p = prm.job[cur.img];
            const int stage = cur.stage, use = cur.use, par = cur.par, row = cur.row, cell = cur.cell;
            const bool jvalid = cur.jvalid, inb = cur.inb;
            const int Xc = cur.Xc, Yc = cur.Yc;

            float fa[8], fb[8], im[3], ms[8];
            if (NP == 1) {
                fu_sample(p, Xc, Yc, h, fl0, fa, fb, im, ms);
            } else {
                // 2 x 2 mean in the order bilinear resampling adds: ((p00 + p01) + (p10 + p11)) * 0.25; this lane holds
                // column xpos of both rows, the other column sits two lanes away
                float fa1[8], fb1[8], im1[3], ms1[8];
                fu_sample(p, 2 * Xc + xpos, 2 * Yc, h, fl0, fa, fb, im, ms);
                fu_sample(p, 2 * Xc + xpos, 2 * Yc + 1, h, fl1, fa1, fb1, im1, ms1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float a0 = fa[c] + __shfl_xor_sync(0xffffffffu, fa[c], 2), a1 = fa1[c] + __shfl_xor_sync(0xffffffffu, fa1[c], 2);
                    fa[c] = (a0 + a1) * 0.25f;
                    const float b0 = fb[c] + __shfl_xor_sync(0xffffffffu, fb[c], 2), b1 = fb1[c] + __shfl_xor_sync(0xffffffffu, fb1[c], 2);
                    fb[c] = (b0 + b1) * 0.25f;
                    const float m0 = ms[c] + __shfl_xor_sync(0xffffffffu, ms[c], 2), m1 = ms1[c] + __shfl_xor_sync(0xffffffffu, ms1[c], 2);
                    ms[c] = (m0 + m1) * 0.25f;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float a0 = im[c] + __shfl_xor_sync(0xffffffffu, im[c], 2), a1 = im1[c] + __shfl_xor_sync(0xffffffffu, im1[c], 2);
                    im[c] = (a0 + a1) * 0.25f;
                }
            }
            if (h) {
                const float inv = 1.0f / (float)p.s;          // IFNet_HDv3.py:87: interpolate(flow) * 1. / scale
#pragma unroll
                for (int c = 2; c < 6; ++c) ms[c] = ms[c] * 1.0f * inv;
            }
            // image piece: [img0 r g b | img1 r g b | 0 0] assembled in the even lane
            uint4 pc_img;
            {
                const float o0 = __shfl_xor_sync(0xffffffffu, im[0], 1), o1 = __shfl_xor_sync(0xffffffffu, im[1], 1), o2 = __shfl_xor_sync(0xffffffffu, im[2], 1);
                __half2* ph = reinterpret_cast<__half2*>(&pc_img);
                ph[0] = __floats2half2_rn(im[0], im[1]);
                ph[1] = __floats2half2_rn(im[2], o0);
                ph[2] = __floats2half2_rn(o1, o2);
                ph[3] = __floats2half2_rn(0.0f, 0.0f);
                if (h) pc_img = make_uint4(0u, 0u, 0u, 0u);
            }
            uint4 pc_a = fu_pack8(fa), pc_b = fu_pack8(fb), pc_m = fu_pack8(ms);
            if (!inb) { pc_a = pc_b = pc_img = pc_m = make_uint4(0u, 0u, 0u, 0u); }

            // the stage is free once the MMAs of the tile that used it two tiles ago have completed
            if (use > 0) mbar_wait(&done_bar[stage], (uint32_t)(use - 1) & 1u);
            if (jvalid && xpos == 0) {
                const int line = row * 9 + cell;
                uint8_t* dst = smem_g + stage * kFuStage + par * kFuPlane + line * 128;
                const int sw = line & 7;
                *reinterpret_cast<uint4*>(dst + (((0 + h) ^ sw) << 4)) = pc_a;
                *reinterpret_cast<uint4*>(dst + (((2 + h) ^ sw) << 4)) = pc_b;
                *reinterpret_cast<uint4*>(dst + (((4 + h) ^ sw) << 4)) = pc_img;
                *reinterpret_cast<uint4*>(dst + (((6 + h) ^ sw) << 4)) = pc_m;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
            cur = nxt; fl0 = nf0; fl1 = nf1;
        }
    } else {
        // ===== MMA issue (first of these warps) + epilogue (all four: TMEM lanes 32 * (warp % 4) ..) =====
        const int q = warp & 3;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t desc_b_hi = make_desc(0, 128);
        const uint64_t desc_a_hi = (desc_b_hi & ~(0x3FFFull << 32)) | ((uint64_t)(9u * 8u) << 32);     // rows of 8 cells, 9 cells apart
        for (int i = 0; i < my_tiles; ++i) {
            const int stage = i & 1, use = i >> 1;
            const uint32_t tmem_d = tmem_base + (uint32_t)stage * 32u;
            if (warp == kFuProdWarps) {
                mbar_wait(&full_bar[stage], (uint32_t)use & 1u);
                if (use > 0) mbar_wait(&tmem_empty[stage], (uint32_t)(use - 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a16 = ((smem + (uint32_t)stage * kFuStage) & 0x3FFFFu) >> 4;
                    const uint32_t b16 = ((smem + kWBase) & 0x3FFFFu) >> 4;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        const int par = (ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0);
                        const int offr = ky != 0 ? 1 : 0, offc = kx != 0 ? 1 : 0;
                        const uint64_t da = desc_a_hi | (uint64_t)(a16 + (uint32_t)(par * (kFuPlane >> 4) + (offr * 9 + offc) * 8));
                        const uint64_t db = desc_b_hi | (uint64_t)(b16 + (uint32_t)tap * (kWTap >> 4));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_f16(tmem_d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (uint32_t)(tap | ks));
                    }
                    tc_commit(&done_bar[stage]);
                }
                __syncwarp();
            }
            // ---- epilogue of tile i: lane = accumulator row = output pixel (oy0 + row / 8, ox0 + row % 8) ----
            const int idx = (int)blockIdx.x + i * (int)gridDim.x;
            const int img = idx / prm.tiles_per_img, m = idx - img * prm.tiles_per_img;
            const int ty = m / prm.tiles_x;
            const int r = q * 32 + lane;
            const int oy = ty * 16 + (r >> 3), ox = (m - ty * prm.tiles_x) * 8 + (r & 7);
            mbar_wait(&done_bar[stage], (uint32_t)use & 1u);
            tc_fence_after();
            uint32_t rr[N];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
            tc_ld16_nowait(taddr, rr);
            if (N == 32) tc_ld16_nowait(taddr + 16, rr + 16);
            tc_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[stage]);
            if (oy < prm.oh && ox < prm.ow) {
                uint4* o4 = reinterpret_cast<uint4*>(prm.out[img] + ((size_t)oy * prm.ow + ox) * N);
#pragma unroll
                for (int c8 = 0; c8 < N / 8; ++c8) {
                    float v[8];
#pragma unroll
                    for (int k2 = 0; k2 < 8; ++k2) {
                        const float a = __uint_as_float(rr[c8 * 8 + k2]) + s_bias[c8 * 8 + k2];
                        v[k2] = a > 0.0f ? a : 0.2f * a;
                    }
                    o4[c8] = fu_pack8(v);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kFuProdWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

template <typename K>
static int fu_launch(K kernel, const FusedParams& prm, size_t smem, cudaStream_t st)
{
    // (all four instantiations share this function's type, so no per-kernel static flag: the call is cheap)
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DRBA_E_UNSUPPORTED;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return DRBA_E_UNSUPPORTED;
    const int grid = prm.total_tiles < sms ? prm.total_tiles : sms;
    kernel<<<grid, kFuThreads, smem, st>>>(prm);
    return DRBA_OK;
}

}  // namespace drba

using namespace drba;

extern "C" int drba_ifnet_block_conv0a_f16(const drba_ifnet_block_input* jobs, int njobs, const void* w, const float* bias, int cout,
                                           int H, int W, int s, void* stream)
{
    if (!jobs || njobs < 1 || njobs > kFuMaxJobs || !w || !bias) return DRBA_E_ARG;
    if (cout != 16 && cout != 32) return DRBA_E_UNSUPPORTED;
    if (s != 1 && s != 2) return DRBA_E_UNSUPPORTED;
    if (H <= 0 || W <= 0 || H % (2 * s) != 0 || W % (2 * s) != 0) return DRBA_E_ARG;
    if (!aligned16(w)) return DRBA_E_ALIGN;
    FusedParams prm;
    memset(&prm, 0, sizeof(prm));
    const int h = H / s, wd = W / s;
    for (int k = 0; k < njobs; ++k) {
        const drba_ifnet_block_input& j = jobs[k];
        if (!j.img0 || !j.img1 || !j.f0 || !j.f1 || !j.flow || !j.tmp_prev || !j.out) return DRBA_E_ARG;
        if (j.s_prev <= 0 || H % j.s_prev != 0 || W % j.s_prev != 0) return DRBA_E_ARG;
        if (!aligned16(j.f0) || !aligned16(j.f1) || !aligned16(j.flow) || !aligned16(j.tmp_prev) || !aligned16(j.out)) return DRBA_E_ALIGN;
        AssembleParams& p = prm.job[k];
        p.img0 = j.img0; p.img1 = j.img1; p.f0 = j.f0; p.f1 = j.f1;
        p.timestep = j.timestep; p.timestep_scalar = j.timestep_scalar; p.flow = j.flow;
        p.prev.p = j.tmp_prev; p.prev.s = j.s_prev; p.prev.h13 = H / j.s_prev; p.prev.w13 = W / j.s_prev; p.prev.pitch = 16;
        p.nfterms = 0;
        p.fterm[0] = p.prev; p.fterm[1] = p.prev;
        p.out = nullptr; p.out_cstride = 64; p.H = H; p.W = W; p.s = s; p.h = h; p.w = wd;
        prm.out[k] = reinterpret_cast<__half*>(j.out);
    }
    prm.njobs = njobs; prm.w = reinterpret_cast<const __half*>(w); prm.bias = bias;
    prm.oh = h / 2; prm.ow = wd / 2;
    prm.tiles_x = (prm.ow + 7) / 8;
    prm.tiles_per_img = prm.tiles_x * ((prm.oh + 15) / 16);
    prm.total_tiles = njobs * prm.tiles_per_img;
    cudaStream_t st = as_stream(stream);
    int rc;
    if (cout == 16) {
        const size_t smem = 2 * kFuStage + 9 * 16 * 128 + 1024;
        rc = s == 1 ? fu_launch(ifnet_fused_conv0a_kernel<1, 16>, prm, smem, st) : fu_launch(ifnet_fused_conv0a_kernel<4, 16>, prm, smem, st);
    } else {
        const size_t smem = 2 * kFuStage + 9 * 32 * 128 + 1024;
        rc = s == 1 ? fu_launch(ifnet_fused_conv0a_kernel<1, 32>, prm, smem, st) : fu_launch(ifnet_fused_conv0a_kernel<4, 32>, prm, smem, st);
    }
    if (rc != DRBA_OK) return rc;
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}
