// Fused non-convolution stages of IFNet 4.26-heavy for sm_100a (batch 1).
//
// Reference: models/rife_426_heavy/IFNet_HDv3.py:126-177 (IFNet.forward), :84-96
// (IFBlock.forward), warplayer.py:8-22 (warp).  Per refinement block the reference runs
// 4 grid_samples (2 images + 2 feature maps at FULL resolution), a 48/52-channel torch.cat
// at full resolution, two F.interpolate down-samples, and after the block an F.interpolate
// up-sample of 13 channels plus slicing/scaling/adding -- ~4 GB of HBM traffic per frame at
// 1080p.  Here:
//
//  ifnet_assemble : builds the block's conv input DIRECTLY at the block's working resolution.
//                   For an integer down-scale s, bilinear align_corners=False resampling is
//                   exactly the mean of the 2x2 pixels at the centre of each s x s cell
//                   (SURVEY.md A.6), so the backward warps are evaluated only at those
//                   4/s^2 positions and the full-resolution warped tensors / concat never
//                   exist.  Output: channels [warp(img0) 3 | warp(img1) 3 | warp(f0) 16 |
//                   warp(f1) 16 | timestep 1 | mask 1 | feat 8 | flow/s 4] (IFNet_HDv3.py:151-
//                   155 + :87-88), NCHW fp32 (exact engine) or NHWC fp16 (tensor-core engine).
//                   Four warps per 32 output pixels, one role each (f0 | f1 | images | timestep,mask,feat,flow);
//                   mask/feat are bilinear taps of the previous block's small lastconv output
//                   (L2 resident) -- a full-resolution mask/feat tensor is never written.
//  ifnet_flow_accum: the only full-resolution state is the fp32 flow [H][W][4]:
//                   flow (+)= s * up(lastconv[0:4])  (IFNet_HDv3.py:91-93, :157), 32 B/px.
//  ifnet_blend    : last flow update + final warps + sigmoid blend (IFNet_HDv3.py:156-167) in
//                   one pass; the last block's flow/mask are never stored at full resolution.
#include "common.cuh"

namespace drba {

constexpr int kIfThreads = 128;

struct WarpTap {
    int i00, i01, i10, i11;   // element offsets y*W+x, or -1 when the tap is out of range
    float w00, w01, w10, w11;
};

// warplayer.py:8-22: border padding, align_corners=True, pixel coordinates
__device__ __forceinline__ WarpTap warp_tap(int x, int y, float fx, float fy, int H, int W)
{
    WarpTap t;
    float sx = fminf(fmaxf((float)x + fx, 0.0f), (float)(W - 1));
    float sy = fminf(fmaxf((float)y + fy, 0.0f), (float)(H - 1));
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float ax = sx - fx0, ay = sy - fy0;
    t.w00 = (1.0f - ax) * (1.0f - ay);
    t.w01 = ax * (1.0f - ay);
    t.w10 = (1.0f - ax) * ay;
    t.w11 = ax * ay;
    const bool vx1 = x0 + 1 < W, vy1 = y0 + 1 < H;
    t.i00 = y0 * W + x0;
    t.i01 = vx1 ? t.i00 + 1 : -1;
    t.i10 = vy1 ? t.i00 + W : -1;
    t.i11 = (vx1 && vy1) ? t.i00 + W + 1 : -1;
    return t;
}

__device__ __forceinline__ float sample_plane(const float* __restrict__ src, const WarpTap& t)
{
    float acc = 0.0f;
    acc += src[t.i00] * t.w00;
    if (t.i01 >= 0) acc += src[t.i01] * t.w01;
    if (t.i10 >= 0) acc += src[t.i10] * t.w10;
    if (t.i11 >= 0) acc += src[t.i11] * t.w11;
    return acc;
}

__device__ __forceinline__ void load16(const float* __restrict__ p, float* v)
{
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 a = q[i];
        v[i * 4 + 0] = a.x; v[i * 4 + 1] = a.y; v[i * 4 + 2] = a.z; v[i * 4 + 3] = a.w;
    }
}
__device__ __forceinline__ void load16(const __half* __restrict__ p, float* v)
{
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint4 a = q[i];
        const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            v[i * 8 + k * 2] = f.x; v[i * 8 + k * 2 + 1] = f.y;
        }
    }
}

// bilinear sample of a 16-channel NHWC feature map, same accumulation order as sample_plane
template <typename FT>
__device__ __forceinline__ void sample_feat16(const FT* __restrict__ f, const WarpTap& t, float* out)
{
    float v[16];
    load16(f + (size_t)t.i00 * 16, v);
#pragma unroll
    for (int c = 0; c < 16; ++c) out[c] = 0.0f + v[c] * t.w00;
    if (t.i01 >= 0) { load16(f + (size_t)t.i01 * 16, v);
#pragma unroll
        for (int c = 0; c < 16; ++c) out[c] += v[c] * t.w01; }
    if (t.i10 >= 0) { load16(f + (size_t)t.i10 * 16, v);
#pragma unroll
        for (int c = 0; c < 16; ++c) out[c] += v[c] * t.w10; }
    if (t.i11 >= 0) { load16(f + (size_t)t.i11 * 16, v);
#pragma unroll
        for (int c = 0; c < 16; ++c) out[c] += v[c] * t.w11; }
}

// ---- lastconv output access -----------------------------------------------------------
// TMP_LAYOUT 0: ConvTranspose output, NCHW fp32 [52][h13/2][w13/2] (PixelShuffle(2) is index
//               arithmetic here: IFNet_HDv3.py:81)
// TMP_LAYOUT 1: pixel-shuffled NHWC fp32 [h13][w13][16] (13 used), written by the tensor-core
//               engine's lastconv epilogue
struct Tmp13 {
    const float* p; int h13, w13, s;   // 13-channel map at 1/s of the full resolution
};

struct Bilin {           // F.interpolate(scale_factor=s, bilinear, align_corners=False) source taps
    int y0, y1, x0, x1;
    float ly, hy, lx, hx;
};

__device__ __forceinline__ Bilin bilin_up(int y, int x, const Tmp13& t)
{
    Bilin b;
    const float r = 1.0f / (float)t.s;                 // ATen: ratio = 1 / scale_factor
    float sy = r * ((float)y + 0.5f) - 0.5f, sx = r * ((float)x + 0.5f) - 0.5f;
    if (sy < 0.0f) sy = 0.0f;
    if (sx < 0.0f) sx = 0.0f;
    b.y0 = (int)sy; b.x0 = (int)sx;
    b.y1 = b.y0 + (b.y0 < t.h13 - 1 ? 1 : 0);
    b.x1 = b.x0 + (b.x0 < t.w13 - 1 ? 1 : 0);
    b.ly = sy - (float)b.y0; b.hy = 1.0f - b.ly;
    b.lx = sx - (float)b.x0; b.hx = 1.0f - b.lx;
    return b;
}

template <int TMP_LAYOUT, int C0, int NC>
__device__ __forceinline__ void load_tmp(const Tmp13& t, int yy, int xx, float* v)
{
    if (TMP_LAYOUT == 0) {
        const int h2 = t.h13 >> 1, w2 = t.w13 >> 1;
        const size_t base = (size_t)(yy >> 1) * w2 + (xx >> 1);
        const int sub = (yy & 1) * 2 + (xx & 1);
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] = t.p[(size_t)((C0 + c) * 4 + sub) * h2 * w2 + base];
    } else {
        const float* q = t.p + ((size_t)yy * t.w13 + xx) * 16;
        if (C0 % 4 == 0) {
            const float4* q4 = reinterpret_cast<const float4*>(q + C0);
#pragma unroll
            for (int c4 = 0; c4 < (NC + 3) / 4; ++c4) {
                const float4 a = q4[c4];
                if (c4 * 4 + 0 < NC) v[c4 * 4 + 0] = a.x;
                if (c4 * 4 + 1 < NC) v[c4 * 4 + 1] = a.y;
                if (c4 * 4 + 2 < NC) v[c4 * 4 + 2] = a.z;
                if (c4 * 4 + 3 < NC) v[c4 * 4 + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) v[c] = q[C0 + c];
        }
    }
}

// channels [C0, C0+NC) of the x s bilinear up-sampling of the 13-channel map at (y, x)
template <int TMP_LAYOUT, int C0, int NC>
__device__ __forceinline__ void up_tmp(const Tmp13& t, int y, int x, float* o)
{
    const Bilin b = bilin_up(y, x, t);
    float a[NC], bb[NC], c[NC], d[NC];
    load_tmp<TMP_LAYOUT, C0, NC>(t, b.y0, b.x0, a);
    load_tmp<TMP_LAYOUT, C0, NC>(t, b.y0, b.x1, bb);
    load_tmp<TMP_LAYOUT, C0, NC>(t, b.y1, b.x0, c);
    load_tmp<TMP_LAYOUT, C0, NC>(t, b.y1, b.x1, d);
#pragma unroll
    for (int k = 0; k < NC; ++k) o[k] = b.hy * (b.hx * a[k] + b.lx * bb[k]) + b.ly * (b.hx * c[k] + b.lx * d[k]);
}

// ---- flow state: flow (+)= s * up(tmp[0:4])  (IFNet_HDv3.py:91-93, :157) ---------------------
template <int TMP_LAYOUT>
__global__ void __launch_bounds__(256)
ifnet_flow_accum_kernel(const Tmp13 t, float* __restrict__ flow, float* __restrict__ planar, int accumulate, int H, int W)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= H * W) return;
    const int y = idx / W, x = idx - y * W;
    float o[4];
    up_tmp<TMP_LAYOUT, 0, 4>(t, y, x, o);
    const float fs = (float)t.s;
    float4 f = make_float4(o[0] * fs, o[1] * fs, o[2] * fs, o[3] * fs);
    float4* dst = reinterpret_cast<float4*>(flow) + idx;
    if (accumulate) {
        const float4 old = *dst;
        f.x = old.x + f.x; f.y = old.y + f.y; f.z = old.z + f.z; f.w = old.w + f.w;
    }
    if (flow) *dst = f;
    if (planar) {
        const size_t HW = (size_t)H * W;
        planar[idx] = f.x; planar[HW + idx] = f.y; planar[2 * HW + idx] = f.z; planar[3 * HW + idx] = f.w;
    }
}

// ---- block input assembly ---------------------------------------------------------------
struct AssembleParams {
    const float* img0; const float* img1;   // [3][H][W] fp32
    const void* f0; const void* f1;         // [H][W][16] FT
    const float* timestep; float timestep_scalar;   // [H][W] or NULL -> scalar
    const float* flow;                      // [H][W][4] fp32 state, or NULL (first block: no warp, 39 channels)
    Tmp13 prev;                             // previous block's lastconv output (mask/feat source)
    void* out; int out_cstride;             // NHWC: channels allocated per pixel
    int H, W, s, h, w;                      // full size, integer scale, h = H/s, w = W/s
};

// Mean of the (up to) 4 sampled positions in the order bilinear resampling adds them:
// 0.5*(0.5*a + 0.5*b) + 0.5*(0.5*c + 0.5*d) == ((a + b) + (c + d)) * 0.25 exactly in fp32.
// a0 collects positions 0,1 (top row), a1 positions 2,3 (bottom row).
template <int K>
__device__ __forceinline__ void acc_pos(float& a0, float& a1, float val)
{
    if (K == 0) a0 = val;
    else if (K == 1) a0 += val;
    else if (K == 2) a1 = val;
    else a1 += val;
}
__device__ __forceinline__ float mean_pos(float a0, float a1, int np) { return np == 1 ? a0 : (a0 + a1) * 0.25f; }

// Output channel of (role, k).  Reference order (IFNet_HDv3.py:151-155 + :88):
//   [warp(img0) 0-2 | warp(img1) 3-5 | warp(f0) 6-21 | warp(f1) 22-37 | timestep 38 | mask 39 | feat 40-47 | flow 48-51]
// NHWC fp16 (tensor-core engine) uses a packed order of four aligned 16-channel segments, one per
// role, so each thread stores 32 contiguous bytes; the host permutes the conv weights to match:
//   [f0 16 | f1 16 | img0 3, img1 3, (first block: timestep), pad | timestep, mask, feat 8, flow 4, pad 2]
__device__ __forceinline__ int ref_channel(int role, int k)
{
    return role == 0 ? 6 + k : (role == 1 ? 22 + k : (role == 2 ? (k < 6 ? k : 38) : 38 + k));
}

template <bool NHWC_HALF>
__device__ __forceinline__ void store16(const AssembleParams& p, int Y, int X, int role, const float* v, int n)
{
    if (NHWC_HALF) {
        uint4 o[2];
        __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int k = 0; k < 8; ++k) oh[k] = __floats2half2_rn(2 * k < n ? v[2 * k] : 0.0f, 2 * k + 1 < n ? v[2 * k + 1] : 0.0f);
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + ((size_t)Y * p.w + X) * p.out_cstride + role * 16);
        dst[0] = o[0]; dst[1] = o[1];
    } else {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)Y * p.w + X;
        const size_t hw = (size_t)p.h * p.w;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < n) o[(size_t)ref_channel(role, k) * hw] = v[k];
    }
}

template <typename FT, int K>
__device__ __forceinline__ void feat_pos(const AssembleParams& p, const FT* f, int role, int x, int y, const float4& fl,
                                         float* a0, float* a1)
{
    // f0 follows flow[:2], f1 follows flow[2:4] (IFNet_HDv3.py:152-153)
    const WarpTap t = warp_tap(x, y, role == 0 ? fl.x : fl.z, role == 0 ? fl.y : fl.w, p.H, p.W);
    float q[16];
    sample_feat16<FT>(f, t, q);
#pragma unroll
    for (int c = 0; c < 16; ++c) acc_pos<K>(a0[c], a1[c], q[c]);
}

template <int K>
__device__ __forceinline__ void img_pos(const AssembleParams& p, int x, int y, const float4& fl, size_t HW, float* a0, float* a1)
{
    const WarpTap t0 = warp_tap(x, y, fl.x, fl.y, p.H, p.W);
    const WarpTap t1 = warp_tap(x, y, fl.z, fl.w, p.H, p.W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        acc_pos<K>(a0[c], a1[c], sample_plane(p.img0 + (size_t)c * HW, t0));
        acc_pos<K>(a0[3 + c], a1[3 + c], sample_plane(p.img1 + (size_t)c * HW, t1));
    }
    acc_pos<K>(a0[6], a1[6], p.timestep ? p.timestep[(size_t)y * p.W + x] : p.timestep_scalar);
}

template <int TMP_LAYOUT, int K>
__device__ __forceinline__ void misc_pos(const AssembleParams& p, int x, int y, const float4& fl, float* a0, float* a1)
{
    acc_pos<K>(a0[0], a1[0], p.timestep ? p.timestep[(size_t)y * p.W + x] : p.timestep_scalar);
    float mf[9];
    up_tmp<TMP_LAYOUT, 4, 9>(p.prev, y, x, mf);     // mask, feat: x s_prev up-sampling of the previous lastconv output
#pragma unroll
    for (int c = 0; c < 9; ++c) acc_pos<K>(a0[1 + c], a1[1 + c], mf[c]);
    acc_pos<K>(a0[10], a1[10], fl.x);
    acc_pos<K>(a0[11], a1[11], fl.y);
    acc_pos<K>(a0[12], a1[12], fl.z);
    acc_pos<K>(a0[13], a1[13], fl.w);
}

template <typename FT, bool NHWC_HALF, int TMP_LAYOUT>
__global__ void __launch_bounds__(kIfThreads)
ifnet_assemble_kernel(const AssembleParams p)
{
    // one warp per role (no divergence inside a warp), one lane per output pixel
    const int role = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + (threadIdx.x & 31);
    if (idx >= p.h * p.w) return;
    const bool first = p.flow == nullptr;
    if (first && role == 3) return;
    const int Y = idx / p.w, X = idx - Y * p.w;
    const int np = p.s == 1 ? 1 : 4;
    const int off = p.s == 1 ? 0 : p.s / 2 - 1;
    const size_t HW = (size_t)p.H * p.W;
    const int bx = p.s * X + off, by = p.s * Y + off;
    const float4* flow4 = reinterpret_cast<const float4*>(p.flow);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    float a0[16], a1[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { a0[c] = 0.0f; a1[c] = 0.0f; }
    int n = 16;
#define DRBA_FOR_POS(CALL)                                                                         \
    {   { const int x = bx, y = by; const float4 fl = first ? zero4 : flow4[(size_t)y * p.W + x]; constexpr int K = 0; CALL; } \
        if (np == 4) {                                                                             \
            { const int x = bx + 1, y = by; const float4 fl = first ? zero4 : flow4[(size_t)y * p.W + x]; constexpr int K = 1; CALL; } \
            { const int x = bx, y = by + 1; const float4 fl = first ? zero4 : flow4[(size_t)y * p.W + x]; constexpr int K = 2; CALL; } \
            { const int x = bx + 1, y = by + 1; const float4 fl = first ? zero4 : flow4[(size_t)y * p.W + x]; constexpr int K = 3; CALL; } \
        } }
    if (role < 2) {
        const FT* f = reinterpret_cast<const FT*>(role == 0 ? p.f0 : p.f1);
        DRBA_FOR_POS((feat_pos<FT, K>(p, f, role, x, y, fl, a0, a1)))
    } else if (role == 2) {
        DRBA_FOR_POS((img_pos<K>(p, x, y, fl, HW, a0, a1)))
        n = first ? 7 : 6;     // the first block carries the timestep here (no mask/feat/flow role)
    } else {
        DRBA_FOR_POS((misc_pos<TMP_LAYOUT, K>(p, x, y, fl, a0, a1)))
        n = 14;
    }
#undef DRBA_FOR_POS
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = mean_pos(a0[c], a1[c], np);
    if (role == 3) {
        const float inv = 1.0f / (float)p.s;   // IFNet_HDv3.py:87: interpolate(flow) * 1. / scale
#pragma unroll
        for (int c = 10; c < 14; ++c) v[c] = v[c] * 1.0f * inv;
    }
    store16<NHWC_HALF>(p, Y, X, role, v, n);
}

// ---- block input assembly, tensor-core engine (NHWC fp16 in / out), L1-friendly lane mapping ----------
// The one-lane-per-pixel kernel above is bound by the L1 data stage (ncu: l1tex 68-86 %): with a lane
// stride of 32-128 B every LDG.128 / STG.128 touches 8-32 lines.  Here adjacent lanes cover adjacent
// bytes: a lane pair shares a pixel's 32 B of features, x-adjacent sample positions sit in adjacent
// lanes (the 2x2 mean becomes two shuffles, same summation order as acc_pos / mean_pos), and the
// 128 B/pixel output row is staged in shared memory and stored as full lines.  Results are bit-identical
// to ifnet_assemble_kernel<__half, true, 1>.
__device__ __forceinline__ void load8(const __half* __restrict__ p, float* v)
{
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        v[k * 2] = f.x; v[k * 2 + 1] = f.y;
    }
}

__device__ __forceinline__ void sample_feat8(const __half* __restrict__ f, const WarpTap& t, float* out)
{
    float v0[8], v1[8], v2[8], v3[8];
    load8(f + (size_t)t.i00 * 16, v0);
    load8(f + (size_t)max(t.i01, 0) * 16, v1);
    load8(f + (size_t)max(t.i10, 0) * 16, v2);
    load8(f + (size_t)max(t.i11, 0) * 16, v3);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float a = 0.0f + v0[c] * t.w00;
        if (t.i01 >= 0) a += v1[c] * t.w01;
        if (t.i10 >= 0) a += v2[c] * t.w10;
        if (t.i11 >= 0) a += v3[c] * t.w11;
        out[c] = a;
    }
}

__device__ __forceinline__ uint4 pack8(const float* v)
{
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    return o;
}

constexpr int kAsmTile = 32;   // output pixels per CTA

// 16 B chunk c of tile pixel px; rotated by 2*px so that pair / quad writers and the 8-lane row readers
// spread over the banks
__device__ __forceinline__ int tile_slot(int px, int c) { return px * 8 + ((c + 2 * px) & 7); }

template <int NP>
__global__ void __launch_bounds__(kIfThreads)
ifnet_assemble_v2_kernel(const AssembleParams p)
{
    __shared__ __align__(16) uint4 tile[kAsmTile * 8];
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hw = p.h * p.w;
    const int base = blockIdx.x * kAsmTile;
    const int off = NP == 1 ? 0 : p.s / 2 - 1;
    const size_t HW = (size_t)p.H * p.W;
    const float4* flow4 = reinterpret_cast<const float4*>(p.flow);
    constexpr int PXW = NP == 1 ? 16 : 8;      // output pixels per warp pass

    if (role < 2) {
        // lane = [pixel | x position | channel half]
        const __half* f = reinterpret_cast<const __half*>(role == 0 ? p.f0 : p.f1) + (lane & 1) * 8;
        const int xpos = NP == 1 ? 0 : (lane >> 1) & 1;
        const int pl = NP == 1 ? lane >> 1 : lane >> 2;
#pragma unroll 1
        for (int pass = 0; pass < kAsmTile / PXW; ++pass) {
            const int px = pass * PXW + pl;
            const int idx = min(base + px, hw - 1);
            const int Y = idx / p.w, X = idx - Y * p.w;
            const int x = p.s * X + off + xpos, y = p.s * Y + off;
            float r0[8];
            {
                const float4 fl = flow4[(size_t)y * p.W + x];
                const WarpTap t = warp_tap(x, y, role == 0 ? fl.x : fl.z, role == 0 ? fl.y : fl.w, p.H, p.W);
                sample_feat8(f, t, r0);
            }
            if (NP == 4) {
                float r1[8];
                const float4 fl = flow4[(size_t)(y + 1) * p.W + x];
                const WarpTap t = warp_tap(x, y + 1, role == 0 ? fl.x : fl.z, role == 0 ? fl.y : fl.w, p.H, p.W);
                sample_feat8(f, t, r1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float a0 = r0[c] + __shfl_xor_sync(0xffffffffu, r0[c], 2);
                    const float a1 = r1[c] + __shfl_xor_sync(0xffffffffu, r1[c], 2);
                    r0[c] = (a0 + a1) * 0.25f;
                }
            }
            if (xpos == 0) tile[tile_slot(px, role * 2 + (lane & 1))] = pack8(r0);
        }
    } else if (role == 2) {
        // lane = [image | pixel | x position]
        const int sel = lane >> 4;
        const int xpos = NP == 1 ? 0 : lane & 1;
        const int pl = NP == 1 ? lane & 15 : (lane >> 1) & 7;
        const float* img = sel ? p.img1 : p.img0;
#pragma unroll 1
        for (int pass = 0; pass < kAsmTile / PXW; ++pass) {
            const int px = pass * PXW + pl;
            const int idx = min(base + px, hw - 1);
            const int Y = idx / p.w, X = idx - Y * p.w;
            const int x = p.s * X + off + xpos, y = p.s * Y + off;
            float r0[3];
            {
                const float4 fl = flow4[(size_t)y * p.W + x];
                const WarpTap t = warp_tap(x, y, sel ? fl.z : fl.x, sel ? fl.w : fl.y, p.H, p.W);
#pragma unroll
                for (int c = 0; c < 3; ++c) r0[c] = sample_plane(img + (size_t)c * HW, t);
            }
            if (NP == 4) {
                float r1[3];
                const float4 fl = flow4[(size_t)(y + 1) * p.W + x];
                const WarpTap t = warp_tap(x, y + 1, sel ? fl.z : fl.x, sel ? fl.w : fl.y, p.H, p.W);
#pragma unroll
                for (int c = 0; c < 3; ++c) r1[c] = sample_plane(img + (size_t)c * HW, t);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float a0 = r0[c] + __shfl_xor_sync(0xffffffffu, r0[c], 1);
                    const float a1 = r1[c] + __shfl_xor_sync(0xffffffffu, r1[c], 1);
                    r0[c] = (a0 + a1) * 0.25f;
                }
            }
            if (xpos == 0) {
                __half* dst = reinterpret_cast<__half*>(&tile[tile_slot(px, 4)]);
#pragma unroll
                for (int c = 0; c < 3; ++c) dst[sel * 3 + c] = __float2half_rn(r0[c]);
                if (sel) *reinterpret_cast<uint32_t*>(dst + 6) = 0u;
                else tile[tile_slot(px, 5)] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    } else {
        // lane = [pixel | sample position]; timestep, mask, feat (x s_prev up-sampling of the previous
        // lastconv output) and flow / s
        constexpr int PXM = NP == 1 ? 32 : 8;
        const int K = NP == 1 ? 0 : lane & 3;
        const int pl = NP == 1 ? lane : lane >> 2;
        const float inv = 1.0f / (float)p.s;   // IFNet_HDv3.py:87: interpolate(flow) * 1. / scale
#pragma unroll 1
        for (int pass = 0; pass < kAsmTile / PXM; ++pass) {
            const int px = pass * PXM + pl;
            const int idx = min(base + px, hw - 1);
            const int Y = idx / p.w, X = idx - Y * p.w;
            const int x = p.s * X + off + (K & 1), y = p.s * Y + off + (K >> 1);
            float v[16];
            const float4 fl = flow4[(size_t)y * p.W + x];
            v[0] = p.timestep ? p.timestep[(size_t)y * p.W + x] : p.timestep_scalar;
            up_tmp<1, 4, 9>(p.prev, y, x, v + 1);
            v[10] = fl.x; v[11] = fl.y; v[12] = fl.z; v[13] = fl.w;
            if (NP == 4) {
#pragma unroll
                for (int c = 0; c < 14; ++c) {
                    const float a = v[c] + __shfl_xor_sync(0xffffffffu, v[c], 1);
                    v[c] = (a + __shfl_xor_sync(0xffffffffu, a, 2)) * 0.25f;
                }
            }
#pragma unroll
            for (int c = 10; c < 14; ++c) v[c] = v[c] * 1.0f * inv;
            v[14] = 0.0f; v[15] = 0.0f;
            if (K == 0) {
                tile[tile_slot(px, 6)] = pack8(v);
                tile[tile_slot(px, 7)] = pack8(v + 8);
            }
        }
    }
    __syncthreads();
    uint4* out4 = reinterpret_cast<uint4*>(p.out);
#pragma unroll
    for (int k = 0; k < kAsmTile * 8 / kIfThreads; ++k) {
        const int i = threadIdx.x + k * kIfThreads;
        const int px = i >> 3, c = i & 7;
        if (base + px < hw) out4[(size_t)(base + px) * 8 + c] = tile[tile_slot(px, c)];
    }
}

// IFNet_HDv3.py:156-167: last flow update, warp both images, blend with sigmoid(mask)
template <int TMP_LAYOUT>
__global__ void __launch_bounds__(kIfThreads)
ifnet_blend_kernel(const float* __restrict__ img0, const float* __restrict__ img1,
                   const float* __restrict__ flow, const Tmp13 t, float* __restrict__ out, int H, int W)
{
    const int idx = blockIdx.x * kIfThreads + threadIdx.x;
    if (idx >= H * W) return;
    const int y = idx / W, x = idx - y * W;
    float o[5];
    up_tmp<TMP_LAYOUT, 0, 5>(t, y, x, o);
    const float fs = (float)t.s;
    float4 fl = make_float4(o[0] * fs, o[1] * fs, o[2] * fs, o[3] * fs);
    if (flow) {
        const float4 old = reinterpret_cast<const float4*>(flow)[idx];
        fl.x = old.x + fl.x; fl.y = old.y + fl.y; fl.z = old.z + fl.z; fl.w = old.w + fl.w;
    }
    const WarpTap t0 = warp_tap(x, y, fl.x, fl.y, H, W);
    const WarpTap t1 = warp_tap(x, y, fl.z, fl.w, H, W);
    const float m = 1.0f / (1.0f + expf(-o[4]));
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float w0 = sample_plane(img0 + (size_t)c * HW, t0);
        const float w1 = sample_plane(img1 + (size_t)c * HW, t1);
        out[(size_t)c * HW + idx] = w0 * m + w1 * (1.0f - m);
    }
}

}  // namespace drba

using namespace drba;

// read on every call so that a test can switch kernels inside one process
static bool env_flag(const char* name)
{
    const char* e = getenv(name);
    return e && e[0] != '\0' && e[0] != '0';
}

static int tmp_ok(const float* tmp, int layout, int H, int W, int s)
{
    if (!tmp || (layout != 0 && layout != 1) || s <= 0 || H % s != 0 || W % s != 0) return DRBA_E_ARG;
    if (layout == 0 && ((H / s) % 2 != 0 || (W / s) % 2 != 0)) return DRBA_E_ARG;
    if (layout == 1 && !aligned16(tmp)) return DRBA_E_ALIGN;
    return DRBA_OK;
}

extern "C" {

int drba_ifnet_assemble(const float* img0, const float* img1, const void* f0, const void* f1, int feat_dtype,
                        const float* timestep, float timestep_scalar,
                        const float* flow, const float* tmp_prev, int tmp_layout, int s_prev,
                        void* out, int out_dtype, int out_cstride, int H, int W, int s, void* stream)
{
    if (H <= 0 || W <= 0 || s <= 0 || (s & (s - 1)) != 0 || H % s != 0 || W % s != 0) return DRBA_E_ARG;
    if (!img0 || !img1 || !f0 || !f1 || !out) return DRBA_E_ARG;
    if (feat_dtype != DRBA_F32 && feat_dtype != DRBA_F16) return DRBA_E_ARG;
    if (out_dtype != DRBA_F32 && out_dtype != DRBA_F16) return DRBA_E_ARG;
    if (out_dtype == DRBA_F16 && (out_cstride < (flow ? 64 : 48) || out_cstride % 8 != 0)) return DRBA_E_ARG;
    if (!aligned16(f0) || !aligned16(f1) || !aligned16(out) || (flow && !aligned16(flow))) return DRBA_E_ALIGN;
    if (flow) {
        const int rc = tmp_ok(tmp_prev, tmp_layout, H, W, s_prev);
        if (rc != DRBA_OK) return rc;
    }
    AssembleParams p;
    p.img0 = img0; p.img1 = img1; p.f0 = f0; p.f1 = f1;
    p.timestep = timestep; p.timestep_scalar = timestep_scalar; p.flow = flow;
    p.prev.p = tmp_prev; p.prev.s = flow ? s_prev : 1; p.prev.h13 = flow ? H / s_prev : 1; p.prev.w13 = flow ? W / s_prev : 1;
    p.out = out; p.out_cstride = out_cstride; p.H = H; p.W = W; p.s = s; p.h = H / s; p.w = W / s;
    const unsigned grid = cdiv((size_t)p.h * p.w, 32);
    cudaStream_t st = as_stream(stream);
    const int key = (feat_dtype == DRBA_F16 ? 4 : 0) | (out_dtype == DRBA_F16 ? 2 : 0) | (tmp_layout == 1 ? 1 : 0);
    if (key == 7 && flow && out_cstride == 64 && !env_flag("DRBA_ASSEMBLE_V1")) {
        const unsigned g2 = cdiv((size_t)p.h * p.w, kAsmTile);
        if (s == 1) ifnet_assemble_v2_kernel<1><<<g2, kIfThreads, 0, st>>>(p);
        else ifnet_assemble_v2_kernel<4><<<g2, kIfThreads, 0, st>>>(p);
        DRBA_RETURN_IF_LAUNCH_FAILED();
        return DRBA_OK;
    }
    switch (key) {
        case 0: ifnet_assemble_kernel<float, false, 0><<<grid, kIfThreads, 0, st>>>(p); break;
        case 1: ifnet_assemble_kernel<float, false, 1><<<grid, kIfThreads, 0, st>>>(p); break;
        case 2: ifnet_assemble_kernel<float, true, 0><<<grid, kIfThreads, 0, st>>>(p); break;
        case 3: ifnet_assemble_kernel<float, true, 1><<<grid, kIfThreads, 0, st>>>(p); break;
        case 4: ifnet_assemble_kernel<__half, false, 0><<<grid, kIfThreads, 0, st>>>(p); break;
        case 5: ifnet_assemble_kernel<__half, false, 1><<<grid, kIfThreads, 0, st>>>(p); break;
        case 6: ifnet_assemble_kernel<__half, true, 0><<<grid, kIfThreads, 0, st>>>(p); break;
        default: ifnet_assemble_kernel<__half, true, 1><<<grid, kIfThreads, 0, st>>>(p); break;
    }
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_ifnet_flow_accum(const float* tmp, int tmp_layout, int s, float* flow, float* planar, int accumulate,
                          int H, int W, void* stream)
{
    if (H <= 0 || W <= 0 || (!flow && !planar) || (accumulate && !flow)) return DRBA_E_ARG;
    const int rc = tmp_ok(tmp, tmp_layout, H, W, s);
    if (rc != DRBA_OK) return rc;
    if (flow && !aligned16(flow)) return DRBA_E_ALIGN;
    Tmp13 t; t.p = tmp; t.s = s; t.h13 = H / s; t.w13 = W / s;
    const unsigned grid = cdiv((size_t)H * W, 256);
    if (tmp_layout == 0) ifnet_flow_accum_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(t, flow, planar, accumulate, H, W);
    else ifnet_flow_accum_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(t, flow, planar, accumulate, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_ifnet_blend(const float* img0, const float* img1, const float* flow, const float* tmp, int tmp_layout, int s,
                     float* out, int H, int W, void* stream)
{
    if (H <= 0 || W <= 0 || !img0 || !img1 || !out) return DRBA_E_ARG;
    const int rc = tmp_ok(tmp, tmp_layout, H, W, s);
    if (rc != DRBA_OK) return rc;
    if (flow && !aligned16(flow)) return DRBA_E_ALIGN;
    Tmp13 t; t.p = tmp; t.s = s; t.h13 = H / s; t.w13 = W / s;
    const unsigned grid = cdiv((size_t)H * W, kIfThreads);
    if (tmp_layout == 0) ifnet_blend_kernel<0><<<grid, kIfThreads, 0, as_stream(stream)>>>(img0, img1, flow, t, out, H, W);
    else ifnet_blend_kernel<1><<<grid, kIfThreads, 0, as_stream(stream)>>>(img0, img1, flow, t, out, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
