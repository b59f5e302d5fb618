// Device helpers shared by ifnet.cu (exact-rounding build, -fmad=false) and ifnet_tc.cu (tensor-core engine,
// FMA contraction allowed): backward-warp taps, lastconv-output access, block-input parameters.
#pragma once
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
    int pitch;                          // TMP_LAYOUT 1: floats per pixel (16; 8 = channels 0..7 only, tmp_layout 2 at the C ABI)
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
        const float* q = t.p + ((size_t)yy * t.w13 + xx) * t.pitch;
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

// ---- block input assembly ---------------------------------------------------------------
struct AssembleParams {
    const float* img0; const float* img1;   // [3][H][W] fp32
    const void* f0; const void* f1;         // [H][W][16] FT
    const float* timestep; float timestep_scalar;   // [H][W] or NULL -> scalar
    const float* flow;                      // [H][W][4] fp32 state, or NULL (first block: no warp, 39 channels)
    Tmp13 prev;                             // previous block's lastconv output (mask/feat source)
    Tmp13 fterm[2]; int nfterms;            // flow given as a SUM of up-sampled lastconv outputs (flow == NULL): coarse blocks
                                            // sample 1/16 or 1/4 of the pixels, so the full-resolution flow state is not
                                            // materialised for them (load_flow below)
    void* out; int out_cstride;             // NHWC: channels allocated per pixel
    int H, W, s, h, w;                      // full size, integer scale, h = H/s, w = W/s
};

// flow at (y, x): the materialised state, or s0 * up(tmp0)[0:4] (+ s1 * up(tmp1)[0:4]) in the order
// ifnet_flow_accum_kernel accumulates them (IFNet_HDv3.py:91-93, :157)
template <int TMP_LAYOUT>
__device__ __forceinline__ float4 load_flow(const AssembleParams& p, int y, int x)
{
    if (p.nfterms == 0) return reinterpret_cast<const float4*>(p.flow)[(size_t)y * p.W + x];
    float o[4];
    up_tmp<TMP_LAYOUT, 0, 4>(p.fterm[0], y, x, o);
    float fs = (float)p.fterm[0].s;
    float4 f = make_float4(o[0] * fs, o[1] * fs, o[2] * fs, o[3] * fs);
    if (p.nfterms > 1) {
        up_tmp<TMP_LAYOUT, 0, 4>(p.fterm[1], y, x, o);
        fs = (float)p.fterm[1].s;
        f.x = f.x + o[0] * fs; f.y = f.y + o[1] * fs; f.z = f.z + o[2] * fs; f.w = f.w + o[3] * fs;
    }
    return f;
}

// ifnet_tc.cu: the L1-friendly NHWC fp16 kernel
void launch_assemble_tc(const AssembleParams& p, cudaStream_t st);
void launch_flow_sum_tc(const Tmp13& t0, const Tmp13& t1, const Tmp13& t2, int nterms, float* flow, int H, int W, cudaStream_t st);

}  // namespace drba
