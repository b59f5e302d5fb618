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
//  ifnet_upsample : lastconv output (13 ch at 1/s) -> bilinear x s -> flow(+=) * s, mask, feat
//                   (IFNet_HDv3.py:91-96, :156-158) into the full-resolution state tensor
//                   [H][W][16] fp32 = {flow 4, mask 1, feat 8, pad 3}.
//  ifnet_blend    : final warps + sigmoid blend (IFNet_HDv3.py:162-167).
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

struct AssembleParams {
    const float* img0; const float* img1;   // [3][H][W] fp32
    const void* f0; const void* f1;         // [H][W][16] FT
    const float* timestep; float timestep_scalar;   // [H][W] or NULL -> scalar
    const float* state;                     // [H][W][16] fp32 or NULL (first block: no warp, 39 channels)
    void* out; int out_cstride;             // NHWC: channels allocated per pixel (>= 52/39, rest zero-filled)
    int H, W, s, h, w;                      // full size, integer scale, h = H/s, w = W/s
};

template <bool NHWC_HALF>
__device__ __forceinline__ void store_channels(const AssembleParams& p, int Y, int X, int c0, const float* v, int n)
{
    if (NHWC_HALF) {
        __half* o = reinterpret_cast<__half*>(p.out) + ((size_t)Y * p.w + X) * p.out_cstride + c0;
        for (int c = 0; c < n; ++c) o[c] = __float2half_rn(v[c]);
    } else {
        float* o = reinterpret_cast<float*>(p.out) + (size_t)c0 * p.h * p.w + (size_t)Y * p.w + X;
        for (int c = 0; c < n; ++c) o[(size_t)c * p.h * p.w] = v[c];
    }
}

// mean of the (up to) 4 sampled positions in the order bilinear resampling adds them:
// 0.5*(0.5*a + 0.5*b) + 0.5*(0.5*c + 0.5*d) == ((a + b) + (c + d)) * 0.25 exactly in fp32
__device__ __forceinline__ float mean4(const float* q, int np)
{
    return np == 1 ? q[0] : ((q[0] + q[1]) + (q[2] + q[3])) * 0.25f;
}

template <typename FT, bool NHWC_HALF>
__global__ void __launch_bounds__(kIfThreads)
ifnet_assemble_kernel(const AssembleParams p)
{
    const int idx = blockIdx.x * kIfThreads + threadIdx.x;
    if (idx >= p.h * p.w) return;
    const int Y = idx / p.w, X = idx - Y * p.w;
    const int np = p.s == 1 ? 1 : 4;
    const int off = p.s == 1 ? 0 : p.s / 2 - 1;
    const size_t HW = (size_t)p.H * p.W;
    const FT* f0 = reinterpret_cast<const FT*>(p.f0);
    const FT* f1 = reinterpret_cast<const FT*>(p.f1);

    int px[4], py[4];
    WarpTap t0[4], t1[4];
    float st[4][16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= np) break;
        py[k] = p.s * Y + off + (k >> 1);
        px[k] = p.s * X + off + (k & 1);
        if (p.state) {
            load16(p.state + ((size_t)py[k] * p.W + px[k]) * 16, st[k]);
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) st[k][c] = 0.0f;
        }
        t0[k] = warp_tap(px[k], py[k], st[k][0], st[k][1], p.H, p.W);
        t1[k] = warp_tap(px[k], py[k], st[k][2], st[k][3], p.H, p.W);
    }
    float q[4], v[16];
    // warped images (first block: the images themselves -- the identity warp is exact)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int k = 0; k < np; ++k) q[k] = sample_plane(p.img0 + (size_t)c * HW, t0[k]);
        v[c] = mean4(q, np);
        for (int k = 0; k < np; ++k) q[k] = sample_plane(p.img1 + (size_t)c * HW, t1[k]);
        v[3 + c] = mean4(q, np);
    }
    store_channels<NHWC_HALF>(p, Y, X, 0, v, 6);
    // warped features
    for (int side = 0; side < 2; ++side) {
        float fq[4][16];
        for (int k = 0; k < np; ++k) sample_feat16<FT>(side == 0 ? f0 : f1, side == 0 ? t0[k] : t1[k], fq[k]);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            for (int k = 0; k < np; ++k) q[k] = fq[k][c];
            v[c] = mean4(q, np);
        }
        store_channels<NHWC_HALF>(p, Y, X, 6 + 16 * side, v, 16);
    }
    // timestep | mask | feat | flow / s
    for (int k = 0; k < np; ++k) q[k] = p.timestep ? p.timestep[(size_t)py[k] * p.W + px[k]] : p.timestep_scalar;
    v[0] = mean4(q, np);
    int nch = 1;
    if (p.state) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {   // mask (state ch 4), feat (5..12)
            for (int k = 0; k < np; ++k) q[k] = st[k][4 + c];
            v[1 + c] = mean4(q, np);
        }
        const float inv = 1.0f / (float)p.s;   // IFNet_HDv3.py:87: interpolate(flow) * 1. / scale
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            for (int k = 0; k < np; ++k) q[k] = st[k][c];
            v[10 + c] = mean4(q, np) * 1.0f * inv;
        }
        nch = 14;
    }
    store_channels<NHWC_HALF>(p, Y, X, 38, v, nch);
    if (NHWC_HALF) {
        const int used = 38 + nch;
        __half* o = reinterpret_cast<__half*>(p.out) + ((size_t)Y * p.w + X) * p.out_cstride;
        for (int c = used; c < p.out_cstride; ++c) o[c] = __float2half_rn(0.0f);
    }
}

// ---- lastconv output -> full-resolution state ----------------------------------------
// TMP_LAYOUT 0: ConvTranspose output, NCHW fp32 [52][H/(2s)][W/(2s)] (PixelShuffle(2) is
//               index arithmetic here: IFNet_HDv3.py:81)
// TMP_LAYOUT 1: pixel-shuffled NHWC fp32 [H/s][W/s][16] (13 used), written by the tensor-core
//               engine's lastconv epilogue
template <int TMP_LAYOUT>
__device__ __forceinline__ void load_tmp13(const float* __restrict__ tmp, int yy, int xx, int h13, int w13, float* v)
{
    if (TMP_LAYOUT == 0) {
        const int h2 = h13 >> 1, w2 = w13 >> 1;
        const size_t base = (size_t)(yy >> 1) * w2 + (xx >> 1);
        const int sub = (yy & 1) * 2 + (xx & 1);
#pragma unroll
        for (int c = 0; c < 13; ++c) v[c] = tmp[(size_t)(c * 4 + sub) * h2 * w2 + base];
    } else {
        float t[16];
        load16(tmp + ((size_t)yy * w13 + xx) * 16, t);
#pragma unroll
        for (int c = 0; c < 13; ++c) v[c] = t[c];
    }
}

template <int TMP_LAYOUT>
__global__ void __launch_bounds__(kIfThreads)
ifnet_upsample_kernel(const float* __restrict__ tmp, float* __restrict__ state, int accumulate, int H, int W, int s)
{
    const int idx = blockIdx.x * kIfThreads + threadIdx.x;
    if (idx >= H * W) return;
    const int y = idx / W, x = idx - y * W;
    const int h13 = H / s, w13 = W / s;
    const float r = 1.0f / (float)s;                   // ATen: ratio = 1 / scale_factor
    float sy = r * ((float)y + 0.5f) - 0.5f, sx = r * ((float)x + 0.5f) - 0.5f;
    if (sy < 0.0f) sy = 0.0f;
    if (sx < 0.0f) sx = 0.0f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h13 - 1 ? 1 : 0), x1 = x0 + (x0 < w13 - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.0f - ly, lx = sx - (float)x0, hx = 1.0f - lx;
    float a[13], b[13], c[13], d[13];
    load_tmp13<TMP_LAYOUT>(tmp, y0, x0, h13, w13, a);
    load_tmp13<TMP_LAYOUT>(tmp, y0, x1, h13, w13, b);
    load_tmp13<TMP_LAYOUT>(tmp, y1, x0, h13, w13, c);
    load_tmp13<TMP_LAYOUT>(tmp, y1, x1, h13, w13, d);
    float o[16];
#pragma unroll
    for (int k = 0; k < 13; ++k) o[k] = hy * (hx * a[k] + lx * b[k]) + ly * (hx * c[k] + lx * d[k]);
    o[13] = o[14] = o[15] = 0.0f;
    float4* dst = reinterpret_cast<float4*>(state + (size_t)idx * 16);
    const float fs = (float)s;
    float4 f = make_float4(o[0] * fs, o[1] * fs, o[2] * fs, o[3] * fs);   // IFNet_HDv3.py:93
    if (accumulate) {                                                     // IFNet_HDv3.py:157
        const float4 old = dst[0];
        f.x = old.x + f.x; f.y = old.y + f.y; f.z = old.z + f.z; f.w = old.w + f.w;
    }
    dst[0] = f;
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    dst[2] = make_float4(o[8], o[9], o[10], o[11]);
    dst[3] = make_float4(o[12], 0.0f, 0.0f, 0.0f);
}

// IFNet_HDv3.py:160-167: warp both images with the final flow, blend with sigmoid(mask)
__global__ void __launch_bounds__(kIfThreads)
ifnet_blend_kernel(const float* __restrict__ img0, const float* __restrict__ img1,
                   const float* __restrict__ state, float* __restrict__ out, int H, int W)
{
    const int idx = blockIdx.x * kIfThreads + threadIdx.x;
    if (idx >= H * W) return;
    const int y = idx / W, x = idx - y * W;
    const float4* st = reinterpret_cast<const float4*>(state + (size_t)idx * 16);
    const float4 fl = st[0];
    const float mk = st[1].x;
    const WarpTap t0 = warp_tap(x, y, fl.x, fl.y, H, W);
    const WarpTap t1 = warp_tap(x, y, fl.z, fl.w, H, W);
    const float m = 1.0f / (1.0f + expf(-mk));
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float w0 = sample_plane(img0 + (size_t)c * HW, t0);
        const float w1 = sample_plane(img1 + (size_t)c * HW, t1);
        out[(size_t)c * HW + idx] = w0 * m + w1 * (1.0f - m);
    }
}

// state[..., 0:4] -> NCHW flow [4][H][W] (calc_flow needs block0's flow as planar tensors)
__global__ void __launch_bounds__(kIfThreads)
ifnet_state_flow_kernel(const float* __restrict__ state, float* __restrict__ flow, int HW)
{
    const int idx = blockIdx.x * kIfThreads + threadIdx.x;
    if (idx >= HW) return;
    const float4 f = reinterpret_cast<const float4*>(state + (size_t)idx * 16)[0];
    flow[idx] = f.x; flow[(size_t)HW + idx] = f.y; flow[(size_t)2 * HW + idx] = f.z; flow[(size_t)3 * HW + idx] = f.w;
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_ifnet_assemble(const float* img0, const float* img1, const void* f0, const void* f1, int feat_dtype,
                        const float* timestep, float timestep_scalar, const float* state,
                        void* out, int out_dtype, int out_cstride, int H, int W, int s, void* stream)
{
    if (H <= 0 || W <= 0 || s <= 0 || (s & (s - 1)) != 0 || H % s != 0 || W % s != 0) return DRBA_E_ARG;
    if (!img0 || !img1 || !f0 || !f1 || !out) return DRBA_E_ARG;
    if (feat_dtype != DRBA_F32 && feat_dtype != DRBA_F16) return DRBA_E_ARG;
    if (out_dtype != DRBA_F32 && out_dtype != DRBA_F16) return DRBA_E_ARG;
    const int nch = state ? 52 : 39;
    if (out_dtype == DRBA_F16 && out_cstride < nch) return DRBA_E_ARG;
    if (!aligned16(f0) || !aligned16(f1) || (state && !aligned16(state))) return DRBA_E_ALIGN;
    AssembleParams p;
    p.img0 = img0; p.img1 = img1; p.f0 = f0; p.f1 = f1;
    p.timestep = timestep; p.timestep_scalar = timestep_scalar; p.state = state;
    p.out = out; p.out_cstride = out_cstride; p.H = H; p.W = W; p.s = s; p.h = H / s; p.w = W / s;
    const unsigned grid = cdiv((size_t)p.h * p.w, kIfThreads);
    cudaStream_t st = as_stream(stream);
    if (feat_dtype == DRBA_F32 && out_dtype == DRBA_F32) ifnet_assemble_kernel<float, false><<<grid, kIfThreads, 0, st>>>(p);
    else if (feat_dtype == DRBA_F16 && out_dtype == DRBA_F16) ifnet_assemble_kernel<__half, true><<<grid, kIfThreads, 0, st>>>(p);
    else if (feat_dtype == DRBA_F32 && out_dtype == DRBA_F16) ifnet_assemble_kernel<float, true><<<grid, kIfThreads, 0, st>>>(p);
    else ifnet_assemble_kernel<__half, false><<<grid, kIfThreads, 0, st>>>(p);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_ifnet_upsample(const float* tmp, int tmp_layout, float* state, int accumulate, int H, int W, int s, void* stream)
{
    if (H <= 0 || W <= 0 || s <= 0 || H % s != 0 || W % s != 0 || !tmp || !state) return DRBA_E_ARG;
    if (tmp_layout != 0 && tmp_layout != 1) return DRBA_E_ARG;
    if (tmp_layout == 0 && ((H / s) % 2 != 0 || (W / s) % 2 != 0)) return DRBA_E_ARG;
    if (!aligned16(state) || (tmp_layout == 1 && !aligned16(tmp))) return DRBA_E_ALIGN;
    const unsigned grid = cdiv((size_t)H * W, kIfThreads);
    if (tmp_layout == 0) ifnet_upsample_kernel<0><<<grid, kIfThreads, 0, as_stream(stream)>>>(tmp, state, accumulate, H, W, s);
    else ifnet_upsample_kernel<1><<<grid, kIfThreads, 0, as_stream(stream)>>>(tmp, state, accumulate, H, W, s);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_ifnet_blend(const float* img0, const float* img1, const float* state, float* out, int H, int W, void* stream)
{
    if (H <= 0 || W <= 0 || !img0 || !img1 || !state || !out) return DRBA_E_ARG;
    if (!aligned16(state)) return DRBA_E_ALIGN;
    ifnet_blend_kernel<<<cdiv((size_t)H * W, kIfThreads), kIfThreads, 0, as_stream(stream)>>>(img0, img1, state, out, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_ifnet_state_flow(const float* state, float* flow, int H, int W, void* stream)
{
    if (H <= 0 || W <= 0 || !state || !flow) return DRBA_E_ARG;
    if (!aligned16(state)) return DRBA_E_ALIGN;
    ifnet_state_flow_kernel<<<cdiv((size_t)H * W, kIfThreads), kIfThreads, 0, as_stream(stream)>>>(state, flow, H * W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
