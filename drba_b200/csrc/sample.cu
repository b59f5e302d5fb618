// Backward warp and bilinear resize (NCHW fp32) for sm_100a.
//
// Stand-alone operator forms of what the fused IFNet kernels (ifnet.cu) do inline;
// they exist for the operator seam (MetricNet.backwarp, warplayer.warp, F.interpolate
// call sites) and as building blocks for parity tests.  HBM-bound gathers: one thread
// per output pixel, coordinates computed once and reused across channels, channel
// planes read/written with unit stride across the warp.
#include "common.cuh"

namespace drba {

constexpr int kSampleThreads = 256;

// models/rife_426_heavy/warplayer.py:8-22 in pixel coordinates
__global__ void __launch_bounds__(kSampleThreads)
backwarp_kernel(const float* __restrict__ in, const float* __restrict__ flow, float* __restrict__ out,
                int N, int C, int H, int W, int pad_mode)
{
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * kSampleThreads + threadIdx.x;
    if (p >= (size_t)N * HW) return;
    const int n = (int)(p / HW);
    const size_t r = p - (size_t)n * HW;
    const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
    float sx = (float)x + flow[((size_t)n * 2) * HW + r];
    float sy = (float)y + flow[((size_t)n * 2 + 1) * HW + r];
    if (pad_mode == DRBA_PAD_BORDER) {
        sx = fminf(fmaxf(sx, 0.0f), (float)(W - 1));
        sy = fminf(fmaxf(sy, 0.0f), (float)(H - 1));
    }
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    // non-finite or huge coordinates sample nothing (zeros); border mode is already clamped
    const bool fin = fabsf(fx0) < 1.0e9f && fabsf(fy0) < 1.0e9f;
    const int x0 = fin ? (int)fx0 : -4, y0 = fin ? (int)fy0 : -4;
    const float ax = sx - fx0, ay = sy - fy0;
    const float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay);
    const float w10 = (1.0f - ax) * ay, w11 = ax * ay;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
    const long long q = (long long)y0 * W + x0;
    const float* src = in + (size_t)n * C * HW;
    float* dst = out + (size_t)n * C * HW + r;
    for (int c = 0; c < C; ++c) {
        const float* s = src + (size_t)c * HW;
        float acc = 0.0f;
        if (vx0 && vy0) acc += s[q] * w00;
        if (vx1 && vy0) acc += s[q + 1] * w01;
        if (vx0 && vy1) acc += s[q + W] * w10;
        if (vx1 && vy1) acc += s[q + W + 1] * w11;
        dst[(size_t)c * HW] = acc;
    }
}

// F.interpolate(mode='bilinear'); ATen's upsample_bilinear2d index math
__global__ void __launch_bounds__(kSampleThreads)
resize_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int H, int W,
                       int OH, int OW, int align_corners, float rh, float rw)
{
    const size_t OHW = (size_t)OH * OW;
    const size_t p = (size_t)blockIdx.x * kSampleThreads + threadIdx.x;
    if (p >= OHW) return;
    const int oy = (int)(p / OW), ox = (int)(p - (size_t)oy * OW);
    float sy = align_corners ? rh * (float)oy : rh * ((float)oy + 0.5f) - 0.5f;
    float sx = align_corners ? rw * (float)ox : rw * ((float)ox + 0.5f) - 0.5f;
    if (!align_corners && sy < 0.0f) sy = 0.0f;
    if (!align_corners && sx < 0.0f) sx = 0.0f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.0f - ly;
    const float lx = sx - (float)x0, hx = 1.0f - lx;
    const size_t i00 = (size_t)y0 * W + x0, i01 = (size_t)y0 * W + x1;
    const size_t i10 = (size_t)y1 * W + x0, i11 = (size_t)y1 * W + x1;
    const size_t HW = (size_t)H * W;
    for (int c = blockIdx.y; c < NC; c += gridDim.y) {
        const float* s = in + (size_t)c * HW;
        out[(size_t)c * OHW + p] = hy * (hx * s[i00] + lx * s[i01]) + ly * (hx * s[i10] + lx * s[i11]);
    }
}


// ---- frame ingest / egress (models/utils/tools.py:33-38, :59-72) -----------------------------------
// to_inp:  uint8 HWC (BGR as decoded) -> float NCHW / 255 -> F.interpolate(size, bilinear, align_corners=False)
// to_out:  F.interpolate(src_size) -> * 255. -> astype(uint8) (C cast: truncation, wraps outside [0, 256))
// One pass each: the fp32 full-size intermediate of the reference never exists, and the host<->device
// copies carry 1 byte per sample instead of 4.
__global__ void __launch_bounds__(kSampleThreads)
frame_ingest_u8_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int H, int W, int OH, int OW,
                       float rh, float rw)
{
    const size_t OHW = (size_t)OH * OW;
    const size_t p = (size_t)blockIdx.x * kSampleThreads + threadIdx.x;
    if (p >= OHW) return;
    const int oy = (int)(p / OW), ox = (int)(p - (size_t)oy * OW);
    float sy = rh * ((float)oy + 0.5f) - 0.5f, sx = rw * ((float)ox + 0.5f) - 0.5f;
    if (sy < 0.0f) sy = 0.0f;
    if (sx < 0.0f) sx = 0.0f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.0f - ly;
    const float lx = sx - (float)x0, hx = 1.0f - lx;
    const unsigned char* r0 = in + ((size_t)y0 * W) * 3;
    const unsigned char* r1 = in + ((size_t)y1 * W) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = (float)r0[x0 * 3 + c] / 255.0f, b = (float)r0[x1 * 3 + c] / 255.0f;
        const float cc = (float)r1[x0 * 3 + c] / 255.0f, d = (float)r1[x1 * 3 + c] / 255.0f;
        out[(size_t)c * OHW + p] = hy * (hx * a + lx * b) + ly * (hx * cc + lx * d);
    }
}

__global__ void __launch_bounds__(kSampleThreads)
frame_egress_u8_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, int H, int W, int OH, int OW,
                       float rh, float rw)
{
    const size_t OHW = (size_t)OH * OW;
    const size_t p = (size_t)blockIdx.x * kSampleThreads + threadIdx.x;
    if (p >= OHW) return;
    const int oy = (int)(p / OW), ox = (int)(p - (size_t)oy * OW);
    float sy = rh * ((float)oy + 0.5f) - 0.5f, sx = rw * ((float)ox + 0.5f) - 0.5f;
    if (sy < 0.0f) sy = 0.0f;
    if (sx < 0.0f) sx = 0.0f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.0f - ly;
    const float lx = sx - (float)x0, hx = 1.0f - lx;
    const size_t i00 = (size_t)y0 * W + x0, i01 = (size_t)y0 * W + x1;
    const size_t i10 = (size_t)y1 * W + x0, i11 = (size_t)y1 * W + x1;
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* s = in + (size_t)c * HW;
        const float v = (hy * (hx * s[i00] + lx * s[i01]) + ly * (hx * s[i10] + lx * s[i11])) * 255.0f;
        // numpy float32 -> uint8: truncate toward zero, keep the low 8 bits
        const float t = truncf(v);
        const int iv = (t >= -2147483648.0f && t < 2147483648.0f) ? (int)t : 0;
        out[p * 3 + c] = (unsigned char)(iv & 0xff);
    }
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_backwarp_f32(const float* in, const float* flow, float* out,
                      int N, int C, int H, int W, int pad_mode, void* stream)
{
    if (N < 0 || C < 0 || H < 0 || W < 0) return DRBA_E_ARG;
    if (pad_mode != DRBA_PAD_BORDER && pad_mode != DRBA_PAD_ZEROS) return DRBA_E_ARG;
    if ((size_t)N * C * H * W == 0) return DRBA_OK;
    if (!in || !flow || !out) return DRBA_E_ARG;
    backwarp_kernel<<<cdiv((size_t)N * H * W, kSampleThreads), kSampleThreads, 0, as_stream(stream)>>>(
        in, flow, out, N, C, H, W, pad_mode);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_resize_bilinear_f32(const float* in, float* out, int N, int C, int H, int W,
                             int OH, int OW, int align_corners, float rh, float rw, void* stream)
{
    if (N < 0 || C < 0 || H <= 0 || W <= 0 || OH < 0 || OW < 0) return DRBA_E_ARG;
    if ((size_t)N * C * OH * OW == 0) return DRBA_OK;
    if (!in || !out) return DRBA_E_ARG;
    if (align_corners) {
        rh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.0f;
        rw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.0f;
    }
    const int NC = N * C;
    dim3 grid(cdiv((size_t)OH * OW, kSampleThreads), NC < 64 ? NC : 64);
    resize_bilinear_kernel<<<grid, kSampleThreads, 0, as_stream(stream)>>>(in, out, NC, H, W, OH, OW,
                                                                           align_corners ? 1 : 0, rh, rw);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_frame_ingest_u8(const unsigned char* in_hwc, float* out_chw, int H, int W, int OH, int OW, void* stream)
{
    if (H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || !in_hwc || !out_chw) return DRBA_E_ARG;
    frame_ingest_u8_kernel<<<cdiv((size_t)OH * OW, kSampleThreads), kSampleThreads, 0, as_stream(stream)>>>(
        in_hwc, out_chw, H, W, OH, OW, (float)H / (float)OH, (float)W / (float)OW);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_frame_egress_u8(const float* in_chw, unsigned char* out_hwc, int H, int W, int OH, int OW, void* stream)
{
    if (H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || !in_chw || !out_hwc) return DRBA_E_ARG;
    frame_egress_u8_kernel<<<cdiv((size_t)OH * OW, kSampleThreads), kSampleThreads, 0, as_stream(stream)>>>(
        in_chw, out_hwc, H, W, OH, OW, (float)H / (float)OH, (float)W / (float)OW);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
