// Glue kernels of the GMFSS path (models/model_gmfss/GMFSS.py:58-190, MetricNet.py:45-65,
// gmflow/geometry.py:87-108) for sm_100a: everything between the conv programs that the reference
// spells as chains of small torch ops (cat / interpolate / grid_sample / norm / mul).
// The conv engine works on NHWC fp16; these kernels are the only places where the NCHW fp32 API tensors
// are packed into / unpacked from that layout, fused with the arithmetic that precedes the conv.
#include "common.cuh"

namespace drba {

constexpr int kGmThreads = 256;
constexpr int kMaxPlanes = 16;

struct PlaneList {
    const float* p[kMaxPlanes];
    float scale[kMaxPlanes];
};

// out[y][x][c] = prelu(scale[c] * plane_c[y][x]) for c < nplanes, 0 for the padding channels
__global__ void __launch_bounds__(kGmThreads)
pack_planes_kernel(const PlaneList pl, int nplanes, size_t HW, float slope, int use_prelu, __half* __restrict__ out, int cstride)
{
    const size_t i = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    if (i >= HW) return;
    float v[kMaxPlanes];
#pragma unroll
    for (int c = 0; c < kMaxPlanes; ++c) {
        float t = 0.0f;
        if (c < nplanes) {
            t = pl.p[c][i] * pl.scale[c];
            if (use_prelu) t = t > 0.0f ? t : slope * t;
        }
        v[c] = t;
    }
    uint4 o[2];
    __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    uint4* dst = reinterpret_cast<uint4*>(out + i * cstride);
    dst[0] = o[0]; dst[1] = o[1];
}

__global__ void __launch_bounds__(kGmThreads)
unpack_kernel(const __half* __restrict__ in, int cstride, float* __restrict__ out, int C, size_t HW, int do_clamp, float lo, float hi)
{
    const size_t i = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    if (i >= HW) return;
    for (int c = 0; c < C; ++c) {
        float v = __half2float(in[i * cstride + c]);
        if (do_clamp == 1) v = fminf(fmaxf(v, lo), hi);
        else if (do_clamp == 2) v = tanhf(v) * 10.0f;       // union MetricNet: Tanh() then * 10 (model_gmfss_union/MetricNet.py:41-42,63)
        out[(size_t)c * HW + i] = v;
    }
}

// zeros-padded bilinear sample at pixel coordinates (grid_sample align_corners=True, padding_mode='zeros')
__device__ __forceinline__ float sample_zeros(const float* __restrict__ src, float sx, float sy, int H, int W)
{
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float ax = sx - fx0, ay = sy - fy0;
    float acc = 0.0f;
    const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
    const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
    if (vx0 && vy0) acc += src[(size_t)y0 * W + x0] * ((1.0f - ax) * (1.0f - ay));
    if (vx1 && vy0) acc += src[(size_t)y0 * W + x0 + 1] * (ax * (1.0f - ay));
    if (vx0 && vy1) acc += src[(size_t)(y0 + 1) * W + x0] * ((1.0f - ax) * ay);
    if (vx1 && vy1) acc += src[(size_t)(y0 + 1) * W + x0 + 1] * (ax * ay);
    return acc;
}

// MetricNet.forward up to the first conv (MetricNet.py:45-60): photometric errors through zero-padded backward
// warps, forward-backward consistency masks (geometry.py:87-108, alpha 0.01, beta 0.5), normalised flows;
// written as the conv input [H][W][16] fp16:
//   img0 3 | img1 3 | -metric0 | -metric1 | flow01 / ((W-1)/2, (H-1)/2) | flow10 / (...) | fwd_occ | bwd_occ | 0 0
__global__ void __launch_bounds__(kGmThreads)
metric_prep_kernel(const float* __restrict__ img0, const float* __restrict__ img1,
                   const float* __restrict__ f01, const float* __restrict__ f10, __half* __restrict__ out, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const size_t i = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    if (i >= HW) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const float u01 = f01[i], v01 = f01[HW + i], u10 = f10[i], v10 = f10[HW + i];
    float v[16];
    float m0 = 0.0f, m1 = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = img0[(size_t)c * HW + i], b = img1[(size_t)c * HW + i];
        v[c] = a; v[3 + c] = b;
        m0 += fabsf(a - sample_zeros(img1 + (size_t)c * HW, (float)x + u01, (float)y + v01, H, W));
        m1 += fabsf(b - sample_zeros(img0 + (size_t)c * HW, (float)x + u10, (float)y + v10, H, W));
    }
    v[6] = -(m0 / 3.0f);
    v[7] = -(m1 / 3.0f);
    const float sx = ((float)W - 1.0f) / 2.0f, sy = ((float)H - 1.0f) / 2.0f;
    v[8] = u01 / sx; v[9] = v01 / sy; v[10] = u10 / sx; v[11] = v10 / sy;
    const float mag = sqrtf(u01 * u01 + v01 * v01) + sqrtf(u10 * u10 + v10 * v10);
    const float wbu = sample_zeros(f10, (float)x + u01, (float)y + v01, H, W);
    const float wbv = sample_zeros(f10 + HW, (float)x + u01, (float)y + v01, H, W);
    const float wfu = sample_zeros(f01, (float)x + u10, (float)y + v10, H, W);
    const float wfv = sample_zeros(f01 + HW, (float)x + u10, (float)y + v10, H, W);
    const float dfx = u01 + wbu, dfy = v01 + wbv, dbx = u10 + wfu, dby = v10 + wfv;
    const float thr = 0.01f * mag + 0.5f;
    v[12] = sqrtf(dfx * dfx + dfy * dfy) > thr ? 1.0f : 0.0f;
    v[13] = sqrtf(dbx * dbx + dby * dby) > thr ? 1.0f : 0.0f;
    v[14] = 0.0f; v[15] = 0.0f;
    uint4 o[2];
    __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
    dst[0] = o[0]; dst[1] = o[1];
}

// Model.inference's per-scale splat inputs (GMFSS.py:88-113): F = t * flow, Z = t * metric at the base
// resolution, then F.interpolate(., 1/s, bilinear, align_corners=False) (* 1/s for the flow).  For integer s this
// resampling is the mean of the 2x2 pixels at the centre of each s x s cell.
__global__ void __launch_bounds__(kGmThreads)
scale_flow_kernel(const float* __restrict__ flow, const float* __restrict__ metric, const float* __restrict__ tmap, float tscalar,
                  int H, int W, int s, float* __restrict__ oflow, float* __restrict__ ometric)
{
    const int h = H / s, w = W / s;
    const size_t hw = (size_t)h * w, HW = (size_t)H * W;
    const size_t i = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    if (i >= hw) return;
    const int Y = (int)(i / w), X = (int)(i - (size_t)Y * w);
    float fu, fv, z;
    if (s == 1) {
        const float t = tmap ? tmap[i] : tscalar;
        fu = t * flow[i]; fv = t * flow[HW + i]; z = t * metric[i];
    } else {
        const int off = s / 2 - 1;
        float au[2], av[2], az[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const size_t j0 = (size_t)(s * Y + off + r) * W + (s * X + off), j1 = j0 + 1;
            const float t0 = tmap ? tmap[j0] : tscalar, t1 = tmap ? tmap[j1] : tscalar;
            au[r] = 0.5f * (t0 * flow[j0]) + 0.5f * (t1 * flow[j1]);
            av[r] = 0.5f * (t0 * flow[HW + j0]) + 0.5f * (t1 * flow[HW + j1]);
            az[r] = 0.5f * (t0 * metric[j0]) + 0.5f * (t1 * metric[j1]);
        }
        const float inv = 1.0f / (float)s;
        fu = (0.5f * au[0] + 0.5f * au[1]) * inv;
        fv = (0.5f * av[0] + 0.5f * av[1]) * inv;
        z = 0.5f * az[0] + 0.5f * az[1];
    }
    oflow[i] = fu; oflow[hw + i] = fv; ometric[i] = z;
}


// ---- GMFSS_union: timestep alignment and swap masks (models/model_gmfss_union/GMFSS.py:118-152) ----------------
// t0w / t1w: the timestep maps forward-warped with their side's flow; g0 / g1: the warped ones-maps.  Holes of
// either side (g < 0.999) set both timesteps to 1 (:121-126).
__global__ void __launch_bounds__(kGmThreads)
union_fix_timesteps_kernel(float* __restrict__ t0w, float* __restrict__ t1w, const float* __restrict__ g0, const float* __restrict__ g1, size_t n)
{
    const size_t i = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    if (i >= n) return;
    if (g0[i] < 0.999f || g1[i] < 0.999f) { t0w[i] = 1.0f; t1w[i] = 1.0f; }
}

// mask0 = t0 / t1 > 25 takes side 2's value into side 1, mask1 = t1 / t0 > 25 the other way (:129-152); both read
// the values from before the swap.  NHWC concat buffer [hw][2C] = [side 1 | side 2].
__global__ void __launch_bounds__(kGmThreads)
union_swap_nhwc_kernel(__half* __restrict__ x, int C, const float* __restrict__ t0, const float* __restrict__ t1, size_t hw)
{
    const size_t i = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    const int G = C / 8;
    if (i >= hw * (size_t)G) return;
    const size_t p = i / G;
    const int g = (int)(i - p * G);
    const bool m0 = t0[p] / t1[p] > 25.0f, m1 = t1[p] / t0[p] > 25.0f;
    if (!m0 && !m1) return;
    uint4* a = reinterpret_cast<uint4*>(x + p * 2 * C + g * 8);
    uint4* b = reinterpret_cast<uint4*>(x + p * 2 * C + C + g * 8);
    const uint4 va = *a, vb = *b;
    if (m0) *a = vb;
    if (m1) *b = va;
}

__global__ void __launch_bounds__(kGmThreads)
union_swap_nchw_kernel(float* __restrict__ a, float* __restrict__ b, int C, const float* __restrict__ t0, const float* __restrict__ t1, size_t hw)
{
    const size_t p = (size_t)blockIdx.x * kGmThreads + threadIdx.x;
    if (p >= hw) return;
    const bool m0 = t0[p] / t1[p] > 25.0f, m1 = t1[p] / t0[p] > 25.0f;
    if (!m0 && !m1) return;
    for (int c = 0; c < C; ++c) {
        const float va = a[(size_t)c * hw + p], vb = b[(size_t)c * hw + p];
        if (m0) a[(size_t)c * hw + p] = vb;
        if (m1) b[(size_t)c * hw + p] = va;
    }
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_pack_planes_nhwc_f16(const float* const* planes, const float* scales, int nplanes, int H, int W,
                              int use_prelu, float slope, void* out, int out_cstride, void* stream)
{
    if (!planes || nplanes < 1 || nplanes > kMaxPlanes || H <= 0 || W <= 0 || !out) return DRBA_E_ARG;
    if (out_cstride < 16 || out_cstride % 8 != 0) return DRBA_E_ARG;
    if (!aligned16(out)) return DRBA_E_ALIGN;
    PlaneList pl;
    for (int c = 0; c < kMaxPlanes; ++c) {
        pl.p[c] = c < nplanes ? planes[c] : nullptr;
        pl.scale[c] = (c < nplanes && scales) ? scales[c] : 1.0f;
        if (c < nplanes && !planes[c]) return DRBA_E_ARG;
    }
    const size_t HW = (size_t)H * W;
    pack_planes_kernel<<<cdiv(HW, kGmThreads), kGmThreads, 0, as_stream(stream)>>>(pl, nplanes, HW, slope, use_prelu, (__half*)out, out_cstride);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_unpack_nhwc_f16(const void* in, int in_cstride, float* out, int C, int H, int W, int do_clamp, float lo, float hi, void* stream)
{
    if (!in || !out || C < 1 || C > in_cstride || H <= 0 || W <= 0) return DRBA_E_ARG;
    const size_t HW = (size_t)H * W;
    unpack_kernel<<<cdiv(HW, kGmThreads), kGmThreads, 0, as_stream(stream)>>>((const __half*)in, in_cstride, out, C, HW, do_clamp, lo, hi);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmfss_metric_prep(const float* img0, const float* img1, const float* flow01, const float* flow10,
                           void* out_nhwc16, int H, int W, void* stream)
{
    if (!img0 || !img1 || !flow01 || !flow10 || !out_nhwc16 || H <= 1 || W <= 1) return DRBA_E_ARG;
    if (!aligned16(out_nhwc16)) return DRBA_E_ALIGN;
    metric_prep_kernel<<<cdiv((size_t)H * W, kGmThreads), kGmThreads, 0, as_stream(stream)>>>(img0, img1, flow01, flow10, (__half*)out_nhwc16, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmfss_scale_flow(const float* flow, const float* metric, const float* tmap, float tscalar,
                          int H, int W, int s, float* out_flow, float* out_metric, void* stream)
{
    if (!flow || !metric || !out_flow || !out_metric || H <= 0 || W <= 0) return DRBA_E_ARG;
    if (s != 1 && s != 2 && s != 4) return DRBA_E_ARG;
    if (H % s != 0 || W % s != 0) return DRBA_E_ARG;
    scale_flow_kernel<<<cdiv((size_t)(H / s) * (W / s), kGmThreads), kGmThreads, 0, as_stream(stream)>>>(
        flow, metric, tmap, tscalar, H, W, s, out_flow, out_metric);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmfss_union_fix_timesteps(float* t0w, float* t1w, const float* g0, const float* g1, size_t n, void* stream)
{
    if (!t0w || !t1w || !g0 || !g1) return DRBA_E_ARG;
    if (n == 0) return DRBA_OK;
    union_fix_timesteps_kernel<<<cdiv(n, kGmThreads), kGmThreads, 0, as_stream(stream)>>>(t0w, t1w, g0, g1, n);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmfss_union_swap_nhwc_f16(void* x, int C, const float* t0, const float* t1, int h, int w, void* stream)
{
    if (!x || !t0 || !t1 || C <= 0 || C % 8 != 0 || h <= 0 || w <= 0) return DRBA_E_ARG;
    if (!aligned16(x)) return DRBA_E_ALIGN;
    const size_t hw = (size_t)h * w;
    union_swap_nhwc_kernel<<<cdiv(hw * (C / 8), kGmThreads), kGmThreads, 0, as_stream(stream)>>>((__half*)x, C, t0, t1, hw);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmfss_union_swap_nchw_f32(float* a, float* b, int C, const float* t0, const float* t1, int h, int w, void* stream)
{
    if (!a || !b || !t0 || !t1 || C <= 0 || h <= 0 || w <= 0) return DRBA_E_ARG;
    const size_t hw = (size_t)h * w;
    union_swap_nchw_kernel<<<cdiv(hw, kGmThreads), kGmThreads, 0, as_stream(stream)>>>(a, b, C, t0, t1, hw);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
