// Fused DistanceRatioMap kernels for sm_100a.
//
// Replaces models/drm.py (get_drm_t :10-62, calc_drm_rife :65-107, calc_drm_gmfss :110-155,
// calc_drm_rife_auxiliary :158-195) and distance_calculator (models/utils/tools.py:77-80):
// ~35 elementwise launches + 4 softsplats + 2 boolean-mask host syncs per call in the
// reference become one scatter kernel + one resolve kernel.
//
//  scatter: read both flows (16 B/px) [+ metrics], distance -> ratio -> time scaling,
//           splat (value*w, w) as ONE red.global.add.v2.f32 per corner into an L2-resident
//           accumulator [map][N*H*W][2];
//  resolve: re-derive the unaligned value from the flows (L2 hits), normalise, apply the
//           `mask < 0.999` hole fill (the ones-mask splat of the reference equals
//           den / (den + 1e-7) of the value splat), write the map, re-zero
//           the accumulator.
// Arithmetic follows SURVEY.md appendix A.2-A.4 operation by operation (this file is
// compiled with -fmad=false so the rounding sequence is the reference's).
#include "common.cuh"

namespace drba {

constexpr int kDrmThreads = 256;

// get_drm_t's branch sequence depends on the scalar t only (drm.py:40-60): computed on the
// host in double like the Python loop, replayed per pixel on the device.
struct DrmSeq {
    unsigned long long bits;  // bit k = 1 -> "x < t" branch at step k
    int len;                  // -1: linear scaling (drm * t * 2)
    float t;
};

static int build_seq(double t, double precision, DrmSeq* s)
{
    double x = 0.5, b = 0.5, l = 0.0, r = 1.0;
    unsigned long long bits = 0;
    int len = 0;
    while (fabs(x - t) > precision) {
        if (x > t) { r = x; x = x - (x - l) * b; if (len >= 64) return DRBA_E_UNSUPPORTED; ++len; }
        if (x < t) { l = x; x = x + (r - x) * b; if (len >= 64) return DRBA_E_UNSUPPORTED; bits |= 1ull << len; ++len; }
        if (!(x > t) && !(x < t) && fabs(x - t) > precision) return DRBA_E_ARG;  // NaN t
    }
    s->bits = bits;
    s->len = len;
    s->t = (float)t;
    return DRBA_OK;
}

__device__ __forceinline__ float drm_time_scale(float drm, const DrmSeq& s)
{
    if (s.len < 0) return drm * s.t * 2.0f;                      // drm.py:74-76
    float xd = drm, ld = drm * 0.0f, rd = drm * 0.0f + 1.0f;     // drm.py:37-38
    for (int k = 0; k < s.len; ++k) {
        if ((s.bits >> k) & 1ull) { ld = xd; xd = xd + (rd - xd) * drm; }   // drm.py:53-60
        else                      { rd = xd; xd = xd - (xd - ld) * drm; }   // drm.py:44-51
    }
    return xd;
}

__global__ void __launch_bounds__(kDrmThreads)
get_drm_t_kernel(const float* __restrict__ drm, float* __restrict__ out, size_t n, DrmSeq s)
{
    const size_t i = (size_t)blockIdx.x * kDrmThreads + threadIdx.x;
    if (i < n) out[i] = drm_time_scale(drm[i], s);
}

struct DrmPix {
    float f10x, f10y, f12x, f12y;
    float u0, u1;  // time-scaled drm10 / drm12 ("drm_t0_unaligned" / "drm_t1_unaligned")
};

template <bool EPS>
__device__ __forceinline__ DrmPix drm_pixel_vals(float f10x, float f10y, float f12x, float f12y, const DrmSeq& s)
{
    DrmPix px;
    px.f10x = f10x; px.f10y = f10y; px.f12x = f12x; px.f12y = f12y;
    float d10 = sqrtf(px.f10x * px.f10x + px.f10y * px.f10y);   // tools.py:77-80
    float d12 = sqrtf(px.f12x * px.f12x + px.f12y * px.f12y);
    if (EPS) { d10 = d10 + 1e-4f; d12 = d12 + 1e-4f; }          // drm.py:67-68 (not in :112-113)
    const float sum = d10 + d12;
    px.u0 = drm_time_scale(d10 / sum, s);
    px.u1 = drm_time_scale(d12 / sum, s);
    return px;
}

template <bool EPS>
__device__ __forceinline__ DrmPix drm_pixel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                                            int n, size_t r, size_t HW, const DrmSeq& s)
{
    return drm_pixel_vals<EPS>(flow10[((size_t)n * 2) * HW + r], flow10[((size_t)n * 2 + 1) * HW + r],
                               flow12[((size_t)n * 2) * HW + r], flow12[((size_t)n * 2 + 1) * HW + r], s);
}

// one (value*w, w) splat with x-neighbour aggregation; every lane of the warp must call it
__device__ __forceinline__ void splat2(float* __restrict__ acc2, int x, int y, int H, int W, float fx, float fy,
                                       float val, float wgt, bool valid, int n, int lane)
{
    Footprint f = footprint(x, y, fx, fy);
    f.ok = f.ok && valid;
    const bool inx0 = f.x0 >= 0 && f.x0 < W, inx1 = f.x0 + 1 >= 0 && f.x0 + 1 < W;
    const bool iny0 = f.y0 >= 0 && f.y0 < H, iny1 = f.y0 + 1 >= 0 && f.y0 + 1 < H;
    const long long q = (long long)f.y0 * W + f.x0;
    const unsigned full = 0xffffffffu;
    const int nx0 = __shfl_down_sync(full, f.x0, 1);
    const int ny0 = __shfl_down_sync(full, f.y0, 1);
    const int nok = __shfl_down_sync(full, (int)f.ok, 1);
    const int nn = __shfl_down_sync(full, n, 1);
    const bool chain_next = f.ok && lane < 31 && nok && nn == n && nx0 == f.x0 + 1 && ny0 == f.y0;
    const bool chain_prev = __shfl_up_sync(full, (int)chain_next, 1) && lane > 0;
    const float v = val * wgt;
    float a0 = v * f.nw, a1 = wgt * f.nw, b0 = v * f.ne, b1 = wgt * f.ne;
    float c0 = v * f.sw, c1 = wgt * f.sw, d0 = v * f.se, d1 = wgt * f.se;
    const float bp0 = __shfl_up_sync(full, b0, 1), bp1 = __shfl_up_sync(full, b1, 1);
    const float dp0 = __shfl_up_sync(full, d0, 1), dp1 = __shfl_up_sync(full, d1, 1);
    if (chain_prev) { a0 += bp0; a1 += bp1; c0 += dp0; c1 += dp1; }
    if (f.ok && inx0 && iny0) red_add_v2(acc2 + q * 2, a0, a1);
    if (f.ok && inx0 && iny1) red_add_v2(acc2 + (q + W) * 2, c0, c1);
    if (!chain_next) {
        if (f.ok && inx1 && iny0) red_add_v2(acc2 + (q + 1) * 2, b0, b1);
        if (f.ok && inx1 && iny1) red_add_v2(acc2 + (q + W + 1) * 2, d0, d1);
    }
}

struct PixIdx { bool valid; int n, x, y; size_t r; };
__device__ __forceinline__ PixIdx pix_index(int N, int H, int W, int threads)
{
    PixIdx i;
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * threads + threadIdx.x;
    i.valid = p < (size_t)N * HW;
    const size_t pc = i.valid ? p : 0;
    i.n = (int)(pc / HW);
    i.r = pc - (size_t)i.n * HW;
    i.y = (int)(i.r / W);
    i.x = (int)(i.r - (size_t)i.y * W);
    return i;
}

// ---- calc_drm_rife / calc_drm_rife_auxiliary ---------------------------------------
template <bool SOFT>
__global__ void __launch_bounds__(kDrmThreads)
drm_rife_scatter_kernel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                        const float* __restrict__ metric10, const float* __restrict__ metric12,
                        float* __restrict__ acc, int N, int H, int W, DrmSeq s, int want01, int want12)
{
    const size_t HW = (size_t)H * W;
    const PixIdx i = pix_index(N, H, W, kDrmThreads);
    const int lane = threadIdx.x & 31;
    const DrmPix px = drm_pixel<true>(flow10, flow12, i.n, i.r, HW, s);
    float* acc01 = acc + (size_t)i.n * HW * 2;
    float* acc12 = acc + ((size_t)N + i.n) * HW * 2;
    if (want01) {   // drm.py:89 / :178: warp(drm_t1_unaligned, flow10 * drm_t1_unaligned)
        const float w = SOFT ? expf(metric10[(size_t)i.n * HW + i.r]) : 1.0f;
        splat2(acc01, i.x, i.y, H, W, px.f10x * px.u1, px.f10y * px.u1, px.u1, w, i.valid, i.n, lane);
    }
    if (want12) {   // drm.py:90 / :179: warp(drm_t0_unaligned, flow12 * drm_t0_unaligned)
        const float w = SOFT ? expf(metric12[(size_t)i.n * HW + i.r]) : 1.0f;
        splat2(acc12, i.x, i.y, H, W, px.f12x * px.u0, px.f12y * px.u0, px.u0, w, i.valid, i.n, lane);
    }
}

__device__ __forceinline__ float resolve_fill(float2* acc2, size_t idx, float unaligned)
{
    const float2 a = acc2[idx];
    acc2[idx] = make_float2(0.f, 0.f);
    const float den = a.y + 0.0000001f;      // softsplat.py:277-280
    const float val = a.x / den;
    const float mask = a.y / den;            // the ones-mask splat (drm.py:95-96)
    return mask < 0.999f ? unaligned : val;  // drm.py:98-102
}

__global__ void __launch_bounds__(kDrmThreads)
drm_rife_resolve_kernel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                        float* __restrict__ acc, float* __restrict__ out01, float* __restrict__ out12,
                        int N, int H, int W, DrmSeq s)
{
    const size_t HW = (size_t)H * W;
    const PixIdx i = pix_index(N, H, W, kDrmThreads);
    if (!i.valid) return;
    const DrmPix px = drm_pixel<true>(flow10, flow12, i.n, i.r, HW, s);
    float2* acc2 = reinterpret_cast<float2*>(acc);
    const size_t idx = (size_t)i.n * HW + i.r;
    if (out01) out01[idx] = resolve_fill(acc2, idx, px.u1);
    if (out12) out12[idx] = resolve_fill(acc2, (size_t)N * HW + idx, px.u0);
}

__device__ __forceinline__ float resolve_val(float ax, float ay, float unaligned)
{
    const float den = ay + 0.0000001f;       // softsplat.py:277-280
    const float val = ax / den;
    const float mask = ay / den;             // the ones-mask splat (drm.py:95-96)
    return mask < 0.999f ? unaligned : val;  // drm.py:98-102
}

// four consecutive pixels per thread (H * W % 4 == 0, 16-byte aligned planes): the same arithmetic per pixel as
// drm_rife_resolve_kernel with 16-byte accesses -- the scalar kernel ran at ~2.5 TB/s on 4- and 8-byte accesses
__global__ void __launch_bounds__(kDrmThreads)
drm_rife_resolve4_kernel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                         float* __restrict__ acc, float* __restrict__ out01, float* __restrict__ out12,
                         int N, int H, int W, DrmSeq s)
{
    const size_t HW = (size_t)H * W, Q = HW >> 2;
    const size_t i = (size_t)blockIdx.x * kDrmThreads + threadIdx.x;
    if (i >= (size_t)N * Q) return;
    const int n = (int)(i / Q);
    const size_t r = (i - (size_t)n * Q) << 2;
    const float4 ax = *reinterpret_cast<const float4*>(flow10 + ((size_t)n * 2) * HW + r);
    const float4 ay = *reinterpret_cast<const float4*>(flow10 + ((size_t)n * 2 + 1) * HW + r);
    const float4 bx = *reinterpret_cast<const float4*>(flow12 + ((size_t)n * 2) * HW + r);
    const float4 by = *reinterpret_cast<const float4*>(flow12 + ((size_t)n * 2 + 1) * HW + r);
    const float f10x[4] = {ax.x, ax.y, ax.z, ax.w}, f10y[4] = {ay.x, ay.y, ay.z, ay.w};
    const float f12x[4] = {bx.x, bx.y, bx.z, bx.w}, f12y[4] = {by.x, by.y, by.z, by.w};
    float u0[4], u1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const DrmPix px = drm_pixel_vals<true>(f10x[k], f10y[k], f12x[k], f12y[k], s);
        u0[k] = px.u0; u1[k] = px.u1;
    }
    const size_t idx = (size_t)n * HW + r;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        float* out = m == 0 ? out01 : out12;
        if (!out) continue;
        const float* un = m == 0 ? u1 : u0;
        float4* a4 = reinterpret_cast<float4*>(acc + ((size_t)m * N * HW + idx) * 2);     // float2 per pixel
        const float4 p01 = a4[0], p23 = a4[1];
        a4[0] = zero; a4[1] = zero;
        float4 o;
        o.x = resolve_val(p01.x, p01.y, un[0]);
        o.y = resolve_val(p01.z, p01.w, un[1]);
        o.z = resolve_val(p23.x, p23.y, un[2]);
        o.w = resolve_val(p23.z, p23.w, un[3]);
        *reinterpret_cast<float4*>(out + idx) = o;
    }
}

// ---- calc_drm_gmfss ------------------------------------------------------------------
template <bool SOFT>
__global__ void __launch_bounds__(kDrmThreads)
drm_gmfss_scatter_kernel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                         const float* __restrict__ metric10, const float* __restrict__ metric12,
                         float* __restrict__ acc, float* __restrict__ drm1t_t01, float* __restrict__ drm1t_t12,
                         int N, int H, int W, DrmSeq s, int want0, int want2)
{
    const size_t HW = (size_t)H * W;
    const PixIdx i = pix_index(N, H, W, kDrmThreads);
    const int lane = threadIdx.x & 31;
    const DrmPix px = drm_pixel<false>(flow10, flow12, i.n, i.r, HW, s);
    const size_t idx = (size_t)i.n * HW + i.r;
    // drm1t_t01 = scaled drm12, drm1t_t12 = scaled drm10 (drm.py:121-128), returned unaligned
    if (i.valid && drm1t_t01) drm1t_t01[idx] = px.u1;
    if (i.valid && drm1t_t12) drm1t_t12[idx] = px.u0;
    float* acc0 = acc + (size_t)i.n * HW * 2;
    float* acc2 = acc + ((size_t)N + i.n) * HW * 2;
    if (want0) {   // drm.py:132: warp(1 - drm1t_t01, flow10, metric10)
        const float w = SOFT ? expf(metric10[idx]) : 1.0f;
        splat2(acc0, i.x, i.y, H, W, px.f10x, px.f10y, 1.0f - px.u1, w, i.valid, i.n, lane);
    }
    if (want2) {   // drm.py:133: warp(1 - drm1t_t12, flow12, metric12)
        const float w = SOFT ? expf(metric12[idx]) : 1.0f;
        splat2(acc2, i.x, i.y, H, W, px.f12x, px.f12y, 1.0f - px.u0, w, i.valid, i.n, lane);
    }
}

// The reference builds its ones mask from the WARPED map (`drm0t_t01.clone() * 0 + 1`,
// drm.py:136), so a NaN in the un-filled drm0t_t01 at pixel i (0/0 distances)
// makes the ones-mask source at i NaN for BOTH mask splats (:139-140): every in-bounds corner
// of pixel i's footprint under flow10 (resp. flow12) gets a NaN mask, and NaN < 0.999 is
// false, so those pixels are NOT hole-filled.  Reproduced with a poison map [2][N*H*W]
// (part of the zero-invariant workspace): pass 1 marks, pass 2 resolves and clears.
__device__ __forceinline__ void poison(float* __restrict__ pois, int x, int y, int H, int W, float fx, float fy)
{
    const Footprint f = footprint(x, y, fx, fy);
    if (!f.ok) return;
    const long long q = (long long)f.y0 * W + f.x0;
    const bool inx0 = f.x0 >= 0 && f.x0 < W, inx1 = f.x0 + 1 >= 0 && f.x0 + 1 < W;
    const bool iny0 = f.y0 >= 0 && f.y0 < H, iny1 = f.y0 + 1 >= 0 && f.y0 + 1 < H;
    if (inx0 && iny0) pois[q] = 1.0f;
    if (inx1 && iny0) pois[q + 1] = 1.0f;
    if (inx0 && iny1) pois[q + W] = 1.0f;
    if (inx1 && iny1) pois[q + W + 1] = 1.0f;
}

__global__ void __launch_bounds__(kDrmThreads)
drm_gmfss_poison_kernel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                        const float* __restrict__ acc, float* __restrict__ pois, int N, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const PixIdx i = pix_index(N, H, W, kDrmThreads);
    if (!i.valid) return;
    const size_t idx = (size_t)i.n * HW + i.r;
    const float v = acc[idx * 2] / (acc[idx * 2 + 1] + 0.0000001f);   // un-filled drm0t_t01
    const float ones = v * 0.0f + 1.0f;                                // drm.py:136
    if (ones != ones) {
        poison(pois + (size_t)i.n * HW, i.x, i.y, H, W,
               flow10[((size_t)i.n * 2) * HW + i.r], flow10[((size_t)i.n * 2 + 1) * HW + i.r]);
        poison(pois + ((size_t)N + i.n) * HW, i.x, i.y, H, W,
               flow12[((size_t)i.n * 2) * HW + i.r], flow12[((size_t)i.n * 2 + 1) * HW + i.r]);
    }
}

__global__ void __launch_bounds__(kDrmThreads)
drm_gmfss_resolve_kernel(const float* __restrict__ flow10, const float* __restrict__ flow12,
                         float* __restrict__ acc, float* __restrict__ pois,
                         float* __restrict__ out0, float* __restrict__ out2, int N, int H, int W, DrmSeq s)
{
    const size_t HW = (size_t)H * W;
    const PixIdx i = pix_index(N, H, W, kDrmThreads);
    if (!i.valid) return;
    const DrmPix px = drm_pixel<false>(flow10, flow12, i.n, i.r, HW, s);
    const size_t idx = (size_t)i.n * HW + i.r;
    float2* acc2 = reinterpret_cast<float2*>(acc);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        float* out = side == 0 ? out0 : out2;
        const size_t j = (size_t)side * N * HW + idx;
        const float2 a = acc2[j];
        const float po = pois[j];
        acc2[j] = make_float2(0.f, 0.f);
        pois[j] = 0.0f;
        if (!out) continue;
        const float den = a.y + 0.0000001f;
        const float val = a.x / den;
        const float mask = po != 0.0f ? __int_as_float(0x7fc00000) : a.y / den;
        const float un = 1.0f - (side == 0 ? px.u1 : px.u0);   // drm.py:123-124
        out[idx] = mask < 0.999f ? un : val;                    // drm.py:143-148
    }
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_get_drm_t_f32(const float* drm, double t, double precision, float* out, size_t n, void* stream)
{
    if (n == 0) return DRBA_OK;
    if (!drm || !out || !(precision > 0.0)) return DRBA_E_ARG;
    DrmSeq s;
    const int rc = build_seq(t, precision, &s);
    if (rc != DRBA_OK) return rc;
    get_drm_t_kernel<<<cdiv(n, kDrmThreads), kDrmThreads, 0, as_stream(stream)>>>(drm, out, n, s);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

size_t drba_drm_workspace_bytes(int N, int H, int W)
{
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    // two (value, weight) accumulators + (gmfss only) two poison maps
    return (size_t)N * H * W * (16 + 8);
}

static int make_seq(double t, int linear, DrmSeq* s)
{
    if (linear) { s->bits = 0; s->len = -1; s->t = (float)t; return DRBA_OK; }
    return build_seq(t, 1e-3, s);   // get_drm_t's default precision (drm.py:10)
}

int drba_drm_rife_f32(double t, const float* flow10, const float* flow12,
                      const float* metric10, const float* metric12, int linear,
                      float* out_t01, float* out_t12, int N, int H, int W,
                      void* ws, size_t ws_bytes, void* stream)
{
    if (N < 0 || H < 0 || W < 0) return DRBA_E_ARG;
    if ((size_t)N * H * W == 0 || (!out_t01 && !out_t12)) return DRBA_OK;
    if (!flow10 || !flow12) return DRBA_E_ARG;
    if (!ws || ws_bytes < (size_t)N * H * W * 16) return DRBA_E_WORKSPACE;
    if (!aligned16(ws)) return DRBA_E_ALIGN;
    DrmSeq s;
    const int rc = make_seq(t, linear, &s);
    if (rc != DRBA_OK) return rc;
    const bool soft = metric10 && metric12;   // drm.py:173
    cudaStream_t st = as_stream(stream);
    const unsigned grid = cdiv((size_t)N * H * W, kDrmThreads);
    if (soft)
        drm_rife_scatter_kernel<true><<<grid, kDrmThreads, 0, st>>>(flow10, flow12, metric10, metric12, (float*)ws,
                                                                     N, H, W, s, out_t01 != nullptr, out_t12 != nullptr);
    else
        drm_rife_scatter_kernel<false><<<grid, kDrmThreads, 0, st>>>(flow10, flow12, nullptr, nullptr, (float*)ws,
                                                                      N, H, W, s, out_t01 != nullptr, out_t12 != nullptr);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    const bool vec4 = ((size_t)H * W) % 4 == 0 && aligned16(flow10) && aligned16(flow12) && aligned16(ws) &&
                      (!out_t01 || aligned16(out_t01)) && (!out_t12 || aligned16(out_t12));
    if (vec4)
        drm_rife_resolve4_kernel<<<cdiv((size_t)N * H * W / 4, kDrmThreads), kDrmThreads, 0, st>>>(flow10, flow12, (float*)ws, out_t01, out_t12, N, H, W, s);
    else
        drm_rife_resolve_kernel<<<grid, kDrmThreads, 0, st>>>(flow10, flow12, (float*)ws, out_t01, out_t12, N, H, W, s);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_drm_gmfss_f32(double t, const float* flow10, const float* flow12,
                       const float* metric10, const float* metric12, int linear,
                       float* drm0t_t01, float* drm1t_t01, float* drm1t_t12, float* drm2t_t12,
                       int N, int H, int W, void* ws, size_t ws_bytes, void* stream)
{
    if (N < 0 || H < 0 || W < 0) return DRBA_E_ARG;
    if ((size_t)N * H * W == 0) return DRBA_OK;
    if (!flow10 || !flow12) return DRBA_E_ARG;
    const size_t px = (size_t)N * H * W;
    if (!ws || ws_bytes < px * 24) return DRBA_E_WORKSPACE;
    if (!aligned16(ws)) return DRBA_E_ALIGN;
    DrmSeq s;
    const int rc = make_seq(t, linear, &s);
    if (rc != DRBA_OK) return rc;
    const bool soft = metric10 && metric12;   // drm.py:119
    cudaStream_t st = as_stream(stream);
    const unsigned grid = cdiv(px, kDrmThreads);
    float* acc = (float*)ws;          // [2][px][2]
    float* pois = acc + px * 4;       // [2][px]
    const bool aligned_wanted = drm0t_t01 || drm2t_t12;
    // side 0 is always splatted when any aligned map is wanted: its NaN pattern drives both masks
    const int want0 = aligned_wanted ? 1 : 0, want2 = drm2t_t12 ? 1 : 0;
    if (soft)
        drm_gmfss_scatter_kernel<true><<<grid, kDrmThreads, 0, st>>>(flow10, flow12, metric10, metric12, acc,
                                                                      drm1t_t01, drm1t_t12, N, H, W, s, want0, want2);
    else
        drm_gmfss_scatter_kernel<false><<<grid, kDrmThreads, 0, st>>>(flow10, flow12, nullptr, nullptr, acc,
                                                                       drm1t_t01, drm1t_t12, N, H, W, s, want0, want2);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    if (!aligned_wanted) return DRBA_OK;
    drm_gmfss_poison_kernel<<<grid, kDrmThreads, 0, st>>>(flow10, flow12, acc, pois, N, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    drm_gmfss_resolve_kernel<<<grid, kDrmThreads, 0, st>>>(flow10, flow12, acc, pois, drm0t_t01, drm2t_t12, N, H, W, s);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
