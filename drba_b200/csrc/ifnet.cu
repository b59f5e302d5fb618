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
#include "ifnet_common.cuh"

namespace drba {

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
    if (!tmp || layout < 0 || layout > 2 || s <= 0 || H % s != 0 || W % s != 0) return DRBA_E_ARG;
    if (layout == 0 && ((H / s) % 2 != 0 || (W / s) % 2 != 0)) return DRBA_E_ARG;
    if (layout >= 1 && !aligned16(tmp)) return DRBA_E_ALIGN;
    return DRBA_OK;
}

extern "C" {

static int assemble_impl(const float* img0, const float* img1, const void* f0, const void* f1, int feat_dtype,
                         const float* timestep, float timestep_scalar,
                         const float* flow, const float* tmp_prev, int tmp_layout, int s_prev,
                         void* out, int out_dtype, int out_cstride, int H, int W, int s, void* stream,
                         const float* term0, int s_term0, const float* term1, int s_term1);

int drba_ifnet_assemble(const float* img0, const float* img1, const void* f0, const void* f1, int feat_dtype,
                        const float* timestep, float timestep_scalar,
                        const float* flow, const float* tmp_prev, int tmp_layout, int s_prev,
                        void* out, int out_dtype, int out_cstride, int H, int W, int s, void* stream)
{
    return assemble_impl(img0, img1, f0, f1, feat_dtype, timestep, timestep_scalar, flow, tmp_prev, tmp_layout, s_prev,
                         out, out_dtype, out_cstride, H, W, s, stream, nullptr, 0, nullptr, 0);
}

int drba_ifnet_assemble_terms(const float* img0, const float* img1, const void* f0, const void* f1,
                              const float* timestep, float timestep_scalar,
                              const float* term0, int s_term0, const float* term1, int s_term1,
                              const float* tmp_prev, int s_prev, void* out, int H, int W, int s, void* stream)
{
    if (!term0 || s_term0 <= 0 || (term1 && s_term1 <= 0)) return DRBA_E_ARG;
    int rc = tmp_ok(term0, 1, H, W, s_term0);
    if (rc == DRBA_OK && term1) rc = tmp_ok(term1, 1, H, W, s_term1);
    if (rc != DRBA_OK) return rc;
    return assemble_impl(img0, img1, f0, f1, DRBA_F16, timestep, timestep_scalar, nullptr, tmp_prev, 1, s_prev,
                         out, DRBA_F16, 64, H, W, s, stream, term0, s_term0, term1, s_term1);
}

int drba_ifnet_flow_sum(const float* tmp0, int s0, const float* tmp1, int s1, const float* tmp2, int s2, int nterms,
                        float* flow, int H, int W, void* stream)
{
    if (H <= 0 || W <= 0 || !flow || nterms < 1 || nterms > 3) return DRBA_E_ARG;
    if (!aligned16(flow)) return DRBA_E_ALIGN;
    const float* tp[3] = {tmp0, tmp1, tmp2};
    const int ss[3] = {s0, s1, s2};
    Tmp13 t[3];
    for (int k = 0; k < 3; ++k) {
        t[k].p = tp[0]; t[k].s = 1; t[k].h13 = H; t[k].w13 = W; t[k].pitch = 16;
        if (k < nterms) {
            const int rc = tmp_ok(tp[k], 1, H, W, ss[k]);
            if (rc != DRBA_OK) return rc;
            t[k].p = tp[k]; t[k].s = ss[k]; t[k].h13 = H / ss[k]; t[k].w13 = W / ss[k];
        }
    }
    launch_flow_sum_tc(t[0], t[1], t[2], nterms, flow, H, W, as_stream(stream));
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

static int assemble_impl(const float* img0, const float* img1, const void* f0, const void* f1, int feat_dtype,
                         const float* timestep, float timestep_scalar,
                         const float* flow, const float* tmp_prev, int tmp_layout, int s_prev,
                         void* out, int out_dtype, int out_cstride, int H, int W, int s, void* stream,
                         const float* term0, int s_term0, const float* term1, int s_term1)
{
    if (H <= 0 || W <= 0 || s <= 0 || (s & (s - 1)) != 0 || H % s != 0 || W % s != 0) return DRBA_E_ARG;
    if (!img0 || !img1 || !f0 || !f1 || !out) return DRBA_E_ARG;
    if (feat_dtype != DRBA_F32 && feat_dtype != DRBA_F16) return DRBA_E_ARG;
    if (out_dtype != DRBA_F32 && out_dtype != DRBA_F16) return DRBA_E_ARG;
    const bool has_flow = flow != nullptr || term0 != nullptr;
    if (out_dtype == DRBA_F16 && (out_cstride < (has_flow ? 64 : 48) || out_cstride % 8 != 0)) return DRBA_E_ARG;
    if (!aligned16(f0) || !aligned16(f1) || !aligned16(out) || (flow && !aligned16(flow))) return DRBA_E_ALIGN;
    if (has_flow) {
        if (tmp_layout == 2) return DRBA_E_ARG;       // the 8-channel form carries flow + mask only
        const int rc = tmp_ok(tmp_prev, tmp_layout, H, W, s_prev);
        if (rc != DRBA_OK) return rc;
    }
    AssembleParams p;
    p.img0 = img0; p.img1 = img1; p.f0 = f0; p.f1 = f1;
    p.timestep = timestep; p.timestep_scalar = timestep_scalar; p.flow = flow;
    p.prev.p = tmp_prev; p.prev.s = has_flow ? s_prev : 1; p.prev.h13 = has_flow ? H / s_prev : 1; p.prev.w13 = has_flow ? W / s_prev : 1; p.prev.pitch = 16;
    p.nfterms = term0 ? (term1 ? 2 : 1) : 0;
    p.fterm[0].p = term0; p.fterm[0].s = term0 ? s_term0 : 1; p.fterm[0].h13 = term0 ? H / s_term0 : 1; p.fterm[0].w13 = term0 ? W / s_term0 : 1; p.fterm[0].pitch = 16;
    p.fterm[1].p = term1; p.fterm[1].s = term1 ? s_term1 : 1; p.fterm[1].h13 = term1 ? H / s_term1 : 1; p.fterm[1].w13 = term1 ? W / s_term1 : 1; p.fterm[1].pitch = 16;
    p.out = out; p.out_cstride = out_cstride; p.H = H; p.W = W; p.s = s; p.h = H / s; p.w = W / s;
    const unsigned grid = cdiv((size_t)p.h * p.w, 32);
    cudaStream_t st = as_stream(stream);
    const int key = (feat_dtype == DRBA_F16 ? 4 : 0) | (out_dtype == DRBA_F16 ? 2 : 0) | (tmp_layout == 1 ? 1 : 0);
    if (key == 7 && has_flow && out_cstride == 64 && (term0 || !env_flag("DRBA_ASSEMBLE_V1"))) {
        launch_assemble_tc(p, st);
        DRBA_RETURN_IF_LAUNCH_FAILED();
        return DRBA_OK;
    }
    if (term0) return DRBA_E_UNSUPPORTED;      // flow terms: tensor-core engine layout only
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
    Tmp13 t; t.p = tmp; t.s = s; t.h13 = H / s; t.w13 = W / s; t.pitch = tmp_layout == 2 ? 8 : 16;
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
    Tmp13 t; t.p = tmp; t.s = s; t.h13 = H / s; t.w13 = W / s; t.pitch = tmp_layout == 2 ? 8 : 16;
    const unsigned grid = cdiv((size_t)H * W, kIfThreads);
    if (tmp_layout == 0) ifnet_blend_kernel<0><<<grid, kIfThreads, 0, as_stream(stream)>>>(img0, img1, flow, t, out, H, W);
    else ifnet_blend_kernel<1><<<grid, kIfThreads, 0, as_stream(stream)>>>(img0, img1, flow, t, out, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
