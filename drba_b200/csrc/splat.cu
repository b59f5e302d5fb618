// Forward (summation) splat for sm_100a.
//
// Replaces models/softsplat/softsplat.py:248-293 (mode prep, normalisation) and the
// CuPy kernel `softsplat_out` (:306-357).  Arithmetic per SURVEY.md appendix A.1.
//
// Design (HBM-bound op; DESIGN.md "softsplat"):
//  * the accumulator is NOT the output tensor: it is a workspace laid out as
//    [N][group][H*W][4] floats (4 channels interleaved per pixel) so that one corner
//    of one source pixel is ONE `red.global.add.v4.f32` (REDG.F32x4) instead of four
//    scalar atomics; the weight channel (1 / m / exp m) rides in the same groups;
//  * lanes of a warp hold x-adjacent source pixels; where the flow is locally smooth
//    lane i's east corners hit the same addresses as lane i+1's west corners, so the
//    east contributions are handed over with a shuffle and only 2 reds per group are
//    issued (warp-aggregated atomics);
//  * the workspace is small enough to stay in the 126 MB L2 (channels are processed
//    in chunks), so atomics never reach HBM; the resolve pass reads it from L2,
//    divides, writes the NCHW output once, and re-zeroes the workspace for the next
//    call.  HBM traffic = read in/flow/metric once + write out once.
#include "common.cuh"

namespace drba {

constexpr int kSplatThreads = 256;
constexpr size_t kSplatL2Budget = 32u << 20;  // accumulator bytes kept L2-resident per chunk

template <int MODE, int VARIANT>
__global__ void __launch_bounds__(kSplatThreads)
splat_scatter_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                     const float* __restrict__ metric, float* __restrict__ acc,
                     int N, int C, int H, int W, int c0, int cc, int ngroups)
{
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)N * HW;
    const size_t p = (size_t)blockIdx.x * kSplatThreads + threadIdx.x;
    const bool valid = p < total;
    const size_t pc = valid ? p : 0;
    const int n = (int)(pc / HW);
    const size_t r = pc - (size_t)n * HW;
    const int y = (int)(r / W);
    const int x = (int)(r - (size_t)y * W);

    const float fx = flow[((size_t)n * 2) * HW + r];
    const float fy = flow[((size_t)n * 2 + 1) * HW + r];
    Footprint f = footprint(x, y, fx, fy);
    f.ok = f.ok && valid;

    float wgt = 1.0f;
    if (MODE == DRBA_SPLAT_LINEAR) wgt = metric[(size_t)n * HW + r];
    if (MODE == DRBA_SPLAT_SOFT) wgt = expf(metric[(size_t)n * HW + r]);

    const bool inx0 = f.x0 >= 0 && f.x0 < W, inx1 = f.x0 + 1 >= 0 && f.x0 + 1 < W;
    const bool iny0 = f.y0 >= 0 && f.y0 < H, iny1 = f.y0 + 1 >= 0 && f.y0 + 1 < H;
    const bool do_nw = f.ok && inx0 && iny0, do_ne = f.ok && inx1 && iny0;
    const bool do_sw = f.ok && inx0 && iny1, do_se = f.ok && inx1 && iny1;
    const long long q = (long long)f.y0 * W + f.x0;  // north-west target (may be out of range)

    // east corners of this lane coincide with the west corners of the next lane?
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool chain_next = false, chain_prev = false;
    if (VARIANT == 0) {
        const int nx0 = __shfl_down_sync(full, f.x0, 1);
        const int ny0 = __shfl_down_sync(full, f.y0, 1);
        const int nok = __shfl_down_sync(full, (int)f.ok, 1);
        const int nn = __shfl_down_sync(full, n, 1);
        chain_next = f.ok && lane < 31 && nok && nn == n && nx0 == f.x0 + 1 && ny0 == f.y0;
        chain_prev = __shfl_up_sync(full, (int)chain_next, 1) && lane > 0;
    }

    float* accn = acc + (size_t)n * ngroups * HW * 4;
    const float* inn = in + ((size_t)n * C + c0) * HW + r;
    const int nlocal = cc + (MODE != DRBA_SPLAT_SUM ? 1 : 0);

    for (int g = 0; g < ngroups; ++g) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lc = g * 4 + k;
            float t = 0.0f;
            if (lc < cc) {
                t = valid ? inn[(size_t)lc * HW] : 0.0f;
                if (MODE == DRBA_SPLAT_LINEAR || MODE == DRBA_SPLAT_SOFT) t = t * wgt;
            } else if (lc < nlocal) {
                t = wgt;
            }
            v[k] = t;
        }
        float a[4], b[4], c[4], d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = v[k] * f.nw;
            b[k] = v[k] * f.ne;
            c[k] = v[k] * f.sw;
            d[k] = v[k] * f.se;
        }
        float* base = accn + (size_t)g * HW * 4;
        if (VARIANT == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float bp = __shfl_up_sync(full, b[k], 1);
                const float dp = __shfl_up_sync(full, d[k], 1);
                if (chain_prev) { a[k] += bp; c[k] += dp; }
            }
            if (do_nw) red_add_v4(base + q * 4, a[0], a[1], a[2], a[3]);
            if (do_sw) red_add_v4(base + (q + W) * 4, c[0], c[1], c[2], c[3]);
            if (!chain_next) {
                if (do_ne) red_add_v4(base + (q + 1) * 4, b[0], b[1], b[2], b[3]);
                if (do_se) red_add_v4(base + (q + W + 1) * 4, d[0], d[1], d[2], d[3]);
            }
        } else {
            // the reference kernel's scheme: one scalar atomic per corner and channel
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (g * 4 + k >= nlocal) break;
                if (do_nw) red_add_f32(base + q * 4 + k, a[k]);
                if (do_ne) red_add_f32(base + (q + 1) * 4 + k, b[k]);
                if (do_sw) red_add_f32(base + (q + W) * 4 + k, c[k]);
                if (do_se) red_add_f32(base + (q + W + 1) * 4 + k, d[k]);
            }
        }
    }
}

__device__ __forceinline__ float pick4(const float4& v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

template <bool NORMALISE>
__global__ void __launch_bounds__(kSplatThreads)
splat_resolve_kernel(float* __restrict__ acc, float* __restrict__ out, int N, int C, int H, int W,
                     int c0, int cc, int ngroups, int eps_mode)
{
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * kSplatThreads + threadIdx.x;
    if (p >= (size_t)N * HW) return;
    const int n = (int)(p / HW);
    const size_t r = p - (size_t)n * HW;
    float4* a4 = reinterpret_cast<float4*>(acc + (size_t)n * ngroups * HW * 4);
    float den = 1.0f;
    if (NORMALISE) {
        const float4 dv = a4[(size_t)(cc >> 2) * HW + r];
        den = splat_den(pick4(dv, cc & 3), eps_mode);
    }
    float* outn = out + ((size_t)n * C + c0) * HW + r;
    for (int g = 0; g < ngroups; ++g) {
        const float4 a = a4[(size_t)g * HW + r];
        a4[(size_t)g * HW + r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lc = g * 4 + k;
            if (lc < cc) outn[(size_t)lc * HW] = NORMALISE ? pick4(a, k) / den : pick4(a, k);
        }
    }
}

// ---- RIFE.calc_flow flow inversion (models/rife.py:59-73) --------------------------
// scatter: value channels are the flow itself -> reads 8 B/px, one v4 red per corner.
__global__ void __launch_bounds__(kSplatThreads)
invert_flow_scatter_kernel(const float* __restrict__ flow, float* __restrict__ acc, int N, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)N * HW;
    const size_t p = (size_t)blockIdx.x * kSplatThreads + threadIdx.x;
    const bool valid = p < total;
    const size_t pc = valid ? p : 0;
    const int n = (int)(pc / HW);
    const size_t r = pc - (size_t)n * HW;
    const int y = (int)(r / W);
    const int x = (int)(r - (size_t)y * W);
    const float fx = flow[((size_t)n * 2) * HW + r];
    const float fy = flow[((size_t)n * 2 + 1) * HW + r];
    Footprint f = footprint(x, y, fx, fy);
    f.ok = f.ok && valid;
    const bool inx0 = f.x0 >= 0 && f.x0 < W, inx1 = f.x0 + 1 >= 0 && f.x0 + 1 < W;
    const bool iny0 = f.y0 >= 0 && f.y0 < H, iny1 = f.y0 + 1 >= 0 && f.y0 + 1 < H;
    const long long q = (long long)f.y0 * W + f.x0;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int nx0 = __shfl_down_sync(full, f.x0, 1);
    const int ny0 = __shfl_down_sync(full, f.y0, 1);
    const int nok = __shfl_down_sync(full, (int)f.ok, 1);
    const int nn = __shfl_down_sync(full, n, 1);
    const bool chain_next = f.ok && lane < 31 && nok && nn == n && nx0 == f.x0 + 1 && ny0 == f.y0;
    const bool chain_prev = __shfl_up_sync(full, (int)chain_next, 1) && lane > 0;
    float a[3] = {fx * f.nw, fy * f.nw, f.nw};
    float b[3] = {fx * f.ne, fy * f.ne, f.ne};
    float c[3] = {fx * f.sw, fy * f.sw, f.sw};
    float d[3] = {fx * f.se, fy * f.se, f.se};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float bp = __shfl_up_sync(full, b[k], 1);
        const float dp = __shfl_up_sync(full, d[k], 1);
        if (chain_prev) { a[k] += bp; c[k] += dp; }
    }
    float* base = acc + (size_t)n * HW * 4;
    if (f.ok && inx0 && iny0) red_add_v4(base + q * 4, a[0], a[1], a[2], 0.f);
    if (f.ok && inx0 && iny1) red_add_v4(base + (q + W) * 4, c[0], c[1], c[2], 0.f);
    if (!chain_next) {
        if (f.ok && inx1 && iny0) red_add_v4(base + (q + 1) * 4, b[0], b[1], b[2], 0.f);
        if (f.ok && inx1 && iny1) red_add_v4(base + (q + W + 1) * 4, d[0], d[1], d[2], 0.f);
    }
}

__global__ void __launch_bounds__(kSplatThreads)
invert_flow_resolve_kernel(float* __restrict__ acc, float* __restrict__ out, int N, int H, int W)
{
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * kSplatThreads + threadIdx.x;
    if (p >= (size_t)N * HW) return;
    const int n = (int)(p / HW);
    const size_t r = p - (size_t)n * HW;
    float4* a4 = reinterpret_cast<float4*>(acc);
    const float4 a = a4[p];
    a4[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float den = a.z + 0.0000001f;
    const float mask = a.z / den;                       // splat_avg(ones) (rife.py:63-64)
    const float big = (float)(H > W ? H : W);           // rife.py:69-70
    float ox = -1.0f * (a.x / den), oy = -1.0f * (a.y / den);
    if (mask < 0.999f) { ox = big; oy = big; }
    out[((size_t)n * 2) * HW + r] = ox * 2.0f;          // rife.py:72-73
    out[((size_t)n * 2 + 1) * HW + r] = oy * 2.0f;
}



// csrc/splat_gather.cu
size_t splat_gather_workspace_bytes(int N, int H, int W);
int splat_gather_launch(const float* in, const float* flow, const float* metric, float* out,
                        int N, int C, int H, int W, int mode, int eps_mode, void* ws, cudaStream_t st, bool tile);
// the list-building cost of the gather path is independent of C: it wins once the scatter path would
// need three or more 4-channel accumulator groups
constexpr int kGatherMinChannels = 9;

}  // namespace drba

using namespace drba;

namespace drba {
// four consecutive pixels per thread (H * W % 4 == 0): same arithmetic as invert_flow_resolve_kernel, 16-byte stores
__global__ void __launch_bounds__(kSplatThreads)
invert_flow_resolve4_kernel(float* __restrict__ acc, float* __restrict__ out, int N, int H, int W)
{
    const size_t HW = (size_t)H * W, Q = HW >> 2;
    const size_t i = (size_t)blockIdx.x * kSplatThreads + threadIdx.x;
    if (i >= (size_t)N * Q) return;
    const int n = (int)(i / Q);
    const size_t r = (i - (size_t)n * Q) << 2;
    float4* a4 = reinterpret_cast<float4*>(acc) + (size_t)n * HW + r;
    const float big = (float)(H > W ? H : W);           // rife.py:69-70
    float ox[4], oy[4];
    float4 a[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = a4[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float den = a[k].z + 0.0000001f;
        const float mask = a[k].z / den;                // splat_avg(ones) (rife.py:63-64)
        float x = -1.0f * (a[k].x / den), y = -1.0f * (a[k].y / den);
        if (mask < 0.999f) { x = big; y = big; }
        ox[k] = x * 2.0f; oy[k] = y * 2.0f;             // rife.py:72-73
    }
    *reinterpret_cast<float4*>(out + ((size_t)n * 2) * HW + r) = make_float4(ox[0], ox[1], ox[2], ox[3]);
    *reinterpret_cast<float4*>(out + ((size_t)n * 2 + 1) * HW + r) = make_float4(oy[0], oy[1], oy[2], oy[3]);
}
}  // namespace drba

extern "C" {

size_t drba_softsplat_workspace_bytes(int N, int C, int H, int W, int mode)
{
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    const size_t group_bytes = (size_t)N * H * W * 16;
    const size_t groups_total = ((size_t)C + (mode != DRBA_SPLAT_SUM ? 1 : 0) + 3) / 4;
    size_t groups_cap = kSplatL2Budget / group_bytes;
    if (groups_cap < 1) groups_cap = 1;
    const size_t scatter = group_bytes * (groups_total < groups_cap ? groups_total : groups_cap);
    const size_t gather = C + (mode != DRBA_SPLAT_SUM ? 1 : 0) >= kGatherMinChannels ? splat_gather_workspace_bytes(N, H, W) : 0;
    return scatter > gather ? scatter : gather;
}

int drba_softsplat_f32_variant(const float* in, const float* flow, const float* metric, float* out,
                               int N, int C, int H, int W, int mode, int eps_mode,
                               void* ws, size_t ws_bytes, int variant, void* stream)
{
    if (N < 0 || C < 0 || H < 0 || W < 0) return DRBA_E_ARG;
    if (mode < DRBA_SPLAT_SUM || mode > DRBA_SPLAT_SOFT) return DRBA_E_ARG;
    if (eps_mode < DRBA_EPS_ADD || eps_mode > DRBA_EPS_NONE) return DRBA_E_ARG;
    if (variant < 0 || variant > 4) return DRBA_E_ARG;
    if ((size_t)N * C * H * W == 0) return DRBA_OK;
    if (!in || !flow || !out) return DRBA_E_ARG;
    if ((mode == DRBA_SPLAT_LINEAR || mode == DRBA_SPLAT_SOFT) && !metric) return DRBA_E_ARG;
    const size_t group_bytes = (size_t)N * H * W * 16;
    if (!ws || ws_bytes < group_bytes) return DRBA_E_WORKSPACE;
    if (!aligned16(ws)) return DRBA_E_ALIGN;
    const int has_w = mode != DRBA_SPLAT_SUM ? 1 : 0;
    // variant: 0 = automatic, 1 = scalar atomics (the reference kernel's scheme), 2 = vector-red scatter,
    //          3 = owner-computes gather, per-target loads (csrc/splat_gather.cu; what variant 0 picks for
    //          C + weight >= 9), 4 = the same with TMA-staged source tiles (measured slower on B200: kept selectable)
    if (variant == 3 || variant == 4 || (variant == 0 && C + has_w >= kGatherMinChannels)) {
        const size_t need = splat_gather_workspace_bytes(N, H, W);
        if (ws_bytes >= need) {
            const int rc = splat_gather_launch(in, flow, metric, out, N, C, H, W, mode, eps_mode, ws, as_stream(stream), variant == 4);
            if (rc != DRBA_E_UNSUPPORTED) return rc;
        } else if (variant == 3 || variant == 4) {
            return DRBA_E_WORKSPACE;
        }
    }
    if (variant == 2) variant = 0;
    if (variant == 3 || variant == 4) variant = 0;
    const size_t groups_fit = ws_bytes / group_bytes;
    const int cc_max = (int)(groups_fit * 4 > (size_t)(C + has_w) ? (size_t)C : groups_fit * 4 - has_w);
    cudaStream_t st = as_stream(stream);
    const unsigned grid = cdiv((size_t)N * H * W, kSplatThreads);
    for (int c0 = 0; c0 < C; c0 += cc_max) {
        const int cc = C - c0 < cc_max ? C - c0 : cc_max;
        const int ngroups = (cc + has_w + 3) / 4;
#define LAUNCH_SCATTER(M)                                                                              \
        if (variant == 0)                                                                              \
            splat_scatter_kernel<M, 0><<<grid, kSplatThreads, 0, st>>>(in, flow, metric, (float*)ws, N, C, H, W, c0, cc, ngroups); \
        else                                                                                           \
            splat_scatter_kernel<M, 1><<<grid, kSplatThreads, 0, st>>>(in, flow, metric, (float*)ws, N, C, H, W, c0, cc, ngroups);
        switch (mode) {
            case DRBA_SPLAT_SUM: LAUNCH_SCATTER(DRBA_SPLAT_SUM) break;
            case DRBA_SPLAT_AVG: LAUNCH_SCATTER(DRBA_SPLAT_AVG) break;
            case DRBA_SPLAT_LINEAR: LAUNCH_SCATTER(DRBA_SPLAT_LINEAR) break;
            default: LAUNCH_SCATTER(DRBA_SPLAT_SOFT) break;
        }
#undef LAUNCH_SCATTER
        DRBA_RETURN_IF_LAUNCH_FAILED();
        if (has_w)
            splat_resolve_kernel<true><<<grid, kSplatThreads, 0, st>>>((float*)ws, out, N, C, H, W, c0, cc, ngroups, eps_mode);
        else
            splat_resolve_kernel<false><<<grid, kSplatThreads, 0, st>>>((float*)ws, out, N, C, H, W, c0, cc, ngroups, eps_mode);
        DRBA_RETURN_IF_LAUNCH_FAILED();
    }
    return DRBA_OK;
}

int drba_softsplat_f32(const float* in, const float* flow, const float* metric, float* out,
                       int N, int C, int H, int W, int mode, int eps_mode,
                       void* ws, size_t ws_bytes, void* stream)
{
    return drba_softsplat_f32_variant(in, flow, metric, out, N, C, H, W, mode, eps_mode, ws, ws_bytes, 0, stream);
}

size_t drba_rife_invert_flow_workspace_bytes(int N, int H, int W)
{
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)N * H * W * 16;
}

int drba_rife_invert_flow_f32(const float* flow_t0, float* out, int N, int H, int W,
                              void* ws, size_t ws_bytes, void* stream)
{
    if (N < 0 || H < 0 || W < 0) return DRBA_E_ARG;
    if ((size_t)N * H * W == 0) return DRBA_OK;
    if (!flow_t0 || !out) return DRBA_E_ARG;
    if (!ws || ws_bytes < (size_t)N * H * W * 16) return DRBA_E_WORKSPACE;
    if (!aligned16(ws)) return DRBA_E_ALIGN;
    cudaStream_t st = as_stream(stream);
    const unsigned grid = cdiv((size_t)N * H * W, kSplatThreads);
    invert_flow_scatter_kernel<<<grid, kSplatThreads, 0, st>>>(flow_t0, (float*)ws, N, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    if (((size_t)H * W) % 4 == 0 && aligned16(out))
        invert_flow_resolve4_kernel<<<cdiv((size_t)N * H * W / 4, kSplatThreads), kSplatThreads, 0, st>>>((float*)ws, out, N, H, W);
    else
        invert_flow_resolve_kernel<<<grid, kSplatThreads, 0, st>>>((float*)ws, out, N, H, W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
