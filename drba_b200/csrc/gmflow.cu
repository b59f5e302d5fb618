// Non-GEMM stages of GMFlow (models/gmflow/*) for sm_100a.  All dense contractions of GMFlow (backbone convs,
// q/k/v/merge/FFN linears, window attention QK^T and PV, global correlation, propagation scores, the
// upsampler head) run on the tcgen05 conv / batched-GEMM engine (conv_tc.cu); the kernels here are what sits
// between them: InstanceNorm, LayerNorm + residual, window (un)packing with the half-window roll of the shifted
// blocks, row softmax with the shift mask, soft-argmax readouts, the local (9x9 / 3x3) branches, feature
// warping and the convex upsampling.  Activations are token-major NHWC fp16 [B][h][w][C]; statistics,
// softmax and flows are fp32.
#include "common.cuh"

namespace drba {

constexpr int kGfThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

// ---- image normalisation (utils.py:58-70) ------------------------------------------------------------
__global__ void __launch_bounds__(kGfThreads)
normalize_img_kernel(const float* __restrict__ in, float* __restrict__ out, size_t HW)
{
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    if (i >= HW) return;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(size_t)c * HW + i] = (in[(size_t)c * HW + i] - mean[c]) / stdv[c];
}

// ---- InstanceNorm2d (affine = False, eps 1e-5) over NHWC fp16 (backbone.py:14-43) --------------------
// stats[c] = {sum, sum of squares} in double.  A thread owns a group of 8 channels (one 16 B load per pixel) and
// walks a strip of pixels; partial sums meet in shared memory, one double atomic pair per channel and block.
__global__ void __launch_bounds__(kGfThreads)
inorm_stats_kernel(const __half* __restrict__ x, int C, size_t HW, int pix_per_block, double* __restrict__ stats)
{
    const int G = C / 8;                         // channel groups
    const int R = kGfThreads / G;                // pixel lanes
    const int g = threadIdx.x % G, r = threadIdx.x / G;
    const size_t p0 = (size_t)blockIdx.x * pix_per_block;
    const size_t p1 = p0 + pix_per_block < HW ? p0 + pix_per_block : HW;
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.0f; ss[j] = 0.0f; }
    if (r < R)
        for (size_t p = p0 + r; p < p1; p += R) {
            const uint4 v = *reinterpret_cast<const uint4*>(x + p * C + g * 8);
            const __half2* hp = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(hp[j]);
                s[2 * j] += f.x; ss[2 * j] += f.x * f.x; s[2 * j + 1] += f.y; ss[2 * j + 1] += f.y * f.y;
            }
        }
    // fixed-order reduction over the R pixel lanes (deterministic); blocks meet in double precision, where the
    // order of the atomics is far below fp32 resolution
    __shared__ float sh[2][kGfThreads * 8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sh[0][(r * G + g) * 8 + j] = s[j]; sh[1][(r * G + g) * 8 + j] = ss[j]; }
    __syncthreads();
    if (threadIdx.x < C) {
        float a = 0.0f, b = 0.0f;
        for (int k = 0; k < R; ++k) { a += sh[0][k * C + threadIdx.x]; b += sh[1][k * C + threadIdx.x]; }
        atomicAdd(stats + 2 * threadIdx.x, (double)a);
        atomicAdd(stats + 2 * threadIdx.x + 1, (double)b);
    }
}

// out = relu?( [skip (raw or IN'd)] + relu?(IN(x)) ), everything NHWC fp16; a thread handles 8 channels
__global__ void __launch_bounds__(kGfThreads)
inorm_apply_kernel(const __half* __restrict__ x, const double* __restrict__ stats, int relu_x,
                   const __half* __restrict__ skip, const double* __restrict__ skip_stats, int final_relu,
                   __half* __restrict__ out, int C, size_t HW)
{
    __shared__ float s_mean[128], s_rstd[128], k_mean[128], k_rstd[128];
    if (threadIdx.x < C) {
        const double n = (double)HW;
        const int c = threadIdx.x;
        const double m = stats[2 * c] / n, var = stats[2 * c + 1] / n - m * m;
        s_mean[c] = (float)m; s_rstd[c] = rsqrtf((float)(var > 0.0 ? var : 0.0) + 1e-5f);
        if (skip_stats) {
            const double sm = skip_stats[2 * c] / n, sv2 = skip_stats[2 * c + 1] / n - sm * sm;
            k_mean[c] = (float)sm; k_rstd[c] = rsqrtf((float)(sv2 > 0.0 ? sv2 : 0.0) + 1e-5f);
        }
    }
    __syncthreads();
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    const int G = C / 8;
    if (i >= HW * (size_t)G) return;
    const int g = (int)(i % G);
    const uint4 xv = *reinterpret_cast<const uint4*>(x + i * 8);
    const __half* xh = reinterpret_cast<const __half*>(&xv);
    uint4 sv = make_uint4(0, 0, 0, 0);
    if (skip) sv = *reinterpret_cast<const uint4*>(skip + i * 8);
    const __half* sh = reinterpret_cast<const __half*>(&sv);
    uint4 ov;
    __half* oh = reinterpret_cast<__half*>(&ov);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        float v = (__half2float(xh[j]) - s_mean[c]) * s_rstd[c];
        if (relu_x) v = fmaxf(v, 0.0f);
        if (skip) {
            float sk = __half2float(sh[j]);
            if (skip_stats) sk = (sk - k_mean[c]) * k_rstd[c];
            v += sk;
        }
        if (final_relu) v = fmaxf(v, 0.0f);
        oh[j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = ov;
}

// ---- windowed sine position (utils.py:73-94) added in place ------------------------------------------
__global__ void __launch_bounds__(kGfThreads)
add_position_kernel(__half* __restrict__ x, const float* __restrict__ pos, int B, int h, int w, int wh, int ww, int C)
{
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    const size_t total = (size_t)B * h * w * C;
    if (i >= total) return;
    const int c = (int)(i % C);
    const size_t p = i / C;
    const int xx = (int)(p % w), yy = (int)((p / w) % h);
    x[i] = __float2half_rn(__half2float(x[i]) + pos[((size_t)(yy % wh) * ww + (xx % ww)) * C + c]);
}

// ---- window packing (utils.py:5-29 split_feature, transformer.py:74-84 roll) ---------------------------
// src tokens [B][h][w][C] -> dst [B*k*k][rows_pad][C] (row l = token (wy, wx) of the window) or, transposed,
// dst [B*k*k][C][rows_pad].  shift: the window grid is laid over the image rolled by (-sh, -sw).
__global__ void __launch_bounds__(kGfThreads)
window_pack_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int B, int h, int w, int C, int k,
                   int sh, int sw, int rows_pad, int transposed)
{
    const int wh = h / k, ww = w / k, Lw = wh * ww;
    const size_t total = (size_t)B * k * k * Lw * (C / 8);
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    if (i >= total) return;
    const int c8 = (int)(i % (C / 8));
    size_t r = i / (C / 8);
    const int l = (int)(r % Lw); r /= Lw;
    const int win = (int)(r % (k * k)), b = (int)(r / (k * k));
    const int wy = l / ww, wx = l - wy * ww;
    const int y = ((win / k) * wh + wy + sh) % h, x = ((win % k) * ww + wx + sw) % w;    // rolled[y'] = orig[(y' + sh) % h]
    const uint4 v = *reinterpret_cast<const uint4*>(src + (((size_t)b * h + y) * w + x) * C + c8 * 8);
    const size_t wb = (size_t)b * k * k + win;
    if (!transposed) {
        *reinterpret_cast<uint4*>(dst + (wb * rows_pad + l) * C + c8 * 8) = v;
    } else {
        const __half* hv = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[(wb * C + c8 * 8 + j) * rows_pad + l] = hv[j];
    }
}

// Transposed packing (V^T for the P V product) through a shared-memory tile: 64 tokens x C channels per block, token
// rows read as whole lines, channel rows written as 128-byte runs (the scalar version above writes 2 bytes per store,
// rows_pad apart: 1.3 ms of GMFlow's 9.5 ms at 1080p).  16-byte chunks are XOR-swizzled by the token group so that
// the column reads of the second phase spread over the banks.  Needs C % 64 == 0 and rows_pad % 8 == 0.
__global__ void __launch_bounds__(kGfThreads)
window_pack_t_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int h, int w, int C, int k,
                     int sh, int sw, int rows_pad)
{
    extern __shared__ __align__(16) __half tile[];        // [64][C]
    const int wh = h / k, ww = w / k, Lw = wh * ww;
    const int l0 = blockIdx.x * 64;
    const int wb = blockIdx.y, win = wb % (k * k), b = wb / (k * k);
    const int c8n = C >> 3;
    for (int e = threadIdx.x; e < 64 * c8n; e += kGfThreads) {
        const int tok = e / c8n, c8 = e - tok * c8n;
        const int l = l0 + tok;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (l < Lw) {
            const int wy = l / ww, wx = l - wy * ww;
            const int y = ((win / k) * wh + wy + sh) % h, x = ((win % k) * ww + wx + sw) % w;
            v = *reinterpret_cast<const uint4*>(src + (((size_t)b * h + y) * w + x) * C + c8 * 8);
        }
        *reinterpret_cast<uint4*>(tile + tok * C + ((c8 ^ ((tok >> 3) & 7)) << 3)) = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * 8; e += kGfThreads) {
        const int tg = e & 7, c = e >> 3;
        const int l = l0 + tg * 8;
        if (l >= Lw) continue;
        uint4 o;
        __half* oh = reinterpret_cast<__half*>(&o);
        const int col = (((c >> 3) ^ tg) << 3) + (c & 7);
#pragma unroll
        for (int j = 0; j < 8; ++j) oh[j] = tile[(tg * 8 + j) * C + col];
        __half* d = dst + ((size_t)wb * C + c) * rows_pad + l;
        if (l + 8 <= Lw) *reinterpret_cast<uint4*>(d) = o;
        else
            for (int j = 0; j < Lw - l; ++j) d[j] = oh[j];      // rows >= Lw keep their (zero) padding
    }
}

// ---- row softmax of attention scores, in place (transformer.py:86-92) ----------------------------------
// S [nwin][Lw][ld] fp16; columns >= Lw are written as zero.  shifted: -100 is added where the query and key
// tokens lie in different regions of the rolled image (transformer.py:20-45).  One warp per row.
__device__ __forceinline__ int shift_region(int l, int win, int k, int wh, int ww, int h, int w, int sh, int sw)
{
    const int wy = l / ww, wx = l - wy * ww;
    const int y = (win / k) * wh + wy, x = (win % k) * ww + wx;      // position in the rolled image
    const int ry = y < h - wh ? 0 : (y < h - sh ? 1 : 2);
    const int rx = x < w - ww ? 0 : (x < w - sw ? 1 : 2);
    return ry * 3 + rx;
}

// grid = (ceil(Lw / 8), windows): the 8 warps of a block take 8 rows of ONE window; the region id of every key
// token (shifted blocks) is tabulated once per block; a row is read once with 16 B loads (a lane owns 8
// consecutive columns of every 256-column slab) and written once.
constexpr int kSoftmaxMaxCols = 2048;
template <int NS>      // 256-column slabs per row
__global__ void __launch_bounds__(kGfThreads)
softmax_rows_kernel(__half* __restrict__ S, int nwin_total, int Lw, int ld, int shifted, int k, int wh, int ww, int h, int w)
{
    __shared__ unsigned char region[kSoftmaxMaxCols];
    const int win_b = blockIdx.y, win = win_b % (k * k);
    const int sh = wh / 2, sw = ww / 2;
    if (shifted) {
        for (int j = threadIdx.x; j < Lw; j += kGfThreads) region[j] = (unsigned char)shift_region(j, win, k, wh, ww, h, w, sh, sw);
        __syncthreads();
    }
    const int q = blockIdx.x * (kGfThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= Lw) return;
    __half* row = S + ((size_t)win_b * Lw + q) * ld;
    const int rq = shifted ? region[q] : 0;
    float v[NS * 8];
    float mx = -3.0e38f;
#pragma unroll
    for (int s8 = 0; s8 < NS; ++s8) {
        const int j0 = s8 * 256 + lane * 8;
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (j0 < ld) raw = *reinterpret_cast<const uint4*>(row + j0);        // ld is a multiple of 16
        const __half* hv = reinterpret_cast<const __half*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = j0 + e;
            float t = -3.0e38f;
            if (j < Lw) {
                t = __half2float(hv[e]);
                if (shifted && region[j] != rq) t += -100.0f;
            }
            v[s8 * 8 + e] = t;
            mx = fmaxf(mx, t);
        }
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NS * 8; ++i) {
        v[i] = v[i] > -1.0e38f ? __expf(v[i] - mx) : 0.0f;
        sum += v[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int s8 = 0; s8 < NS; ++s8) {
        const int j0 = s8 * 256 + lane * 8;
        if (j0 < ld) {
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(v[s8 * 8 + 2 * e] * inv, v[s8 * 8 + 2 * e + 1] * inv);
            *reinterpret_cast<uint4*>(row + j0) = o;
        }
    }
}

// rows longer than 2048 columns (GMFlow inputs above 544x960 with 2 splits): three passes over the row
constexpr int kSoftmaxGenericCols = 8192;
__global__ void __launch_bounds__(kGfThreads)
softmax_rows_generic_kernel(__half* __restrict__ S, int nwin_total, int Lw, int ld, int shifted, int k, int wh, int ww, int h, int w)
{
    __shared__ unsigned char region[kSoftmaxGenericCols];
    const int win_b = blockIdx.y, win = win_b % (k * k);
    const int sh = wh / 2, sw = ww / 2;
    if (shifted) {
        for (int j = threadIdx.x; j < Lw; j += kGfThreads) region[j] = (unsigned char)shift_region(j, win, k, wh, ww, h, w, sh, sw);
        __syncthreads();
    }
    const int q = blockIdx.x * (kGfThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= Lw) return;
    __half* row = S + ((size_t)win_b * Lw + q) * ld;
    const int rq = shifted ? region[q] : 0;
    float mx = -3.0e38f;
    for (int j = lane; j < Lw; j += 32) {
        float v = __half2float(row[j]);
        if (shifted && region[j] != rq) v += -100.0f;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < Lw; j += 32) {
        float v = __half2float(row[j]);
        if (shifted && region[j] != rq) v += -100.0f;
        sum += __expf(v - mx);
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < ld; j += 32) {
        float o = 0.0f;
        if (j < Lw) {
            float v = __half2float(row[j]);
            if (shifted && region[j] != rq) v += -100.0f;
            o = __expf(v - mx) * inv;
        }
        row[j] = __float2half_rn(o);
    }
}

// ---- LayerNorm + residual (transformer.py:177-188) -----------------------------------------------------
// m: [rows][C] fp16 in WINDOW order when k > 0 (rows = (b, win, l)), else token order.
// out[token] = src[token] + LN(m[row(token)]) * gamma + beta        (cat == 0)
// cat[token] = [src[token] | LN(m) * gamma + beta]                   (cat == 1: FFN input, 2C channels)
// One warp per token, C = 128 (4 channels per lane).
__global__ void __launch_bounds__(kGfThreads)
ln_residual_kernel(const __half* __restrict__ src, const __half* __restrict__ m, const float* __restrict__ gamma,
                   const float* __restrict__ beta, __half* __restrict__ out, int B, int h, int w, int k, int sh, int sw,
                   int rows_pad, int cat)
{
    constexpr int C = 128;
    const int warp = (int)(((size_t)blockIdx.x * kGfThreads + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    const int L = h * w;
    if (warp >= B * L) return;
    const int b = warp / L, t = warp - b * L;
    size_t mrow = (size_t)warp;
    if (k > 0) {
        const int y = t / w, x = t - y * w;
        const int wh = h / k, ww = w / k;
        const int yr = (y - sh + h) % h, xr = (x - sw + w) % w;           // position in the rolled image
        const int win = (yr / wh) * k + xr / ww, l = (yr % wh) * ww + xr % ww;
        mrow = ((size_t)b * k * k + win) * rows_pad + l;
    }
    const uint2 raw = *reinterpret_cast<const uint2*>(m + mrow * C + lane * 4);
    const __half2* hp = reinterpret_cast<const __half2*>(&raw);
    const float2 a = __half22float2(hp[0]), c2 = __half22float2(hp[1]);
    float v[4] = {a.x, a.y, c2.x, c2.y};
    const float mean = warp_sum(v[0] + v[1] + v[2] + v[3]) * (1.0f / C);
    float d = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] -= mean; d += v[i] * v[i]; }
    const float rstd = rsqrtf(warp_sum(d) * (1.0f / C) + 1e-5f);
    const uint2 sraw = *reinterpret_cast<const uint2*>(src + (size_t)warp * C + lane * 4);
    const __half2* sp = reinterpret_cast<const __half2*>(&sraw);
    const float2 s0 = __half22float2(sp[0]), s1 = __half22float2(sp[1]);
    const float sv[4] = {s0.x, s0.y, s1.x, s1.y};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = v[i] * rstd * gamma[lane * 4 + i] + beta[lane * 4 + i];
    uint2 w2;
    __half2* wp = reinterpret_cast<__half2*>(&w2);
    if (!cat) {
        wp[0] = __floats2half2_rn(sv[0] + o[0], sv[1] + o[1]);
        wp[1] = __floats2half2_rn(sv[2] + o[2], sv[3] + o[3]);
        *reinterpret_cast<uint2*>(out + (size_t)warp * C + lane * 4) = w2;
    } else {
        *reinterpret_cast<uint2*>(out + (size_t)warp * 2 * C + lane * 4) = sraw;
        wp[0] = __floats2half2_rn(o[0], o[1]);
        wp[1] = __floats2half2_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(out + (size_t)warp * 2 * C + C + lane * 4) = w2;
    }
}

// ---- soft readout of a score matrix (matching.py:31-41 global matching; transformer.py:359-363) ---------
// out[row] = sum_j softmax(S[row])_j * val[j] (- own grid position when subtract_grid).  S [rows][ld] fp16,
// val [cols][2] fp32 or NULL -> val[j] = (j % w, j / w).  One warp per row.
__global__ void __launch_bounds__(kGfThreads)
soft_readout_kernel(const __half* __restrict__ S, int rows, int cols, int ld, const float* __restrict__ val, int w,
                    int subtract_grid, float scale, float* __restrict__ out /* [2][rows] planar */)
{
    const int warp = (int)(((size_t)blockIdx.x * kGfThreads + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const __half* row = S + (size_t)warp * ld;
    float mx = -3.0e38f;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, scale * __half2float(row[j]));
    mx = warp_max(mx);
    float sum = 0.0f, ax = 0.0f, ay = 0.0f;
    for (int j = lane; j < cols; j += 32) {
        const float e = __expf(scale * __half2float(row[j]) - mx);
        const float vx = val ? val[2 * j] : (float)(j % w), vy = val ? val[2 * j + 1] : (float)(j / w);
        sum += e; ax += e * vx; ay += e * vy;
    }
    sum = warp_sum(sum); ax = warp_sum(ax); ay = warp_sum(ay);
    if (lane == 0) {
        float ox = ax / sum, oy = ay / sum;
        if (subtract_grid) { ox -= (float)(warp % w); oy -= (float)(warp / w); }
        out[warp] = ox;
        out[rows + warp] = oy;
    }
}

// ---- local correlation soft-argmax (matching.py:46-89), radius r, features [h][w][128] fp16 --------------
// One warp per pixel; lanes own 4 channels each; taps outside the image get -1e4 (their features are zero).
__global__ void __launch_bounds__(kGfThreads)
local_match_kernel(const __half* __restrict__ f0, const __half* __restrict__ f1, int h, int w, int r, float* __restrict__ flow /* [2][h*w] */)
{
    constexpr int C = 128;
    const int warp = (int)(((size_t)blockIdx.x * kGfThreads + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= h * w) return;
    const int y = warp / w, x = warp - y * w;
    const uint2 qa = *reinterpret_cast<const uint2*>(f0 + (size_t)warp * C + lane * 4);
    const __half2* qp = reinterpret_cast<const __half2*>(&qa);
    const float2 q0 = __half22float2(qp[0]), q1 = __half22float2(qp[1]);
    const float scale = rsqrtf((float)C);
    float mx = -3.0e38f, sum = 0.0f, ax = 0.0f, ay = 0.0f;
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx) {
            const int yy = y + dy, xx = x + dx;
            float s;
            if (yy < 0 || yy >= h || xx < 0 || xx >= w) {
                s = -1e4f;
            } else {
                const uint2 ka = *reinterpret_cast<const uint2*>(f1 + ((size_t)yy * w + xx) * C + lane * 4);
                const __half2* kp = reinterpret_cast<const __half2*>(&ka);
                const float2 k0 = __half22float2(kp[0]), k1 = __half22float2(kp[1]);
                s = warp_sum(q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y) * scale;
            }
            // online softmax
            const float nm = fmaxf(mx, s);
            const float corr = __expf(mx - nm), e = __expf(s - nm);
            sum = sum * corr + e; ax = ax * corr + e * (float)xx; ay = ay * corr + e * (float)yy;
            mx = nm;
        }
    if (lane == 0) {
        flow[warp] = ax / sum - (float)x;
        flow[(size_t)h * w + warp] = ay / sum - (float)y;
    }
}

// ---- local-window flow propagation (transformer.py:366-409), radius 1: q, kmap [h][w][128] fp16 --------
// F.unfold pads with zeros: an outside tap has score 0 (not masked) and flow 0.
__global__ void __launch_bounds__(kGfThreads)
local_propagate_kernel(const __half* __restrict__ q, const __half* __restrict__ kmap, const float* __restrict__ flow,
                       int h, int w, float* __restrict__ out)
{
    constexpr int C = 128;
    const int warp = (int)(((size_t)blockIdx.x * kGfThreads + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= h * w) return;
    const int y = warp / w, x = warp - y * w;
    const size_t HW = (size_t)h * w;
    const uint2 qa = *reinterpret_cast<const uint2*>(q + (size_t)warp * C + lane * 4);
    const __half2* qp = reinterpret_cast<const __half2*>(&qa);
    const float2 q0 = __half22float2(qp[0]), q1 = __half22float2(qp[1]);
    const float scale = rsqrtf((float)C);
    float s[9], fx[9], fy[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        s[t] = 0.0f; fx[t] = 0.0f; fy[t] = 0.0f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            const uint2 ka = *reinterpret_cast<const uint2*>(kmap + ((size_t)yy * w + xx) * C + lane * 4);
            const __half2* kp = reinterpret_cast<const __half2*>(&ka);
            const float2 k0 = __half22float2(kp[0]), k1 = __half22float2(kp[1]);
            s[t] = warp_sum(q0.x * k0.x + q0.y * k0.y + q1.x * k1.x + q1.y * k1.y) * scale;
            fx[t] = flow[(size_t)yy * w + xx]; fy[t] = flow[HW + (size_t)yy * w + xx];
        }
    }
    float mx = s[0];
#pragma unroll
    for (int t = 1; t < 9; ++t) mx = fmaxf(mx, s[t]);
    float sum = 0.0f, ax = 0.0f, ay = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { const float e = __expf(s[t] - mx); sum += e; ax += e * fx[t]; ay += e * fy[t]; }
    if (lane == 0) { out[warp] = ax / sum; out[HW + warp] = ay / sum; }
}

// ---- feature warp for the refinement scale (gmflow.py:117-123, geometry.py:76-84): zeros padding -------
__global__ void __launch_bounds__(kGfThreads)
warp_feature_kernel(const __half* __restrict__ f, const float* __restrict__ flow, __half* __restrict__ out, int h, int w)
{
    constexpr int C = 128;
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    const size_t HW = (size_t)h * w;
    if (i >= HW * (C / 8)) return;
    const int c8 = (int)(i % (C / 8));
    const size_t p = i / (C / 8);
    const int y = (int)(p / w), x = (int)(p - (size_t)y * w);
    const float sx = (float)x + flow[p], sy = (float)y + flow[HW + p];
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float ax = sx - fx0, ay = sy - fy0;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    const float wts[4] = {(1.0f - ax) * (1.0f - ay), ax * (1.0f - ay), (1.0f - ax) * ay, ax * ay};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
        if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(f + ((size_t)yy * w + xx) * C + c8 * 8);
        const __half2* hp = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t2 = __half22float2(hp[j]); acc[2 * j] += t2.x * wts[t]; acc[2 * j + 1] += t2.y * wts[t]; }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(out + p * C + c8 * 8) = o;
}

// ---- upsampler input: cat(flow, feature) -> NHWC fp16 [h][w][144] (gmflow.py:78) ------------------------
__global__ void __launch_bounds__(kGfThreads)
upsampler_input_kernel(const float* __restrict__ flow, const __half* __restrict__ feat, __half* __restrict__ out, size_t HW)
{
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    if (i >= HW * 18) return;
    const int g = (int)(i % 18);
    const size_t p = i / 18;
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
    // channel order of the reference: flow x, flow y, feature 0..127; padded to 144
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        float v = 0.0f;
        if (c == 0) v = flow[p];
        else if (c == 1) v = flow[HW + p];
        else if (c < 130) v = __half2float(feat[p * 128 + (c - 2)]);
        oh[j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(out + p * 144 + g * 8) = o;
}

// ---- convex upsampling (gmflow.py:80-90): mask [h][w][144] fp16 (9 x 4 x 4), flow [2][h][w] -> [2][4h][4w] ---
__global__ void __launch_bounds__(kGfThreads)
convex_upsample_kernel(const __half* __restrict__ mask, const float* __restrict__ flow, float* __restrict__ out, int h, int w)
{
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    const size_t HW = (size_t)h * w;
    if (i >= HW * 16) return;
    const int sub = (int)(i % 16);
    const size_t p = i / 16;
    const int y = (int)(p / w), x = (int)(p - (size_t)y * w);
    const int ky = sub / 4, kx = sub % 4;
    float m[9], mx = -3.0e38f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { m[t] = __half2float(mask[p * 144 + t * 16 + sub]); mx = fmaxf(mx, m[t]); }
    float sum = 0.0f, ax = 0.0f, ay = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float e = __expf(m[t] - mx);
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        float fx = 0.0f, fy = 0.0f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) { fx = 4.0f * flow[(size_t)yy * w + xx]; fy = 4.0f * flow[HW + (size_t)yy * w + xx]; }
        sum += e; ax += e * fx; ay += e * fy;
    }
    const size_t W4 = (size_t)w * 4, o = ((size_t)y * 4 + ky) * W4 + (size_t)x * 4 + kx;
    out[o] = ax / sum;
    out[HW * 16 + o] = ay / sum;
}

// out = alpha * a + beta * b (fp32; b may be NULL): flow = flow + flow_pred, flow * 2 after up-sampling
__global__ void __launch_bounds__(kGfThreads)
axpby_f32_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta, float* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * kGfThreads + threadIdx.x;
    if (i < n) out[i] = b ? alpha * a[i] + beta * b[i] : alpha * a[i];
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_gmflow_normalize_img(const float* in, float* out, int H, int W, void* stream)
{
    if (!in || !out || H <= 0 || W <= 0) return DRBA_E_ARG;
    const size_t HW = (size_t)H * W;
    normalize_img_kernel<<<cdiv(HW, kGfThreads), kGfThreads, 0, as_stream(stream)>>>(in, out, HW);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_inorm_stats(const void* x, int C, int H, int W, double* stats_zeroed, void* stream)
{
    if (!x || !stats_zeroed || C <= 0 || C > 128 || C % 8 != 0 || H <= 0 || W <= 0) return DRBA_E_ARG;
    if (!aligned16(x)) return DRBA_E_ALIGN;
    const size_t HW = (size_t)H * W;
    int ppb = 512;
    while (ppb > 64 && HW / ppb < 2 * kNumSMs) ppb >>= 1;     // enough blocks to fill the machine
    inorm_stats_kernel<<<cdiv(HW, ppb), kGfThreads, 0, as_stream(stream)>>>((const __half*)x, C, HW, ppb, stats_zeroed);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_inorm_apply(const void* x, const double* stats, int relu_x, const void* skip, const double* skip_stats,
                            int final_relu, void* out, int C, int H, int W, void* stream)
{
    if (!x || !stats || !out || C <= 0 || C % 8 != 0 || H <= 0 || W <= 0) return DRBA_E_ARG;
    if (!aligned16(x) || !aligned16(out) || (skip && !aligned16(skip))) return DRBA_E_ALIGN;
    const size_t n = (size_t)H * W * (C / 8);
    inorm_apply_kernel<<<cdiv(n, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)x, stats, relu_x, (const __half*)skip,
                                                                                skip_stats, final_relu, (__half*)out, C, (size_t)H * W);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_add_position(void* x, const float* pos, int B, int h, int w, int wh, int ww, int C, void* stream)
{
    if (!x || !pos || B <= 0 || h <= 0 || w <= 0 || wh <= 0 || ww <= 0 || C <= 0) return DRBA_E_ARG;
    const size_t n = (size_t)B * h * w * C;
    add_position_kernel<<<cdiv(n, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((__half*)x, pos, B, h, w, wh, ww, C);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_window_pack(const void* src, void* dst, int B, int h, int w, int C, int k, int shifted, int rows_pad, int transposed, void* stream)
{
    if (!src || !dst || B <= 0 || h <= 0 || w <= 0 || C <= 0 || C % 8 != 0 || k <= 0 || h % k != 0 || w % k != 0) return DRBA_E_ARG;
    if (rows_pad < (h / k) * (w / k)) return DRBA_E_ARG;
    if (!aligned16(src) || !aligned16(dst)) return DRBA_E_ALIGN;
    const int sh = shifted ? (h / k) / 2 : 0, sw = shifted ? (w / k) / 2 : 0;
    const size_t n = (size_t)B * h * w * (C / 8);
    if (transposed && C % 64 == 0 && C <= 256 && rows_pad % 8 == 0 && (size_t)B * k * k <= 65535) {
        const int Lw = (h / k) * (w / k);
        const dim3 grid((Lw + 63) / 64, B * k * k);
        window_pack_t_kernel<<<grid, kGfThreads, (size_t)64 * C * 2, as_stream(stream)>>>((const __half*)src, (__half*)dst, h, w, C, k, sh, sw, rows_pad);
        DRBA_RETURN_IF_LAUNCH_FAILED();
        return DRBA_OK;
    }
    window_pack_kernel<<<cdiv(n, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)src, (__half*)dst, B, h, w, C, k, sh, sw, rows_pad, transposed);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_softmax_rows(void* S, int nwin_total, int Lw, int ld, int shifted, int k, int h, int w, void* stream)
{
    if (!S || nwin_total <= 0 || Lw <= 0 || ld < Lw || k <= 0 || h % k != 0 || w % k != 0 || (h / k) * (w / k) != Lw) return DRBA_E_ARG;
    if (ld > kSoftmaxGenericCols) return DRBA_E_UNSUPPORTED;
    const dim3 grid(cdiv((size_t)Lw, kGfThreads / 32), nwin_total);
    if (ld > kSoftmaxMaxCols) {
        softmax_rows_generic_kernel<<<grid, kGfThreads, 0, as_stream(stream)>>>((__half*)S, nwin_total, Lw, ld, shifted, k, h / k, w / k, h, w);
        DRBA_RETURN_IF_LAUNCH_FAILED();
        return DRBA_OK;
    }
    if (ld % 8 != 0 || !aligned16(S)) return DRBA_E_ALIGN;
    if (ld <= 512) softmax_rows_kernel<2><<<grid, kGfThreads, 0, as_stream(stream)>>>((__half*)S, nwin_total, Lw, ld, shifted, k, h / k, w / k, h, w);
    else softmax_rows_kernel<8><<<grid, kGfThreads, 0, as_stream(stream)>>>((__half*)S, nwin_total, Lw, ld, shifted, k, h / k, w / k, h, w);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_ln_residual(const void* src, const void* m, const float* gamma, const float* beta, void* out,
                            int B, int h, int w, int C, int k, int shifted, int rows_pad, int cat, void* stream)
{
    if (!src || !m || !gamma || !beta || !out || C != 128 || B <= 0 || h <= 0 || w <= 0) return DRBA_E_ARG;
    if (k > 0 && (h % k != 0 || w % k != 0 || rows_pad < (h / k) * (w / k))) return DRBA_E_ARG;
    const int sh = (k > 0 && shifted) ? (h / k) / 2 : 0, sw = (k > 0 && shifted) ? (w / k) / 2 : 0;
    const size_t threads = (size_t)B * h * w * 32;
    ln_residual_kernel<<<cdiv(threads, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)src, (const __half*)m, gamma, beta, (__half*)out,
                                                                                      B, h, w, k, sh, sw, rows_pad, cat);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_soft_readout(const void* S, int rows, int cols, int ld, const float* val, int w, int subtract_grid, float scale, float* out, void* stream)
{
    if (!S || !out || rows <= 0 || cols <= 0 || ld < cols || w <= 0) return DRBA_E_ARG;
    soft_readout_kernel<<<cdiv((size_t)rows * 32, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)S, rows, cols, ld, val, w, subtract_grid, scale, out);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_local_match(const void* f0, const void* f1, int h, int w, int C, int radius, float* flow, void* stream)
{
    if (!f0 || !f1 || !flow || C != 128 || h <= 0 || w <= 0 || radius <= 0) return DRBA_E_ARG;
    local_match_kernel<<<cdiv((size_t)h * w * 32, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)f0, (const __half*)f1, h, w, radius, flow);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_local_propagate(const void* q, const void* kmap, const float* flow, int h, int w, int C, float* out, void* stream)
{
    if (!q || !kmap || !flow || !out || C != 128 || h <= 0 || w <= 0) return DRBA_E_ARG;
    local_propagate_kernel<<<cdiv((size_t)h * w * 32, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)q, (const __half*)kmap, flow, h, w, out);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_warp_feature(const void* f, const float* flow, void* out, int h, int w, int C, void* stream)
{
    if (!f || !flow || !out || C != 128 || h <= 0 || w <= 0) return DRBA_E_ARG;
    warp_feature_kernel<<<cdiv((size_t)h * w * (C / 8), kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)f, flow, (__half*)out, h, w);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_upsampler_input(const float* flow, const void* feat, void* out144, int h, int w, void* stream)
{
    if (!flow || !feat || !out144 || h <= 0 || w <= 0) return DRBA_E_ARG;
    upsampler_input_kernel<<<cdiv((size_t)h * w * 18, kGfThreads), kGfThreads, 0, as_stream(stream)>>>(flow, (const __half*)feat, (__half*)out144, (size_t)h * w);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_gmflow_convex_upsample(const void* mask144, const float* flow, float* out, int h, int w, void* stream)
{
    if (!mask144 || !flow || !out || h <= 0 || w <= 0) return DRBA_E_ARG;
    convex_upsample_kernel<<<cdiv((size_t)h * w * 16, kGfThreads), kGfThreads, 0, as_stream(stream)>>>((const __half*)mask144, flow, out, h, w);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

int drba_axpby_f32(const float* a, float alpha, const float* b, float beta, float* out, size_t n, void* stream)
{
    if (!a || !out) return DRBA_E_ARG;
    if (n == 0) return DRBA_OK;
    axpby_f32_kernel<<<cdiv(n, kGfThreads), kGfThreads, 0, as_stream(stream)>>>(a, alpha, b, beta, out, n);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
