/*
 * drba_oracle.c -- CPU restatement of the DRBA hot-path arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under drba_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / CPU baseline.
 *
 * Parity status: PINNED.  Every function below is checked bit-exactly (fp32)
 * or to <=1e-6 against outputs of the reference itself, run in the build
 * container by tests/golden/make_golden.py (fixtures committed under
 * tests/golden/), see tests/test_oracle_golden.py.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the DRBA checkout, commit a99ce27).  Plain scalar C, fp32, compiled with
 * -ffp-contract=off so no FMA contraction changes the rounding.
 *
 * Layout everywhere: NCHW contiguous float32, flow channel 0 = x, 1 = y.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* mode / eps enums shared with include/drba_b200.h (same numeric values) */
enum { ORC_SUM = 0, ORC_AVG = 1, ORC_LINEAR = 2, ORC_SOFT = 3 };
enum { ORC_ADDEPS = 0, ORC_ZEROEPS = 1, ORC_CLIPEPS = 2, ORC_NOEPS = 3 /* unknown suffix: no branch of :273-290 fires */ };

/* ------------------------------------------------------------------------ */
/* Summation splat.                                                          */
/* models/softsplat/softsplat.py:306-357 (CuPy kernel `softsplat_out`) and   */
/* its CPU twin models/softsplat/softsplat_torch.py:110-176.                 */
/* Accumulation order follows the torch twin (corner-major: all north-west   */
/* contributions in pixel order, then NE, SW, SE -- softsplat_torch.py:146-  */
/* 174), which makes fp32 results bit-identical to the reference CPU path.   */
/* Non-finite targets are skipped (softsplat.py:323-324), out-of-bounds      */
/* corners are skipped (softsplat.py:342-356).                               */
/* ------------------------------------------------------------------------ */
ORC_API void orc_splat_sum(const float* in, const float* flow, float* out,
                           int N, int C, int H, int W)
{
    const size_t HW = (size_t)H * W;
    memset(out, 0, sizeof(float) * (size_t)N * C * HW);
    for (int corner = 0; corner < 4; ++corner) {
        const int dx = corner & 1, dy = corner >> 1;
        for (int n = 0; n < N; ++n) {
            const float* fx = flow + (size_t)n * 2 * HW;
            const float* fy = fx + HW;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const size_t p = (size_t)y * W + x;
                    const float X = (float)x + fx[p];
                    const float Y = (float)y + fy[p];
                    if (!isfinite(X) || !isfinite(Y)) continue;
                    const float flX = floorf(X), flY = floorf(Y);
                    /* the reference converts the floor to int32 */
                    if (flX < -2147483000.0f || flX > 2147483000.0f ||
                        flY < -2147483000.0f || flY > 2147483000.0f) continue;
                    const int x0 = (int)flX, y0 = (int)flY;
                    const int tx = x0 + dx, ty = y0 + dy;
                    if (tx < 0 || tx >= W || ty < 0 || ty >= H) continue;
                    /* NW=(x0+1-X)(y0+1-Y) NE=(X-x0)(y0+1-Y) SW=(x0+1-X)(Y-y0) SE=(X-x0)(Y-y0) */
                    const float wx = dx ? (X - (float)x0) : ((float)(x0 + 1) - X);
                    const float wy = dy ? (Y - (float)y0) : ((float)(y0 + 1) - Y);
                    const float w = wx * wy;
                    const size_t q = (size_t)ty * W + tx;
                    for (int c = 0; c < C; ++c) {
                        const size_t base = ((size_t)n * C + c) * HW;
                        out[base + q] += in[base + p] * w;
                    }
                }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* softsplat(tenIn, tenFlow, tenMetric, strMode)                             */
/* models/softsplat/softsplat.py:248-293 (mode prep :260-267, normalise      */
/* :273-290).  `metric` may be NULL for sum / avg.  `work` must hold         */
/* N*(C+1)*H*W*2 floats (augmented input + augmented output); if NULL it is  */
/* malloc'ed here.                                                           */
/* ------------------------------------------------------------------------ */
ORC_API int orc_softsplat(const float* in, const float* flow, const float* metric,
                          float* out, int N, int C, int H, int W,
                          int mode, int eps_mode, float* work)
{
    const size_t HW = (size_t)H * W;
    if (mode == ORC_SUM) {
        orc_splat_sum(in, flow, out, N, C, H, W);
        return 0;
    }
    if ((mode == ORC_LINEAR || mode == ORC_SOFT) && metric == NULL) return -1;
    const int C1 = C + 1;
    float* own = NULL;
    if (work == NULL) {
        own = (float*)malloc(sizeof(float) * 2 * (size_t)N * C1 * HW);
        if (!own) return -2;
        work = own;
    }
    float* aug_in = work;
    float* aug_out = work + (size_t)N * C1 * HW;
    for (int n = 0; n < N; ++n) {
        const float* m = metric ? metric + (size_t)n * HW : NULL;
        float* wch = aug_in + ((size_t)n * C1 + C) * HW;
        for (size_t p = 0; p < HW; ++p) {
            float wv;
            if (mode == ORC_AVG) wv = 1.0f;                 /* :260-261 */
            else if (mode == ORC_LINEAR) wv = m[p];         /* :263-264 */
            else wv = expf(m[p]);                           /* :266-267 */
            wch[p] = wv;
        }
        for (int c = 0; c < C; ++c) {
            const float* src = in + ((size_t)n * C + c) * HW;
            float* dst = aug_in + ((size_t)n * C1 + c) * HW;
            if (mode == ORC_AVG) memcpy(dst, src, sizeof(float) * HW);
            else for (size_t p = 0; p < HW; ++p) dst[p] = src[p] * wch[p];
        }
    }
    orc_splat_sum(aug_in, flow, aug_out, N, C1, H, W);
    for (int n = 0; n < N; ++n) {
        const float* den = aug_out + ((size_t)n * C1 + C) * HW;
        for (int c = 0; c < C; ++c) {
            const float* num = aug_out + ((size_t)n * C1 + c) * HW;
            float* dst = out + ((size_t)n * C + c) * HW;
            for (size_t p = 0; p < HW; ++p) {
                float d = den[p];
                if (eps_mode == ORC_ADDEPS) d = d + 0.0000001f;          /* :277,:280 */
                else if (eps_mode == ORC_ZEROEPS) d = (d == 0.0f) ? 1.0f : d; /* :283 */
                else if (eps_mode == ORC_CLIPEPS) d = d < 0.0000001f ? 0.0000001f : d; /* :286 */
                dst[p] = num[p] / d;
            }
        }
    }
    free(own);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* distance_calculator: models/utils/tools.py:77-80                          */
/* ------------------------------------------------------------------------ */
static inline float orc_dist(float u, float v) { return sqrtf(u * u + v * v); }

/* ------------------------------------------------------------------------ */
/* get_drm_t: models/drm.py:10-62.  Scalar bisection on t mirrored per pixel */
/* with the per-pixel step factor b_drm = drm.  Note both `if`s can fire in  */
/* one loop iteration (drm.py:44 and :53 are not elif).  The scalar state    */
/* is Python double arithmetic.                                              */
/* ------------------------------------------------------------------------ */
ORC_API void orc_get_drm_t(const float* drm, double t, double precision, float* out, size_t n)
{
    /* record the branch sequence first (it depends on scalar t only) */
    unsigned char seq[256];
    int len = 0;
    double x = 0.5, b = 0.5, l = 0.0, r = 1.0;
    while (fabs(x - t) > precision && len < 254) {
        if (x > t) { r = x; x = x - (x - l) * b; seq[len++] = 0; }
        if (x < t) { l = x; x = x + (r - x) * b; seq[len++] = 1; }
    }
    for (size_t i = 0; i < n; ++i) {
        const float bd = drm[i];
        float xd = drm[i], ld = drm[i] * 0.0f, rd = drm[i] * 0.0f + 1.0f;
        for (int k = 0; k < len; ++k) {
            if (seq[k] == 0) { rd = xd; xd = xd - (xd - ld) * bd; }
            else             { ld = xd; xd = xd + (rd - xd) * bd; }
        }
        out[i] = xd;
    }
}

/* hole fill shared by every DRM variant: drm.py:95-102 (and :138-148, :180-189)
 * mask = den / (den + 1e-7) is what an 'avg'/'soft' splat of a ones map returns. */
static void orc_fill(float* warped, const float* ones_warped, const float* unaligned, size_t n)
{
    for (size_t i = 0; i < n; ++i)
        if (ones_warped[i] < 0.999f) warped[i] = unaligned[i];
}

/* ------------------------------------------------------------------------ */
/* calc_drm_rife (drm.py:65-107) and calc_drm_rife_auxiliary (drm.py:158-195)*/
/* metric10/metric12 NULL  -> 'avg' (calc_drm_rife, or aux without metrics)  */
/* Outputs: drm_t1_t01, drm_t1_t12, each [N,1,H,W].                          */
/* ------------------------------------------------------------------------ */
ORC_API int orc_drm_rife(double t, const float* flow10, const float* flow12,
                         const float* metric10, const float* metric12, int linear,
                         float* out_t01, float* out_t12, int N, int H, int W)
{
    const size_t HW = (size_t)H * W, n1 = (size_t)N * HW;
    float* buf = (float*)malloc(sizeof(float) * (n1 * 6 + n1 * 2 * 2));
    if (!buf) return -2;
    float* drm10 = buf, *drm12 = buf + n1, *u0 = buf + 2 * n1, *u1 = buf + 3 * n1;
    float* ones = buf + 4 * n1, *mask = buf + 5 * n1;
    float* sf10 = buf + 6 * n1, *sf12 = buf + 8 * n1;
    const int mode = (metric10 && metric12) ? ORC_SOFT : ORC_AVG;
    for (int n = 0; n < N; ++n)
        for (size_t p = 0; p < HW; ++p) {
            const size_t i = (size_t)n * HW + p;
            const float d10 = orc_dist(flow10[(size_t)n * 2 * HW + p], flow10[(size_t)n * 2 * HW + HW + p]) + 1e-4f;
            const float d12 = orc_dist(flow12[(size_t)n * 2 * HW + p], flow12[(size_t)n * 2 * HW + HW + p]) + 1e-4f;
            drm10[i] = d10 / (d10 + d12);
            drm12[i] = d12 / (d10 + d12);
            ones[i] = drm10[i] * 0.0f + 1.0f;
        }
    if (linear) {
        const float tf = (float)t;
        for (size_t i = 0; i < n1; ++i) { u0[i] = drm10[i] * tf * 2.0f; u1[i] = drm12[i] * tf * 2.0f; }
    } else {
        orc_get_drm_t(drm10, t, 1e-3, u0, n1);
        orc_get_drm_t(drm12, t, 1e-3, u1, n1);
    }
    for (int n = 0; n < N; ++n)
        for (int c = 0; c < 2; ++c)
            for (size_t p = 0; p < HW; ++p) {
                const size_t f = ((size_t)n * 2 + c) * HW + p, i = (size_t)n * HW + p;
                sf10[f] = flow10[f] * u1[i];   /* drm.py:89  flow10 * drm_t1_unaligned */
                sf12[f] = flow12[f] * u0[i];   /* drm.py:90  flow12 * drm_t0_unaligned */
            }
    int rc = 0;
    rc |= orc_softsplat(u1, sf10, metric10, out_t01, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    rc |= orc_softsplat(ones, sf10, metric10, mask, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    orc_fill(out_t01, mask, u1, n1);
    rc |= orc_softsplat(u0, sf12, metric12, out_t12, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    rc |= orc_softsplat(ones, sf12, metric12, mask, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    orc_fill(out_t12, mask, u0, n1);
    free(buf);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* calc_drm_gmfss: drm.py:110-155.  No +1e-4 on the distances: 0/0 = NaN is  */
/* kept.  Outputs in dict order: drm0t_t01, drm1t_t01, drm1t_t12, drm2t_t12. */
/* ------------------------------------------------------------------------ */
ORC_API int orc_drm_gmfss(double t, const float* flow10, const float* flow12,
                          const float* metric10, const float* metric12, int linear,
                          float* drm0t_t01, float* drm1t_t01, float* drm1t_t12, float* drm2t_t12,
                          int N, int H, int W)
{
    const size_t HW = (size_t)H * W, n1 = (size_t)N * HW;
    float* buf = (float*)malloc(sizeof(float) * n1 * 6);
    if (!buf) return -2;
    float* drm10 = buf, *drm12 = buf + n1, *un0 = buf + 2 * n1, *un2 = buf + 3 * n1;
    float* ones = buf + 4 * n1, *mask = buf + 5 * n1;
    const int mode = (metric10 && metric12) ? ORC_SOFT : ORC_AVG;
    for (int n = 0; n < N; ++n)
        for (size_t p = 0; p < HW; ++p) {
            const size_t i = (size_t)n * HW + p;
            const float d10 = orc_dist(flow10[(size_t)n * 2 * HW + p], flow10[(size_t)n * 2 * HW + HW + p]);
            const float d12 = orc_dist(flow12[(size_t)n * 2 * HW + p], flow12[(size_t)n * 2 * HW + HW + p]);
            drm10[i] = d10 / (d10 + d12);
            drm12[i] = d12 / (d10 + d12);
        }
    if (linear) {
        const float tf = (float)t;
        for (size_t i = 0; i < n1; ++i) { drm1t_t01[i] = drm12[i] * tf * 2.0f; drm1t_t12[i] = drm10[i] * tf * 2.0f; }
    } else {
        orc_get_drm_t(drm12, t, 1e-3, drm1t_t01, n1);
        orc_get_drm_t(drm10, t, 1e-3, drm1t_t12, n1);
    }
    for (size_t i = 0; i < n1; ++i) { un0[i] = 1.0f - drm1t_t01[i]; un2[i] = 1.0f - drm1t_t12[i]; }
    int rc = 0;
    rc |= orc_softsplat(un0, flow10, metric10, drm0t_t01, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    rc |= orc_softsplat(un2, flow12, metric12, drm2t_t12, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    /* ones_mask = drm0t_t01.clone()*0 + 1 (drm.py:136): NaN in the warped map stays NaN */
    for (size_t i = 0; i < n1; ++i) ones[i] = drm0t_t01[i] * 0.0f + 1.0f;
    rc |= orc_softsplat(ones, flow10, metric10, mask, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    orc_fill(drm0t_t01, mask, un0, n1);
    rc |= orc_softsplat(ones, flow12, metric12, mask, N, 1, H, W, mode, ORC_ADDEPS, NULL);
    orc_fill(drm2t_t12, mask, un2, n1);
    free(buf);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Backward warp: models/rife_426_heavy/warplayer.py:8-22 (border padding,   */
/* align_corners=True) and the zeros-padding twin used by                    */
/* models/model_gmfss/MetricNet.py:10-20 / models/gmflow/geometry.py:60-67.  */
/* Restated in pixel coordinates: the normalised grid of the reference       */
/* (linspace(-1,1,W) + flow/((W-1)/2)) un-normalises under align_corners=True*/
/* to x + flow_x exactly up to fp32 rounding (validated <= 2e-6).            */
/* pad_mode: 0 = border (clamp coordinate), 1 = zeros.                       */
/* ------------------------------------------------------------------------ */
ORC_API void orc_backwarp(const float* in, const float* flow, float* out,
                          int N, int C, int H, int W, int pad_mode)
{
    const size_t HW = (size_t)H * W;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x;
                float sx = (float)x + flow[(size_t)n * 2 * HW + p];
                float sy = (float)y + flow[(size_t)n * 2 * HW + HW + p];
                if (pad_mode == 0) {
                    sx = fminf(fmaxf(sx, 0.0f), (float)(W - 1));
                    sy = fminf(fmaxf(sy, 0.0f), (float)(H - 1));
                }
                const float fx0 = floorf(sx), fy0 = floorf(sy);
                const int x0 = (int)fx0, y0 = (int)fy0;
                const float ax = sx - fx0, ay = sy - fy0;
                const float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay);
                const float w10 = (1.0f - ax) * ay, w11 = ax * ay;
                for (int c = 0; c < C; ++c) {
                    const float* src = in + ((size_t)n * C + c) * HW;
                    float acc = 0.0f;
                    #define TAP(yy, xx, ww) \
                        if ((xx) >= 0 && (xx) < W && (yy) >= 0 && (yy) < H) acc += src[(size_t)(yy) * W + (xx)] * (ww);
                    TAP(y0, x0, w00) TAP(y0, x0 + 1, w01) TAP(y0 + 1, x0, w10) TAP(y0 + 1, x0 + 1, w11)
                    #undef TAP
                    out[((size_t)n * C + c) * HW + p] = acc;
                }
            }
}

/* ------------------------------------------------------------------------ */
/* Bilinear resize, F.interpolate(mode='bilinear') as used by                */
/* IFNet_HDv3.py:85-92 (align_corners=False, scale_factor given) and         */
/* models/utils/tools.py:72 (size given).  rh/rw are the source-per-dest     */
/* coordinate ratios: 1/scale_factor when a scale_factor is passed, in/out   */
/* when a size is passed.  align_corners=False: src=(dst+0.5)*r-0.5 clamped  */
/* at 0; align_corners=True: src=dst*(in-1)/(out-1).                         */
/* ------------------------------------------------------------------------ */
ORC_API void orc_resize_bilinear(const float* in, float* out, int N, int C,
                                 int H, int W, int OH, int OW,
                                 int align_corners, float rh, float rw)
{
    if (align_corners) {
        rh = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.0f;
        rw = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.0f;
    }
    for (int nc = 0; nc < N * C; ++nc) {
        const float* src = in + (size_t)nc * H * W;
        float* dst = out + (size_t)nc * OH * OW;
        for (int oy = 0; oy < OH; ++oy) {
            float sy = align_corners ? rh * oy : rh * ((float)oy + 0.5f) - 0.5f;
            if (!align_corners && sy < 0.0f) sy = 0.0f;
            const int y0 = (int)sy;
            const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
            const float ly = sy - (float)y0, hy = 1.0f - ly;
            for (int ox = 0; ox < OW; ++ox) {
                float sx = align_corners ? rw * ox : rw * ((float)ox + 0.5f) - 0.5f;
                if (!align_corners && sx < 0.0f) sx = 0.0f;
                const int x0 = (int)sx;
                const int x1 = x0 + (x0 < W - 1 ? 1 : 0);
                const float lx = sx - (float)x0, hx = 1.0f - lx;
                dst[(size_t)oy * OW + ox] =
                    hy * (hx * src[(size_t)y0 * W + x0] + lx * src[(size_t)y0 * W + x1]) +
                    ly * (hx * src[(size_t)y1 * W + x0] + lx * src[(size_t)y1 * W + x1]);
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* RIFE.calc_flow's flow inversion: models/rife.py:59-73.                    */
/* flow_t0 [N,2,H,W] (flow from the mid frame towards a source frame) ->     */
/* out = 2 * fill( -splat_avg(flow_t0, flow_t0), holes <- max(H,W) ).        */
/* ------------------------------------------------------------------------ */
ORC_API int orc_rife_invert_flow(const float* flow_t0, float* out, int N, int H, int W)
{
    const size_t HW = (size_t)H * W;
    float* ones = (float*)malloc(sizeof(float) * (size_t)N * HW * 4);
    if (!ones) return -2;
    float* mask = ones + (size_t)N * HW * 2;
    int rc = orc_softsplat(flow_t0, flow_t0, NULL, out, N, 2, H, W, ORC_AVG, ORC_ADDEPS, NULL);
    for (size_t i = 0; i < (size_t)N * 2 * HW; ++i) { out[i] = -1.0f * out[i]; ones[i] = out[i] * 0.0f + 1.0f; }
    rc |= orc_softsplat(ones, flow_t0, NULL, mask, N, 2, H, W, ORC_AVG, ORC_ADDEPS, NULL);
    const float big = (float)(H > W ? H : W);
    for (size_t i = 0; i < (size_t)N * 2 * HW; ++i) {
        if (mask[i] < 0.999f) out[i] = ones[i] * big;
        out[i] = out[i] * 2.0f;
    }
    free(ones);
    return rc;
}

ORC_API int orc_version(void) { return 1; }
