/*
 * drba_b200.h -- C ABI of libdrba_b200.so (sm_100a).
 *
 * This is the drop-in boundary for DRBA's per-triplet interpolation hot path.
 * The reference's only FFI precedent is the CuPy launch in
 * models/softsplat/softsplat.py:362-367 (raw data_ptr()s + dims, caller-allocated
 * output, torch's current stream); every entry point below keeps that contract:
 *
 *   - plain pointers and sizes, no torch / C++ types;
 *   - all pointers are DEVICE pointers to contiguous tensors (NCHW unless noted);
 *   - the call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never allocates, never synchronises, is re-entrant;
 *   - returns 0, a negative DRBA_E_* argument error, or a positive cudaError_t.
 *
 * Workspaces: ops that scatter need a zero-filled accumulator.  The caller owns it.
 * Contract: a workspace handed to any drba_* op must be all-zero on entry (use
 * drba_workspace_clear once after allocation); every op leaves it all-zero on exit
 * (the resolve pass re-zeroes what it reads), so it can be reused without memsets.
 *
 * Reference citations are relative to the DRBA checkout (commit a99ce27).
 */
#ifndef DRBA_B200_H
#define DRBA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DRBA_API __attribute__((visibility("default")))
#else
#define DRBA_API
#endif

/* error codes (negative; positive return values are cudaError_t) */
#define DRBA_OK 0
#define DRBA_E_ARG (-1)         /* NULL / non-positive dimension / bad enum        */
#define DRBA_E_WORKSPACE (-2)   /* workspace missing or too small                  */
#define DRBA_E_UNSUPPORTED (-3) /* valid request this build cannot serve           */
#define DRBA_E_ALIGN (-4)       /* pointer not aligned as required (16 B)          */
#define DRBA_E_BARRIER (-5)     /* a persistent conv program timed out at its grid barrier: results invalid (sticky) */

/* softsplat modes / eps variants: models/softsplat/softsplat.py:260-267, :273-290 */
enum { DRBA_SPLAT_SUM = 0, DRBA_SPLAT_AVG = 1, DRBA_SPLAT_LINEAR = 2, DRBA_SPLAT_SOFT = 3 };
enum { DRBA_EPS_ADD = 0, DRBA_EPS_ZERO = 1, DRBA_EPS_CLIP = 2,
       DRBA_EPS_NONE = 3 /* unknown "-suffix": the reference divides by the raw denominator (:273-290 fall through) */ };
/* padding of the backward warp: warplayer.py:22 (border) / MetricNet.py:20 (zeros) */
enum { DRBA_PAD_BORDER = 0, DRBA_PAD_ZEROS = 1 };
/* activation dtypes of the conv engine */
enum { DRBA_F32 = 0, DRBA_F16 = 1 };

DRBA_API int drba_version(void);
DRBA_API const char* drba_error_string(int code);
/* zero-fill a workspace (once, after allocation) */
DRBA_API int drba_workspace_clear(void* ws, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------
 * softsplat(tenIn, tenFlow, tenMetric, strMode)
 * replaces models/softsplat/softsplat.py:248-293 + kernel :306-357
 * (twin: softsplat_torch.py:20-179).  fp32 NCHW: in [N,C,H,W], flow [N,2,H,W]
 * (ch0 = x, ch1 = y, pixels), metric [N,1,H,W] or NULL (sum/avg), out [N,C,H,W].
 * The accumulator lives in `ws` (channel-interleaved groups of 4, L2 resident);
 * larger C is processed in channel chunks sized to the workspace given.
 * ------------------------------------------------------------------------- */
DRBA_API size_t drba_softsplat_workspace_bytes(int N, int C, int H, int W, int mode);
DRBA_API int drba_softsplat_f32(const float* in, const float* flow, const float* metric, float* out,
                                int N, int C, int H, int W, int mode, int eps_mode,
                                void* ws, size_t ws_bytes, void* stream);
/* variant selector for measurements: 0 = aggregated vector reds (default),
 * 1 = one scalar atomic per corner and channel (the reference kernel's scheme) */
DRBA_API int drba_softsplat_f32_variant(const float* in, const float* flow, const float* metric, float* out,
                                        int N, int C, int H, int W, int mode, int eps_mode,
                                        void* ws, size_t ws_bytes, int variant, void* stream);

/* ---------------------------------------------------------------------------
 * Splat with reusable lists (csrc/splat_gather.cu): GMFSS forward-warps several tensors with the same
 * (flow, metric) -- the half-resolution image and a 64/128/192-channel feature map,
 * models/model_gmfss/GMFSS.py:96-115.  build: count / scan / fill / sort+weights, once per (flow, metric),
 * batch 1; apply: a pure gather, deterministic; release: re-zeroes the workspace (library contract).
 *   apply_nhwc_f16: in [H][W][in_cstride] fp16 (C % 8 == 0 channels used), out [H][W][out_cstride] fp16 written
 *                   at channel offset out_coffset (straight into a concat buffer), optional scalar PReLU on write
 *   apply_nchw_f32: in / out [C][H][W] fp32
 * normalise != 0: divide by the splatted weight (avg / linear / soft, softsplat.py:273-290).
 * ------------------------------------------------------------------------- */
DRBA_API size_t drba_splat_lists_workspace_bytes(int H, int W);
DRBA_API int drba_splat_lists_build(const float* flow, const float* metric, int mode, int H, int W,
                                    void* ws, size_t ws_bytes, void* stream);
DRBA_API int drba_splat_lists_apply_nhwc_f16(const void* ws, const void* in, int C, int in_cstride,
                                             void* out, int out_cstride, int out_coffset, int H, int W,
                                             int normalise, int eps_mode, int use_prelu, float slope, void* stream);
DRBA_API int drba_splat_lists_apply_nchw_f32(const void* ws, const float* in, float* out, int C, int H, int W,
                                             int normalise, int eps_mode, void* stream);
DRBA_API int drba_splat_lists_release(void* ws, int H, int W, void* stream);

/* ---------------------------------------------------------------------------
 * RIFE.calc_flow's flow inversion, models/rife.py:59-73:
 *   out = 2 * fill(-splat_avg(flow_t0, flow_t0), holes <- max(H, W))
 * flow_t0, out: [N,2,H,W] fp32.  ws >= N*H*W*16 bytes.
 * ------------------------------------------------------------------------- */
DRBA_API size_t drba_rife_invert_flow_workspace_bytes(int N, int H, int W);
DRBA_API int drba_rife_invert_flow_f32(const float* flow_t0, float* out, int N, int H, int W,
                                       void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * get_drm_t(drm, t, precision): models/drm.py:10-62.  n elements, fp32.
 * ------------------------------------------------------------------------- */
DRBA_API int drba_get_drm_t_f32(const float* drm, double t, double precision, float* out, size_t n, void* stream);

/* ---------------------------------------------------------------------------
 * calc_drm_rife (models/drm.py:65-107) and calc_drm_rife_auxiliary (:158-195)
 * fused with distance_calculator (models/utils/tools.py:77-80).
 * flow10, flow12 [N,2,H,W]; metric10/metric12 [N,1,H,W] or both NULL ('avg');
 * out_t01 ('drm_t1_t01') / out_t12 ('drm_t1_t12') [N,1,H,W]; either may be NULL
 * when the caller needs only one map.  ws >= N*H*W*16 bytes.
 * ------------------------------------------------------------------------- */
DRBA_API size_t drba_drm_workspace_bytes(int N, int H, int W);
DRBA_API int drba_drm_rife_f32(double t, const float* flow10, const float* flow12,
                               const float* metric10, const float* metric12, int linear,
                               float* out_t01, float* out_t12, int N, int H, int W,
                               void* ws, size_t ws_bytes, void* stream);
/* calc_drm_gmfss (models/drm.py:110-155).  Outputs in the reference's dict order;
 * any may be NULL.  No epsilon on the distances: 0/0 = NaN propagates as in the
 * reference. */
DRBA_API int drba_drm_gmfss_f32(double t, const float* flow10, const float* flow12,
                                const float* metric10, const float* metric12, int linear,
                                float* drm0t_t01, float* drm1t_t01, float* drm1t_t12, float* drm2t_t12,
                                int N, int H, int W, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Backward warp (grid_sample bilinear, align_corners=True) in pixel coordinates:
 * models/rife_426_heavy/warplayer.py:8-22 (border), models/model_gmfss/MetricNet.py:10-20
 * and models/gmflow/geometry.py:60-67 (zeros).  in/out [N,C,H,W], flow [N,2,H,W].
 * ------------------------------------------------------------------------- */
DRBA_API int drba_backwarp_f32(const float* in, const float* flow, float* out,
                               int N, int C, int H, int W, int pad_mode, void* stream);

/* ---------------------------------------------------------------------------
 * F.interpolate(mode='bilinear'): IFNet_HDv3.py:85-92, tools.py:72, GMFSS.py:64-77.
 * rh/rw = source-per-destination ratio (1/scale_factor, or in/out when a size is
 * given); ignored when align_corners != 0.
 * ------------------------------------------------------------------------- */
DRBA_API int drba_resize_bilinear_f32(const float* in, float* out, int N, int C, int H, int W,
                                      int OH, int OW, int align_corners, float rh, float rw, void* stream);

/* ---------------------------------------------------------------------------
 * Frame ingest / egress: models/utils/tools.py:33-38 (to_tensor / to_cv2) fused with :59-72
 * (to_inp / to_out / resize).  in_hwc / out_hwc: uint8 [H][W][3] as cv2 decodes / encodes (BGR);
 * float side: NCHW [3][H'][W'] in [0,1].  Bilinear, align_corners=False, ATen's scale = in/out.
 *   ingest: out = interpolate(float(in) / 255, (OH, OW))
 *   egress: out = astype(uint8)(interpolate(in, (OH, OW)) * 255.)   (truncation; wraps like numpy)
 * ------------------------------------------------------------------------- */
DRBA_API int drba_frame_ingest_u8(const unsigned char* in_hwc, float* out_chw, int H, int W, int OH, int OW, void* stream);
DRBA_API int drba_frame_egress_u8(const float* in_chw, unsigned char* out_hwc, int H, int W, int OH, int OW, void* stream);

/* ---------------------------------------------------------------------------
 * Scene-cut detection (csrc/scene.cu): check_scene(x1, x2, threshold) of models/utils/tools.py:27-30 =
 * bilinear 32x32 thumbnails (align_corners=False) + ssim_matlab (models/pytorch_msssim/__init__.py:83-136:
 * 11^3 Gaussian volume window, replicate padding, value range probed from x1) + `ssim < threshold`, one CTA per
 * frame pair, no host synchronisation (the reference has three per pair).
 * Pair i reads the NCHW [3][H][W] fp32 frames x1 + i * pair_stride1 and x2 + i * pair_stride2 (strides in
 * elements; npairs = 1 for two separate frames; a contiguous clip [N][3][H][W] checks all N - 1 neighbour pairs
 * with x2 = x1 + 3HW and both strides 3HW).  ssim_out [npairs] / flag_out [npairs] (1 = scene cut) may be NULL
 * (not both) and may point to mapped pinned host memory.
 * ------------------------------------------------------------------------- */
DRBA_API int drba_check_scene_f32(const float* x1, const float* x2, long long pair_stride1, long long pair_stride2,
                                  int npairs, int H, int W, float threshold, float* ssim_out, int* flag_out, void* stream);

/* ---------------------------------------------------------------------------
 * fp32 direct convolution (exact engine).  Generic form covering every conv on the
 * IFNet path (models/rife_426_heavy/IFNet_HDv3.py:11-25 conv, :28-47 Head, :50-59 ResConv,
 * :80 lastconv ConvTranspose2d(4,2,1) as four phase launches):
 *   out[n,co,oy*OS+PY,ox*OS+PX] = act(bias[co] + res[same] + sum_t sum_ci in[n,ci,oy*S+dy[t],ox*S+dx[t]] * w[t][ci][co])
 * in_strides / out_strides: element strides {n, c, y, x}; `res` (optional) is addressed
 * like `out`.  w is packed [T][Cin][Cout] fp32.  act: 0 none, 1 LeakyReLU(0.2).
 * ------------------------------------------------------------------------- */
DRBA_API int drba_conv2d_direct_f32(const float* in, const float* w, const float* bias, const void* res, void* out,
                                    int out_dtype /* DRBA_F32 | DRBA_F16: dtype of out (and res) */,
                                    int N, int Cin, int H, int W, const long long* in_strides,
                                    int Cout, int OH, int OW, const long long* out_strides,
                                    int S, int OS, int PY, int PX, int T, const int* dy, const int* dx,
                                    int act, void* stream);

/* ---------------------------------------------------------------------------
 * tcgen05 implicit-GEMM convolution (tensor-core engine), fp16 operands, fp32 accumulate.
 * Replaces the cuDNN fp16 convs the reference runs under torch.autocast for
 * IFNet_HDv3.py:62-96 (conv0, ResConv x 8, lastconv).  Batch 1.
 *   in   : NHWC fp16 [H][W][Cin], Cin % 16 == 0 (zero-padded channels)
 *   w    : fp16 [G][T][cout_pad][Cin];  bias: fp32 [G][cout_pad];  dy, dx: int [G][T] tap offsets
 *   out[oy][ox][co] = act(bias + res + sum_t sum_ci in[S*oy + dy[t]][S*ox + dx[t]][ci] * w[g][t][co][ci])
 *   epilogue 0: NHWC fp16, optional residual `res` of the same geometry, act 0 none / 1 LeakyReLU(0.2)
 *               out_os 1: [OH][OW][out_cstride], G = 1
 *               out_os 2: ConvTranspose2d(k4,s2,p1) as G = 4 phase convs (g = py*2+px, T = 4):
 *                         [2*OH][2*OW][out_cstride], phase g lands at (2*oy + py, 2*ox + px)  (Head.cnn3)
 *   epilogue 1: lastconv = ConvTranspose2d(Cin,52,4,2,1)+PixelShuffle(2): G = 4 phases (py*2+px),
 *               T = 4, cout_pad = 64, cout = 52; out = NHWC fp32 [4*OH][4*OW][out_cstride], out_cstride = 16
 *               (13 used) or 8 (channels 0..7 only: flow, mask, feat 0..2).  Equivalent 3x3 form: G = 1, T = 9 (canonical
 *               taps), cout_pad = 256 = 4 phases x 64 with zero weights on the taps a phase does not use (halo mode, one
 *               phase's weights resident per CTA)
 * ------------------------------------------------------------------------- */
DRBA_API int drba_conv_tc_f16(const void* in, int H, int W, int Cin,
                              const void* w, const float* bias, int G, int T, const int* dy, const int* dx,
                              int cout_pad, int cout, int S, int OH, int OW,
                              int epilogue, int act, const void* res, void* out, int out_cstride, int out_os, void* stream);

/* A program of conv layers executed by ONE persistent launch (csrc/conv_tc.cu): layer i+1 starts
 * after a grid-wide barrier, so layer i+1 may read what layer i wrote (e.g. a whole IFBlock:
 * conv0a, conv0b, 8 x ResConv, lastconv -- IFNet_HDv3.py:84-96).  Up to two independent images
 * (in[k] / res[k] / out[k], same geometry, shared weights) are processed side by side.
 * Field meaning as in drba_conv_tc_f16; act: 0 none, 1 LeakyReLU(0.2), 2 PReLU(slope[cout_pad]), 3 ReLU,
 * 4 PReLU(scalar slope0), 5 GELU (erf).
 * sync_ws: 8 zero bytes of device memory (8-byte aligned) owned by the caller, required when
 * nlayers > 1; zero again when the launch completes.  Do not run two programs that share a device
 * concurrently on different streams (each expects to own all SMs between its barriers). */
#define DRBA_CONV_MAX_LAYERS 12
typedef struct drba_conv_layer {
    const void* in[2];
    const void* res[2];
    void* out[2];
    const void* w;
    const float* bias;
    const float* slope;
    int H, W, Cin, G, T;
    int dy[36], dx[36];
    int cout_pad, cout, S, OH, OW, epilogue, act, out_cstride, out_os;
    /* epilogue 0 extensions (GMFSS FeatureNet / MetricNet / GridNet, models/model_gmfss/FusionNet.py:6-33:
     * pre-activation residual blocks whose inputs are consumed raw AND through scalar PReLUs):
     * value = bias + conv + res + res2; out = act(value); out1 = act1(value); out2 = act2(value).
     * act / act1 / act2 = 4: scalar PReLU with slope0 / slope1 / slope2.  NULL pointers are skipped. */
    const void* res2[2];
    void* out1[2];
    void* out2[2];
    int act1, act2;
    float slope0, slope1, slope2;
    /* batched GEMM (GMFlow attention / correlation, models/gmflow/transformer.py:8-17, matching.py:7-43):
     * in = A [H = batch][W = M rows][Cin = K], w = B [batch][cout_pad rows][K] (K-major), G = T = S = 1, no bias:
     * out[b][m][n] = sum_k A[b][m][k] * B[b][n][k].  bias may be NULL for any epilogue-0 layer. */
    int bgemm;
} drba_conv_layer;
DRBA_API int drba_conv_tc_program_f16(const drba_conv_layer* layers, int nlayers, int nimg, void* sync_ws, void* stream);
/* The grid barrier of a program needs every CTA resident: the launcher clamps the grid to what the device can hold
 * (cudaOccupancyMaxActiveBlocksPerMultiprocessor x SM count).  If CTAs still fail to arrive in time (MPS / MIG / a
 * foreign kernel holding SMs) the wait is abandoned so that the GPU does not hang, and a STICKY error word in mapped
 * host memory is set: drba_conv_tc_program_f16 returns DRBA_E_BARRIER from then on, and drba_conv_tc_status() (no
 * synchronisation; a plain host read) reports it for work that was replayed from a CUDA graph. */
DRBA_API int drba_conv_tc_status(void);
/* debug: while a device buffer of 4096 int64 is registered, CTA 0 of every conv launch writes clock64() stamps
 * of its pipeline events into it (NULL switches tracing off; scripts/trace_conv.py decodes). */
DRBA_API int drba_conv_tc_debug_trace(void* dev_buf_4096_i64);

/* ---------------------------------------------------------------------------
 * GMFSS glue (csrc/gmfss.cu).  The conv engine works on NHWC fp16; these are the only places where the NCHW
 * fp32 API tensors are packed into / unpacked from that layout, fused with the arithmetic around the convs.
 *   pack_planes : out[y][x][c] = prelu?(scales[c] * planes[c][y][x]), c < nplanes <= 16, zero padding to 16
 *                 (`planes` / `scales` are HOST arrays of device pointers / floats; scales may be NULL)
 *   unpack      : NHWC fp16 -> NCHW fp32 planes, optional clamp (GMFSS.py:190)
 *   metric_prep : MetricNet.forward up to its first conv (MetricNet.py:45-60; fb check geometry.py:87-108)
 *   scale_flow  : F = t * flow, Z = t * metric resampled to 1/s (s = 1, 2, 4), flow additionally * 1/s
 *                 (GMFSS.py:88-113); t is a per-pixel map [H][W] or, when tmap == NULL, the scalar
 * ------------------------------------------------------------------------- */
DRBA_API int drba_pack_planes_nhwc_f16(const float* const* planes, const float* scales, int nplanes, int H, int W,
                                       int use_prelu, float slope, void* out, int out_cstride, void* stream);
DRBA_API int drba_unpack_nhwc_f16(const void* in, int in_cstride, float* out, int C, int H, int W,
                                  int do_clamp, float lo, float hi, void* stream);
DRBA_API int drba_gmfss_metric_prep(const float* img0, const float* img1, const float* flow01, const float* flow10,
                                    void* out_nhwc16, int H, int W, void* stream);
DRBA_API int drba_gmfss_scale_flow(const float* flow, const float* metric, const float* tmap, float tscalar,
                                   int H, int W, int s, float* out_flow, float* out_metric, void* stream);
/* GMFSS_union (models/model_gmfss_union/GMFSS.py:118-152): holes of either warped ones-map set both warped timestep
 * maps to 1; then where t0/t1 > 25 side 1 takes side 2's values and where t1/t0 > 25 the other way round, on the
 * NHWC fp16 concat buffer [h][w][2C] = [side 1 | side 2] or on two NCHW fp32 tensors.  drba_unpack_nhwc_f16 with
 * do_clamp == 2 applies tanh(x) * 10 (union MetricNet head). */
DRBA_API int drba_gmfss_union_fix_timesteps(float* t0w, float* t1w, const float* g0, const float* g1, size_t n, void* stream);
DRBA_API int drba_gmfss_union_swap_nhwc_f16(void* x, int C, const float* t0, const float* t1, int h, int w, void* stream);
DRBA_API int drba_gmfss_union_swap_nchw_f32(float* a, float* b, int C, const float* t0, const float* t1, int h, int w, void* stream);

/* ---------------------------------------------------------------------------
 * GMFlow glue (csrc/gmflow.cu; models/gmflow/*).  Every dense contraction of GMFlow runs on
 * drba_conv_tc_program_f16 (convs, linears as 1x1 convs, attention / correlation as batched GEMMs); these are
 * the stages in between.  Token tensors are NHWC fp16 [B][h][w][C]; flows are planar fp32 [2][h][w].
 *   normalize_img   : (img - mean) / std, utils.py:58-70
 *   inorm_stats/apply: InstanceNorm2d (eps 1e-5, no affine) of backbone.py:14-43; stats = double [C][2] (sum, sum^2),
 *                     zero on entry; apply: out = relu?( skip(+its own IN) + relu?(IN(x)) )
 *   add_position    : windowed sine position encoding (utils.py:73-94), pos = fp32 [wh][ww][C]
 *   window_pack     : split into k x k attention windows, optionally after the half-window roll of the shifted
 *                     blocks (transformer.py:74-84); dst [B*k*k][rows_pad][C] or transposed [B*k*k][C][rows_pad]
 *   softmax_rows    : in-place row softmax of the scores [nwin][Lw][ld] incl. the shift mask (transformer.py:20-45)
 *   ln_residual     : out = src + LayerNorm(m) (m in window order when k > 0), or cat = [src | LayerNorm(m)]
 *   soft_readout    : out[row] = sum_j softmax(scale * S[row])_j * val[j] (val NULL: grid coordinates), matching.py:31-41
 *   local_match     : (2r+1)^2 correlation soft-argmax, matching.py:46-89;  local_propagate: transformer.py:366-409
 *   warp_feature    : bilinear zero-padded warp of a feature map (gmflow.py:117-123)
 *   upsampler_input / convex_upsample: gmflow.py:68-90 around the two-conv mask head
 * ------------------------------------------------------------------------- */
DRBA_API int drba_gmflow_normalize_img(const float* in, float* out, int H, int W, void* stream);
DRBA_API int drba_gmflow_inorm_stats(const void* x, int C, int H, int W, double* stats_zeroed, void* stream);
DRBA_API int drba_gmflow_inorm_apply(const void* x, const double* stats, int relu_x, const void* skip, const double* skip_stats,
                                     int final_relu, void* out, int C, int H, int W, void* stream);
DRBA_API int drba_gmflow_add_position(void* x, const float* pos, int B, int h, int w, int wh, int ww, int C, void* stream);
DRBA_API int drba_gmflow_window_pack(const void* src, void* dst, int B, int h, int w, int C, int k, int shifted, int rows_pad,
                                     int transposed, void* stream);
DRBA_API int drba_gmflow_softmax_rows(void* S, int nwin_total, int Lw, int ld, int shifted, int k, int h, int w, void* stream);
DRBA_API int drba_gmflow_ln_residual(const void* src, const void* m, const float* gamma, const float* beta, void* out,
                                     int B, int h, int w, int C, int k, int shifted, int rows_pad, int cat, void* stream);
DRBA_API int drba_gmflow_soft_readout(const void* S, int rows, int cols, int ld, const float* val, int w, int subtract_grid,
                                      float scale, float* out, void* stream);
DRBA_API int drba_gmflow_local_match(const void* f0, const void* f1, int h, int w, int C, int radius, float* flow, void* stream);
DRBA_API int drba_gmflow_local_propagate(const void* q, const void* kmap, const float* flow, int h, int w, int C, float* out, void* stream);
DRBA_API int drba_gmflow_warp_feature(const void* f, const float* flow, void* out, int h, int w, int C, void* stream);
DRBA_API int drba_gmflow_upsampler_input(const float* flow, const void* feat, void* out144, int h, int w, void* stream);
DRBA_API int drba_gmflow_convex_upsample(const void* mask144, const float* flow, float* out, int h, int w, void* stream);
DRBA_API int drba_axpby_f32(const float* a, float alpha, const float* b, float beta, float* out, size_t n, void* stream);

/* ---------------------------------------------------------------------------
 * Fused non-conv stages of IFNet.forward (IFNet_HDv3.py:126-177), batch 1.
 * Feature maps f0/f1: [H][W][16] (NHWC) of feat_dtype.  The only full-resolution state is
 * flow: [H][W][4] fp32.  "tmp" = a block's lastconv output, 13 channels at 1/s resolution:
 *   tmp_layout 0: ConvTranspose output NCHW fp32 [52][H/2s][W/2s] (exact engine)
 *   tmp_layout 1: pixel-shuffled NHWC fp32 [H/s][W/s][16]       (tensor-core engine)
 *   tmp_layout 2: the same with 8 floats per pixel (flow + mask + 3 feat channels: what flow_accum / blend read;
 *                 written by the last IFBlock's lastconv with out_cstride = 8; not accepted by drba_ifnet_assemble)
 *
 * drba_ifnet_assemble: conv input of one IFBlock at 1/s resolution = warp + cat + resize
 *   (IFNet_HDv3.py:151-155 + :85-88).  flow == NULL: first block (39 channels, no warp).
 *   Otherwise mask/feat are taken from tmp_prev (x s_prev bilinear, IFNet_HDv3.py:92-96).
 *   out_dtype DRBA_F32: NCHW [52|39][H/s][W/s] in the reference's channel order;
 *   out_dtype DRBA_F16: NHWC [H/s][W/s][out_cstride] in the packed order
 *     [f0 16 | f1 16 | img0 3, img1 3, (first block: timestep), 0.. | timestep, mask, feat 8, flow 4, 0, 0]
 *     (48 channels for the first block, 64 otherwise; conv weights are permuted to match).
 * drba_ifnet_flow_accum: flow (+)= s * up(tmp[0:4]) (IFNet_HDv3.py:91-93, :157); `planar`
 *   (optional) also receives the result as [4][H][W] (RIFE.calc_flow needs planar flows).
 * drba_ifnet_blend: last flow update + warps + sigmoid blend (IFNet_HDv3.py:156-167):
 *   out[3][H][W] = warp(img0, F[:2]) * m + warp(img1, F[2:4]) * (1 - m),
 *   F = flow + s * up(tmp[0:4]) (flow may be NULL), m = sigmoid(up(tmp[4])).
 * ------------------------------------------------------------------------- */
DRBA_API int drba_ifnet_assemble(const float* img0, const float* img1, const void* f0, const void* f1, int feat_dtype,
                                 const float* timestep, float timestep_scalar,
                                 const float* flow, const float* tmp_prev, int tmp_layout, int s_prev,
                                 void* out, int out_dtype, int out_cstride, int H, int W, int s, void* stream);
/* Tensor-core engine only (NHWC fp16 features in, [H/s][W/s][64] fp16 out, lastconv outputs NHWC fp32 pitch 16):
 * the block input with the flow given as a SUM of up-sampled lastconv outputs, s_term0 * up(term0)[0:4]
 * (+ s_term1 * up(term1)[0:4]), evaluated at the block's own sample positions -- the coarse blocks (scale 8, 4) read
 * 1/16 and 1/4 of the pixels, so the full-resolution flow state need not exist for them; drba_ifnet_flow_sum then
 * writes flow = sum of up to three terms in one pass (the state blocks 3, 4 and the blend read).  Same sums, same
 * order as consecutive drba_ifnet_flow_accum calls (IFNet_HDv3.py:91-93, :157). */
DRBA_API int drba_ifnet_assemble_terms(const float* img0, const float* img1, const void* f0, const void* f1,
                                       const float* timestep, float timestep_scalar,
                                       const float* term0, int s_term0, const float* term1, int s_term1,
                                       const float* tmp_prev, int s_prev, void* out, int H, int W, int s, void* stream);
/* Tensor-core engine, blocks 3 and 4 (block scale s = 2 / 1): the block input is assembled INSIDE the kernel that runs the
 * block's first conv (conv0a: 3x3, stride 2, 64 packed channels -> cout = 16 / 32, LeakyReLU 0.2; IFNet_HDv3.py:66-69), so the
 * 64-channel input (267 MB per frame at 1088 x 1920, scale 1) is never written to memory.  Same arithmetic as
 * drba_ifnet_assemble (DRBA_F16 output) followed by drba_conv_tc_f16; replaces both for these blocks.
 *   jobs[k].out: [H/2s][W/2s][cout] fp16;  w: [9][cout][64] fp16 (taps ky*3+kx, packed input channel order);  bias [cout] fp32;
 *   tmp_prev: previous block's lastconv output, NHWC fp32, 16 floats per pixel, at 1/s_prev.  Up to two jobs per launch. */
typedef struct drba_ifnet_block_input {
    const float* img0; const float* img1;       /* [3][H][W] fp32 */
    const void* f0; const void* f1;             /* [H][W][16] fp16 (drba Head output) */
    const float* timestep; float timestep_scalar; /* [H][W] fp32, or NULL -> scalar */
    const float* flow;                          /* [H][W][4] fp32 flow state */
    const float* tmp_prev; int s_prev;
    void* out;
} drba_ifnet_block_input;
DRBA_API int drba_ifnet_block_conv0a_f16(const drba_ifnet_block_input* jobs, int njobs, const void* w, const float* bias, int cout,
                                         int H, int W, int s, void* stream);
DRBA_API int drba_ifnet_flow_sum(const float* tmp0, int s0, const float* tmp1, int s1, const float* tmp2, int s2, int nterms,
                                 float* flow, int H, int W, void* stream);
DRBA_API int drba_ifnet_flow_accum(const float* tmp, int tmp_layout, int s, float* flow, float* planar, int accumulate,
                                   int H, int W, void* stream);
DRBA_API int drba_ifnet_blend(const float* img0, const float* img1, const float* flow, const float* tmp, int tmp_layout,
                              int s, float* out, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRBA_B200_H */
