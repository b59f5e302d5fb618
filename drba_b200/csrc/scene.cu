// Scene-cut detection: check_scene(x1, x2, threshold) of models/utils/tools.py:27-30 in ONE kernel.
//
//   x1, x2 -> F.interpolate(x, (32, 32), bilinear, align_corners=False)                       (tools.py:28-29)
//          -> ssim_matlab: 3-D Gaussian window (11^3, sigma 1.5) over the (C, H, W) volume with replicate padding 5,
//             SSIM map, mean                                            (models/pytorch_msssim/__init__.py:83-136)
//          -> flag = ssim < threshold                                                         (tools.py:30)
//
// The reference spends ~25 launches and THREE host synchronisations per frame pair on this (torch.max / torch.min of
// the value-range probe and the final comparison).  Here one CTA per pair does everything in shared memory and
// writes {ssim, flag} to device memory (or mapped pinned host memory): the host reads the flag a window later
// without ever stalling the stream (drba_b200.tools.SceneDetector).
//
// The Gaussian volume window is separable and replicate padding is a per-axis clamp, so three 1-D passes (W, H, C)
// equal the reference's conv3d up to fp32 summation order (measured: |dssim| <= 2e-6).  The work is 2 x 3 x 32 x 32
// samples: latency, not bandwidth -- a single CTA, 1024 threads, three voxels per thread.
#include "common.cuh"

namespace drba {

constexpr int kSceneThreads = 1024;
constexpr int kThumb = 32;
constexpr int kVox = 3 * kThumb * kThumb;      // 3072 voxels of the (C, H, W) volume

struct SceneGauss { float g[11]; };

__device__ __forceinline__ float thumb_sample(const float* __restrict__ s, int H, int W, int oy, int ox, float rh, float rw)
{
    // ATen upsample_bilinear2d, align_corners = False (same arithmetic as resize_bilinear_kernel, sample.cu)
    float sy = rh * ((float)oy + 0.5f) - 0.5f, sx = rw * ((float)ox + 0.5f) - 0.5f;
    if (sy < 0.0f) sy = 0.0f;
    if (sx < 0.0f) sx = 0.0f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.0f - ly;
    const float lx = sx - (float)x0, hx = 1.0f - lx;
    return hy * (hx * s[(size_t)y0 * W + x0] + lx * s[(size_t)y0 * W + x1]) + ly * (hx * s[(size_t)y1 * W + x0] + lx * s[(size_t)y1 * W + x1]);
}

// one CTA per frame pair: pair i = (x1 + i * stride1, x2 + i * stride2), frames NCHW [3][H][W] fp32
__global__ void __launch_bounds__(kSceneThreads)
check_scene_kernel(const float* __restrict__ x1s, const float* __restrict__ x2s, long long stride1, long long stride2, int H, int W,
                   float rh, float rw, float threshold, SceneGauss gw, float* __restrict__ ssim_out, int* __restrict__ flag_out)
{
    extern __shared__ float sm[];
    float* t1 = sm;                  // thumbnails [3][32][32]
    float* t2 = sm + kVox;
    float* A = sm + 2 * kVox;        // blur ping-pong
    float* B = sm + 3 * kVox;
    __shared__ float red[32];
    __shared__ float s_max, s_min;
    const int tid = threadIdx.x;
    const float* x1 = x1s + (long long)blockIdx.x * stride1;
    const float* x2 = x2s + (long long)blockIdx.x * stride2;
    const size_t HW = (size_t)H * W;

    float vmax = -INFINITY, vmin = INFINITY;
    for (int v = tid; v < kVox; v += kSceneThreads) {
        const int c = v >> 10, oy = (v >> 5) & 31, ox = v & 31;
        const float a = thumb_sample(x1 + c * HW, H, W, oy, ox, rh, rw);
        t1[v] = a;
        t2[v] = thumb_sample(x2 + c * HW, H, W, oy, ox, rh, rw);
        vmax = fmaxf(vmax, a);
        vmin = fminf(vmin, a);
    }
    // value range probe of img1 (pytorch_msssim/__init__.py:85-94)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    if ((tid & 31) == 0) { red[tid >> 5] = vmax; }
    __syncthreads();
    if (tid < 32) {
        float m = red[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (tid == 0) s_max = m;
    }
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = vmin; }
    __syncthreads();
    if (tid < 32) {
        float m = red[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (tid == 0) s_min = m;
    }
    __syncthreads();
    const float Lr = (s_max > 128.0f ? 255.0f : 1.0f) - (s_min < -0.5f ? -1.0f : 0.0f);

    // five blurred volumes: x1, x2, x1*x1, x2*x2, x1*x2 -- each: load -> W pass -> H pass -> C pass (registers)
    float q[5][3];
#pragma unroll
    for (int qi = 0; qi < 5; ++qi) {
        for (int v = tid; v < kVox; v += kSceneThreads) {
            const float a = t1[v], b = t2[v];
            A[v] = qi == 0 ? a : (qi == 1 ? b : (qi == 2 ? a * a : (qi == 3 ? b * b : a * b)));
        }
        __syncthreads();
        for (int v = tid; v < kVox; v += kSceneThreads) {          // along W
            const int x = v & 31, base = v & ~31;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 11; ++k) acc += gw.g[k] * A[base + min(max(x + k - 5, 0), 31)];
            B[v] = acc;
        }
        __syncthreads();
        for (int v = tid; v < kVox; v += kSceneThreads) {          // along H
            const int x = v & 31, y = (v >> 5) & 31, cb = v & ~1023;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 11; ++k) acc += gw.g[k] * B[cb + (min(max(y + k - 5, 0), 31) << 5) + x];
            A[v] = acc;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 3; ++j) {                              // along C (3 channels, replicate-padded to 13)
            const int v = tid + j * kSceneThreads;
            const int c = v >> 10, r = v & 1023;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 11; ++k) acc += gw.g[k] * A[(min(max(c + k - 5, 0), 2) << 10) + r];
            q[qi][j] = acc;
        }
        __syncthreads();
    }
    const float C1 = (0.01f * Lr) * (0.01f * Lr), C2 = (0.03f * Lr) * (0.03f * Lr);
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float mu1 = q[0][j], mu2 = q[1][j];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = q[2][j] - mu1_sq, s2 = q[3][j] - mu2_sq, s12 = q[4][j] - mu12;
        const float v1 = 2.0f * s12 + C2, v2 = s1 + s2 + C2;
        sum += ((2.0f * mu12 + C1) * v1) / ((mu1_sq + mu2_sq + C1) * v2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0) red[tid >> 5] = sum;
    __syncthreads();
    if (tid < 32) {
        float m = red[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
        if (tid == 0) {
            const float ssim = m / (float)kVox;
            if (ssim_out) ssim_out[blockIdx.x] = ssim;
            if (flag_out) flag_out[blockIdx.x] = ssim < threshold ? 1 : 0;
            __threadfence_system();      // the outputs may live in mapped host memory
        }
    }
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_check_scene_f32(const float* x1, const float* x2, long long pair_stride1, long long pair_stride2, int npairs,
                         int H, int W, float threshold, float* ssim_out, int* flag_out, void* stream)
{
    if (!x1 || !x2 || (!ssim_out && !flag_out)) return DRBA_E_ARG;
    if (npairs < 1 || H < 1 || W < 1) return DRBA_E_ARG;
    // gaussian(11, 1.5) as the reference builds it: float32 exps, normalised in fp32 (pytorch_msssim/__init__.py:9-11)
    SceneGauss gw;
    float s = 0.0f;
    for (int k = 0; k < 11; ++k) { gw.g[k] = (float)exp(-(double)((k - 5) * (k - 5)) / (2.0 * 1.5 * 1.5)); s += gw.g[k]; }
    for (int k = 0; k < 11; ++k) gw.g[k] /= s;
    static bool attr = false;
    const size_t smem = (size_t)4 * kVox * sizeof(float);
    if (!attr) {
        cudaFuncSetAttribute(check_scene_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    check_scene_kernel<<<npairs, kSceneThreads, smem, as_stream(stream)>>>(x1, x2, pair_stride1, pair_stride2, H, W, (float)H / 32.0f,
                                                                           (float)W / 32.0f, threshold, gw, ssim_out, flag_out);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
