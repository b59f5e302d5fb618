// Shared helpers for the libdrba_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include "../../include/drba_b200.h"

#define DRBA_RETURN_IF_LAUNCH_FAILED()                 \
    do {                                               \
        cudaError_t e__ = cudaGetLastError();          \
        if (e__ != cudaSuccess) return (int)e__;       \
    } while (0)

namespace drba {

constexpr int kNumSMs = 148;  // B200

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// vector reductions to global memory (REDG.E.ADD.F32x4 / F32x2): one L2 atomic
// transaction per 16 / 8 bytes instead of one per float
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// Bilinear forward-warp footprint of one source pixel (softsplat.py:312-357).
struct Footprint {
    int x0, y0;
    float nw, ne, sw, se;
    bool ok;  // finite target (softsplat.py:323-324) that fits an int32
};

__device__ __forceinline__ Footprint footprint(int x, int y, float fx, float fy) {
    Footprint f;
    const float X = (float)x + fx;
    const float Y = (float)y + fy;
    const float flx = floorf(X), fly = floorf(Y);
    f.ok = isfinite(X) && isfinite(Y) && fabsf(flx) < 2147483000.0f && fabsf(fly) < 2147483000.0f;
    f.x0 = f.ok ? (int)flx : 0;
    f.y0 = f.ok ? (int)fly : 0;
    const float wx1 = X - (float)f.x0, wx0 = (float)(f.x0 + 1) - X;
    const float wy1 = Y - (float)f.y0, wy0 = (float)(f.y0 + 1) - Y;
    f.nw = wx0 * wy0;
    f.ne = wx1 * wy0;
    f.sw = wx0 * wy1;
    f.se = wx1 * wy1;
    return f;
}

// splat denominators: softsplat.py:273-290
__device__ __forceinline__ float splat_den(float d, int eps_mode) {
    if (eps_mode == DRBA_EPS_ADD) return d + 0.0000001f;
    if (eps_mode == DRBA_EPS_ZERO) return d == 0.0f ? 1.0f : d;
    if (eps_mode == DRBA_EPS_NONE) return d;
    return d < 0.0000001f ? 0.0000001f : d;
}

}  // namespace drba
