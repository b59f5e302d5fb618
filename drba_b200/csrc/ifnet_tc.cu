// Block-input assembly of the tensor-core engine (NHWC fp16 in / out).  Same arithmetic as ifnet_assemble_kernel in
// ifnet.cu, but this file is built with FMA contraction (the fp32 engine, which is compared with the fp32 oracle op
// by op, stays in ifnet.cu under -fmad=false): the result is rounded to fp16 anyway, and the kernel was
// instruction-issue bound (ncu: 77 % issue active) with a separate multiply and add per tap and channel.
#include "ifnet_common.cuh"
#ifndef DRBA_ASM_MINB4
#define DRBA_ASM_MINB4 10
#endif
#ifndef DRBA_ASM_MINB1
#define DRBA_ASM_MINB1 12
#endif

namespace drba {

// ---- block input assembly, tensor-core engine (NHWC fp16 in / out), L1-friendly lane mapping ----------
// The one-lane-per-pixel kernel above is bound by the L1 data stage (ncu: l1tex 68-86 %): with a lane
// stride of 32-128 B every LDG.128 / STG.128 touches 8-32 lines.  Here adjacent lanes cover adjacent
// bytes: a lane pair shares a pixel's 32 B of features, x-adjacent sample positions sit in adjacent
// lanes (the 2x2 mean becomes two shuffles, same summation order as acc_pos / mean_pos), and the
// 128 B/pixel output row is staged in shared memory and stored as full lines.  Results are bit-identical
// to ifnet_assemble_kernel<__half, true, 1>.
__device__ __forceinline__ void load8(const __half* __restrict__ p, float* v)
{
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        v[k * 2] = f.x; v[k * 2 + 1] = f.y;
    }
}

__device__ __forceinline__ void sample_feat8(const __half* __restrict__ f, const WarpTap& t, float* out)
{
    float v0[8], v1[8], v2[8], v3[8];
    load8(f + (size_t)t.i00 * 16, v0);
    load8(f + (size_t)max(t.i01, 0) * 16, v1);
    load8(f + (size_t)max(t.i10, 0) * 16, v2);
    load8(f + (size_t)max(t.i11, 0) * 16, v3);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float a = 0.0f + v0[c] * t.w00;
        if (t.i01 >= 0) a += v1[c] * t.w01;
        if (t.i10 >= 0) a += v2[c] * t.w10;
        if (t.i11 >= 0) a += v3[c] * t.w11;
        out[c] = a;
    }
}

__device__ __forceinline__ uint4 pack8(const float* v)
{
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    return o;
}

constexpr int kAsmTile = 32;   // output pixels per CTA

// 16 B chunk c of tile pixel px; rotated by 2*px so that pair / quad writers and the 8-lane row readers
// spread over the banks
__device__ __forceinline__ int tile_slot(int px, int c) { return px * 8 + ((c + 2 * px) & 7); }

// (min blocks per SM: the kernel is latency / L1 bound -- ncu r2: 42 % occupancy at 72 registers, issue 66 %, L1 wavefronts 65 % --
// so registers are capped for occupancy: 48 (s >= 2) / 40 (s = 1) measured best: block-input time 0.727 -> 0.691 ms per window)
template <int NP>
__global__ void __launch_bounds__(kIfThreads, NP == 4 ? DRBA_ASM_MINB4 : DRBA_ASM_MINB1)
ifnet_assemble_v2_kernel(const AssembleParams p)
{
    __shared__ __align__(16) uint4 tile[kAsmTile * 8];
    // Flow given as a SUM of up-sampled lastconv outputs (coarse blocks, NP = 4): evaluated ONCE per sample position
    // -- 32 pixels x 4 positions = one per thread -- and shared; each of the four roles used to re-evaluate it (two
    // bilinear up-samplings = 24 loads per position: block2's kernel took 57 us for 130 k output pixels).
    __shared__ __align__(16) float4 sflow[NP == 4 ? kAsmTile * 4 : 1];
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grid = (ceil(w / 32), h): a CTA owns 32 x-adjacent pixels of ONE output row, so no pass divides by the row length
    // (a signed division is ~25 instructions, seven passes per CTA, and the kernel is issue-bound)
    const int Y = blockIdx.y, X0 = blockIdx.x * kAsmTile;
    const int off = NP == 1 ? 0 : p.s / 2 - 1;
    const size_t HW = (size_t)p.H * p.W;
    constexpr int PXW = NP == 1 ? 16 : 8;      // output pixels per warp pass
    const bool shared_flow = NP == 4 && p.nfterms > 0;
    if (shared_flow) {
        const int px = threadIdx.x >> 2, K = threadIdx.x & 3;
        const int X = min(X0 + px, p.w - 1);
        sflow[threadIdx.x] = load_flow<1>(p, p.s * Y + off + (K >> 1), p.s * X + off + (K & 1));
        __syncthreads();
    }
    // flow at sample position K = 2 * (row offset) + (column offset) of tile pixel px
    auto flow_at = [&](int px, int K, int y, int x) -> float4 {
        return shared_flow ? sflow[px * 4 + K] : load_flow<1>(p, y, x);
    };

    if (role < 2) {
        // lane = [pixel | x position | channel half]
        const __half* f = reinterpret_cast<const __half*>(role == 0 ? p.f0 : p.f1) + (lane & 1) * 8;
        const int xpos = NP == 1 ? 0 : (lane >> 1) & 1;
        const int pl = NP == 1 ? lane >> 1 : lane >> 2;
#pragma unroll 1
        for (int pass = 0; pass < kAsmTile / PXW; ++pass) {
            const int px = pass * PXW + pl;
            const int X = min(X0 + px, p.w - 1);
            const int x = p.s * X + off + xpos, y = p.s * Y + off;
            float r0[8];
            {
                const float4 fl = flow_at(px, xpos, y, x);
                const WarpTap t = warp_tap(x, y, role == 0 ? fl.x : fl.z, role == 0 ? fl.y : fl.w, p.H, p.W);
                sample_feat8(f, t, r0);
            }
            if (NP == 4) {
                float r1[8];
                const float4 fl = flow_at(px, 2 + xpos, y + 1, x);
                const WarpTap t = warp_tap(x, y + 1, role == 0 ? fl.x : fl.z, role == 0 ? fl.y : fl.w, p.H, p.W);
                sample_feat8(f, t, r1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float a0 = r0[c] + __shfl_xor_sync(0xffffffffu, r0[c], 2);
                    const float a1 = r1[c] + __shfl_xor_sync(0xffffffffu, r1[c], 2);
                    r0[c] = (a0 + a1) * 0.25f;
                }
            }
            if (xpos == 0) tile[tile_slot(px, role * 2 + (lane & 1))] = pack8(r0);
        }
    } else if (role == 2) {
        // lane = [image | pixel | x position]
        const int sel = lane >> 4;
        const int xpos = NP == 1 ? 0 : lane & 1;
        const int pl = NP == 1 ? lane & 15 : (lane >> 1) & 7;
        const float* img = sel ? p.img1 : p.img0;
#pragma unroll 1
        for (int pass = 0; pass < kAsmTile / PXW; ++pass) {
            const int px = pass * PXW + pl;
            const int X = min(X0 + px, p.w - 1);
            const int x = p.s * X + off + xpos, y = p.s * Y + off;
            float r0[3];
            {
                const float4 fl = flow_at(px, xpos, y, x);
                const WarpTap t = warp_tap(x, y, sel ? fl.z : fl.x, sel ? fl.w : fl.y, p.H, p.W);
#pragma unroll
                for (int c = 0; c < 3; ++c) r0[c] = sample_plane(img + (size_t)c * HW, t);
            }
            if (NP == 4) {
                float r1[3];
                const float4 fl = flow_at(px, 2 + xpos, y + 1, x);
                const WarpTap t = warp_tap(x, y + 1, sel ? fl.z : fl.x, sel ? fl.w : fl.y, p.H, p.W);
#pragma unroll
                for (int c = 0; c < 3; ++c) r1[c] = sample_plane(img + (size_t)c * HW, t);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float a0 = r0[c] + __shfl_xor_sync(0xffffffffu, r0[c], 1);
                    const float a1 = r1[c] + __shfl_xor_sync(0xffffffffu, r1[c], 1);
                    r0[c] = (a0 + a1) * 0.25f;
                }
            }
            if (xpos == 0) {
                __half* dst = reinterpret_cast<__half*>(&tile[tile_slot(px, 4)]);
#pragma unroll
                for (int c = 0; c < 3; ++c) dst[sel * 3 + c] = __float2half_rn(r0[c]);
                if (sel) *reinterpret_cast<uint32_t*>(dst + 6) = 0u;
                else tile[tile_slot(px, 5)] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    } else {
        // lane = [pixel | sample position]; timestep, mask, feat (x s_prev up-sampling of the previous
        // lastconv output) and flow / s
        constexpr int PXM = NP == 1 ? 32 : 8;
        const int K = NP == 1 ? 0 : lane & 3;
        const int pl = NP == 1 ? lane : lane >> 2;
        const float inv = 1.0f / (float)p.s;   // IFNet_HDv3.py:87: interpolate(flow) * 1. / scale
#pragma unroll 1
        for (int pass = 0; pass < kAsmTile / PXM; ++pass) {
            const int px = pass * PXM + pl;
            const int X = min(X0 + px, p.w - 1);
            const int x = p.s * X + off + (K & 1), y = p.s * Y + off + (K >> 1);
            float v[16];
            const float4 fl = flow_at(px, K, y, x);
            v[0] = p.timestep ? p.timestep[(size_t)y * p.W + x] : p.timestep_scalar;
            up_tmp<1, 4, 9>(p.prev, y, x, v + 1);
            v[10] = fl.x; v[11] = fl.y; v[12] = fl.z; v[13] = fl.w;
            if (NP == 4) {
#pragma unroll
                for (int c = 0; c < 14; ++c) {
                    const float a = v[c] + __shfl_xor_sync(0xffffffffu, v[c], 1);
                    v[c] = (a + __shfl_xor_sync(0xffffffffu, a, 2)) * 0.25f;
                }
            }
#pragma unroll
            for (int c = 10; c < 14; ++c) v[c] = v[c] * 1.0f * inv;
            v[14] = 0.0f; v[15] = 0.0f;
            if (K == 0) {
                tile[tile_slot(px, 6)] = pack8(v);
                tile[tile_slot(px, 7)] = pack8(v + 8);
            }
        }
    }
    __syncthreads();
    uint4* out4 = reinterpret_cast<uint4*>(p.out) + ((size_t)Y * p.w + X0) * 8;
#pragma unroll
    for (int k = 0; k < kAsmTile * 8 / kIfThreads; ++k) {
        const int i = threadIdx.x + k * kIfThreads;
        const int px = i >> 3, c = i & 7;
        if (X0 + px < p.w) out4[(size_t)px * 8 + c] = tile[tile_slot(px, c)];
    }
}

// (tensor-core engine only, hence in this FMA-contracting file: the kernel is issue-bound on three bilinear up-samplings per pixel)
// flow = s0 * up(tmp0) + s1 * up(tmp1) + s2 * up(tmp2), summed in block order: what three consecutive
// ifnet_flow_accum launches leave behind, in ONE write-only pass (the coarse blocks' assemble kernels evaluate
// the flow at their own sample positions, so the first full-resolution flow is needed before block 3 only)
__global__ void __launch_bounds__(256)
ifnet_flow_sum_kernel(const Tmp13 t0, const Tmp13 t1, const Tmp13 t2, int nterms, float* __restrict__ flow, int H, int W)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= H * W) return;
    const int y = idx / W, x = idx - y * W;
    float o[4];
    up_tmp<1, 0, 4>(t0, y, x, o);
    float fs = (float)t0.s;
    float4 f = make_float4(o[0] * fs, o[1] * fs, o[2] * fs, o[3] * fs);
    if (nterms > 1) {
        up_tmp<1, 0, 4>(t1, y, x, o);
        fs = (float)t1.s;
        f.x = f.x + o[0] * fs; f.y = f.y + o[1] * fs; f.z = f.z + o[2] * fs; f.w = f.w + o[3] * fs;
    }
    if (nterms > 2) {
        up_tmp<1, 0, 4>(t2, y, x, o);
        fs = (float)t2.s;
        f.x = f.x + o[0] * fs; f.y = f.y + o[1] * fs; f.z = f.z + o[2] * fs; f.w = f.w + o[3] * fs;
    }
    reinterpret_cast<float4*>(flow)[idx] = f;
}

void launch_flow_sum_tc(const Tmp13& t0, const Tmp13& t1, const Tmp13& t2, int nterms, float* flow, int H, int W, cudaStream_t st)
{
    ifnet_flow_sum_kernel<<<cdiv((size_t)H * W, 256), 256, 0, st>>>(t0, t1, t2, nterms, flow, H, W);
}

void launch_assemble_tc(const AssembleParams& p, cudaStream_t st)
{
    const dim3 grid(cdiv((size_t)p.w, kAsmTile), (unsigned)p.h);
    if (p.s == 1) ifnet_assemble_v2_kernel<1><<<grid, kIfThreads, 0, st>>>(p);
    else ifnet_assemble_v2_kernel<4><<<grid, kIfThreads, 0, st>>>(p);
}

}  // namespace drba
