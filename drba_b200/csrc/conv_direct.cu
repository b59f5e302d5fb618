// fp32 direct convolution (CUDA cores) -- the exact-precision conv engine.
//
// Serves the fp32 mode of the IFNet path (models/rife_426_heavy/IFNet_HDv3.py: conv :11-25,
// Head :28-47, ResConv :50-59, IFBlock :62-96) so that the CUDA path can be compared with
// the fp32 oracle at 1e-4 level; the throughput engine is the tcgen05 implicit GEMM in
// conv_tc.cu.  One generic kernel covers every layer shape on the path:
//
//   out[n, co, oy*OS+PY, ox*OS+PX] = act( bias[co] + res[...] +
//        sum_{t<T} sum_{ci} in[n, ci, oy*S + dy[t], ox*S + dx[t]] * w[t][ci][co] )
//
//   3x3 s1 p1   : T=9, S=1, d = k-1, OS=1
//   3x3 s2 p1   : T=9, S=2, d = k-1, OS=1
//   ConvT k4 s2 p1 (IFNet_HDv3.py:80, :34) : four phase launches (PY,PX in {0,1}), T=4, S=1,
//                 OS=2; phase 0 uses kernel rows {1,3} at dy {0,-1}, phase 1 rows {0,2} at dy {+1,0}
//   ResConv     : beta is folded into w and bias on the host; res = x; act = LeakyReLU(0.2)
//
// Input and output are addressed through element strides, so NCHW and NHWC tensors (and
// channel slices of larger tensors) are all served without copies.
#include "common.cuh"

namespace drba {

constexpr int kTile = 16;      // 16x16 output positions per CTA
constexpr int kCoT = 16;       // output channels per thread
constexpr int kCiT = 8;        // input channels staged per iteration
constexpr int kMaxTaps = 49;      // up to 7x7 (GMFlow backbone.conv1)

struct DirectConvParams {
    const float* in; const float* w; const float* bias; const float* res; float* out;
    int N, Cin, H, W;                 // input extent
    long long in_sn, in_sc, in_sy, in_sx;
    int Cout, OH, OW;                 // output positions computed by this launch
    long long out_sn, out_sc, out_sy, out_sx;   // strides of the destination (also used for res)
    int S, OS, PY, PX;
    int T;
    int dy[kMaxTaps], dx[kMaxTaps];
    int min_dy, min_dx, PH, PW;       // staged patch geometry
    int act;                          // 0 none, 1 LeakyReLU(0.2)
    int out_half;                     // store fp16 instead of fp32 (res, if any, is then fp16 too)
};

__global__ void __launch_bounds__(kTile * kTile)
direct_conv_kernel(const DirectConvParams p)
{
    extern __shared__ float smem[];
    float* patch = smem;                                   // [kCiT][PH][PW]
    float* wsm = smem + kCiT * p.PH * p.PW;                // [T][kCiT][kCoT]
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kTile + tx;
    const int co_tiles = (p.Cout + kCoT - 1) / kCoT;
    const int n = blockIdx.z / co_tiles;
    const int co0 = (blockIdx.z - n * co_tiles) * kCoT;
    const int ox0 = blockIdx.x * kTile, oy0 = blockIdx.y * kTile;
    const int iy0 = oy0 * p.S + p.min_dy, ix0 = ox0 * p.S + p.min_dx;

    float acc[kCoT];
#pragma unroll
    for (int k = 0; k < kCoT; ++k) acc[k] = 0.0f;

    const float* inn = p.in + (long long)n * p.in_sn;
    const int patch_elems = p.PH * p.PW;
    for (int ci0 = 0; ci0 < p.Cin; ci0 += kCiT) {
        for (int e = tid; e < kCiT * patch_elems; e += kTile * kTile) {
            const int ci = e / patch_elems, r = e - ci * patch_elems;
            const int py = r / p.PW, px = r - py * p.PW;
            const int iy = iy0 + py, ix = ix0 + px, c = ci0 + ci;
            float v = 0.0f;
            if (c < p.Cin && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                v = inn[c * p.in_sc + iy * p.in_sy + ix * p.in_sx];
            patch[e] = v;
        }
        for (int e = tid; e < p.T * kCiT * kCoT; e += kTile * kTile) {
            const int t = e / (kCiT * kCoT), r = e - t * (kCiT * kCoT);
            const int ci = r / kCoT, co = r - ci * kCoT;
            float v = 0.0f;
            if (ci0 + ci < p.Cin && co0 + co < p.Cout)
                v = p.w[((long long)t * p.Cin + ci0 + ci) * p.Cout + co0 + co];
            wsm[e] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int t = 0; t < p.T; ++t) {
            const int off = (ty * p.S + p.dy[t] - p.min_dy) * p.PW + tx * p.S + p.dx[t] - p.min_dx;
#pragma unroll
            for (int ci = 0; ci < kCiT; ++ci) {
                const float v = patch[ci * patch_elems + off];
                const float4* w4 = reinterpret_cast<const float4*>(wsm + (t * kCiT + ci) * kCoT);
#pragma unroll
                for (int k4 = 0; k4 < kCoT / 4; ++k4) {
                    const float4 w = w4[k4];
                    acc[k4 * 4 + 0] = fmaf(v, w.x, acc[k4 * 4 + 0]);
                    acc[k4 * 4 + 1] = fmaf(v, w.y, acc[k4 * 4 + 1]);
                    acc[k4 * 4 + 2] = fmaf(v, w.z, acc[k4 * 4 + 2]);
                    acc[k4 * 4 + 3] = fmaf(v, w.w, acc[k4 * 4 + 3]);
                }
            }
        }
        __syncthreads();
    }

    const int ox = ox0 + tx, oy = oy0 + ty;
    if (ox >= p.OW || oy >= p.OH) return;
    const long long o = (long long)n * p.out_sn + (long long)(oy * p.OS + p.PY) * p.out_sy + (long long)(ox * p.OS + p.PX) * p.out_sx;
#pragma unroll
    for (int k = 0; k < kCoT; ++k) {
        const int co = co0 + k;
        if (co >= p.Cout) break;
        float v = acc[k] + (p.bias ? p.bias[co] : 0.0f);
        if (p.res) v += p.out_half ? __half2float(reinterpret_cast<const __half*>(p.res)[o + co * p.out_sc]) : p.res[o + co * p.out_sc];
        if (p.act == 1) v = v > 0.0f ? v : 0.2f * v;
        if (p.out_half) reinterpret_cast<__half*>(p.out)[o + co * p.out_sc] = __float2half_rn(v);
        else p.out[o + co * p.out_sc] = v;
    }
}


// ---- Head.cnn0 of the tensor-core engine (IFNet_HDv3.py:31: Conv2d(3, 16, 3, 2, 1) + LeakyReLU) ------------------
// Planar fp32 image in, NHWC fp16 out.  The generic kernel above stages 8-channel patches and 16x16 tiles for a
// layer that has 27 inputs per output: 97 us at 1080p against ~42 MB of traffic.  Here one thread owns two
// vertically adjacent outputs (weights are read once for both, broadcast from shared memory), a warp covers 32
// consecutive columns so every image row segment is read as whole lines.  Same fmaf order as the generic kernel
// (tap-major, input channel inner), so the results are identical.
__global__ void __launch_bounds__(256)
head_conv0_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  __half* __restrict__ out, int H, int W, long long in_sn, long long in_sc, long long in_sy,
                  int OH, int OW, long long out_sn, long long out_sy)
{
    __shared__ __align__(16) float wsm[27 * 16];
    __shared__ float bsm[16];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int e = tid; e < 27 * 16; e += 256) wsm[e] = w[e];
    if (tid < 16) bsm[tid] = bias ? bias[tid] : 0.0f;
    __syncthreads();
    const int ox = blockIdx.x * 32 + threadIdx.x;
    const int oy = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (ox >= OW || oy >= OH) return;
    const float* inn = in + (long long)blockIdx.z * in_sn;
    float acc[2][16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { acc[0][k] = 0.0f; acc[1][k] = 0.0f; }
    // input rows 2*oy - 1 .. 2*oy + 3: output oy uses rows 0..2 of them, output oy + 1 rows 2..4
    float v[5][3][3];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const int iy = 2 * oy - 1 + r;
        const bool yok = iy >= 0 && iy < H;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                v[r][ci][kx] = (yok && ix >= 0 && ix < W) ? inn[ci * in_sc + iy * in_sy + ix] : 0.0f;
            }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const float4* w4 = reinterpret_cast<const float4*>(wsm + (t * 3 + ci) * 16);
            const float a = v[t / 3][ci][t % 3], b = v[t / 3 + 2][ci][t % 3];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const float4 ww = w4[k4];
                acc[0][k4 * 4 + 0] = fmaf(a, ww.x, acc[0][k4 * 4 + 0]); acc[1][k4 * 4 + 0] = fmaf(b, ww.x, acc[1][k4 * 4 + 0]);
                acc[0][k4 * 4 + 1] = fmaf(a, ww.y, acc[0][k4 * 4 + 1]); acc[1][k4 * 4 + 1] = fmaf(b, ww.y, acc[1][k4 * 4 + 1]);
                acc[0][k4 * 4 + 2] = fmaf(a, ww.z, acc[0][k4 * 4 + 2]); acc[1][k4 * 4 + 2] = fmaf(b, ww.z, acc[1][k4 * 4 + 2]);
                acc[0][k4 * 4 + 3] = fmaf(a, ww.w, acc[0][k4 * 4 + 3]); acc[1][k4 * 4 + 3] = fmaf(b, ww.w, acc[1][k4 * 4 + 3]);
            }
        }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (oy + j >= OH) break;
        uint4 o[2];
        __half* oh = reinterpret_cast<__half*>(o);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float x = acc[j][k] + bsm[k];
            x = x > 0.0f ? x : 0.2f * x;
            oh[k] = __float2half_rn(x);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)blockIdx.z * out_sn + (long long)(oy + j) * out_sy + (long long)ox * 16);
        dst[0] = o[0]; dst[1] = o[1];
    }
}

}  // namespace drba

using namespace drba;

extern "C" {

int drba_conv2d_direct_f32(const float* in, const float* w, const float* bias, const void* res, void* out, int out_dtype,
                           int N, int Cin, int H, int W, const long long* in_strides,
                           int Cout, int OH, int OW, const long long* out_strides,
                           int S, int OS, int PY, int PX, int T, const int* dy, const int* dx,
                           int act, void* stream)
{
    if (N < 0 || Cin <= 0 || H <= 0 || W <= 0 || Cout <= 0 || OH < 0 || OW < 0) return DRBA_E_ARG;
    if (T <= 0 || T > kMaxTaps || S <= 0 || OS <= 0 || !dy || !dx || !in_strides || !out_strides) return DRBA_E_ARG;
    if (act != 0 && act != 1) return DRBA_E_ARG;
    if (out_dtype != DRBA_F32 && out_dtype != DRBA_F16) return DRBA_E_ARG;
    if ((size_t)N * OH * OW == 0) return DRBA_OK;
    if (!in || !w || !out) return DRBA_E_ARG;
    {
        // Head.cnn0 fast path: 3 -> 16 channels, 3x3 stride 2 pad 1, LeakyReLU, planar fp32 -> NHWC fp16
        bool canon = Cin == 3 && Cout == 16 && T == 9 && S == 2 && OS == 1 && PY == 0 && PX == 0 && act == 1 && !res &&
                     out_dtype == DRBA_F16 && in_strides[3] == 1 && out_strides[1] == 1 && out_strides[3] == 16 &&
                     out_strides[2] % 8 == 0 && out_strides[0] % 8 == 0 && aligned16(out) && N <= 65535;
        for (int t = 0; canon && t < 9; ++t) canon = dy[t] == t / 3 - 1 && dx[t] == t % 3 - 1;
        if (canon) { const char* e = getenv("DRBA_DIRECT_GENERIC"); canon = !(e && e[0] == '1'); }   // tests: force the generic kernel
        if (canon) {
            dim3 grid((OW + 31) / 32, (OH + 15) / 16, N);
            head_conv0_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, w, bias, (__half*)out, H, W, in_strides[0], in_strides[1],
                                                                           in_strides[2], OH, OW, out_strides[0], out_strides[2]);
            DRBA_RETURN_IF_LAUNCH_FAILED();
            return DRBA_OK;
        }
    }
    DirectConvParams p;
    p.in = in; p.w = w; p.bias = bias; p.res = (const float*)res; p.out = (float*)out; p.out_half = out_dtype == DRBA_F16;
    p.N = N; p.Cin = Cin; p.H = H; p.W = W;
    p.in_sn = in_strides[0]; p.in_sc = in_strides[1]; p.in_sy = in_strides[2]; p.in_sx = in_strides[3];
    p.Cout = Cout; p.OH = OH; p.OW = OW;
    p.out_sn = out_strides[0]; p.out_sc = out_strides[1]; p.out_sy = out_strides[2]; p.out_sx = out_strides[3];
    p.S = S; p.OS = OS; p.PY = PY; p.PX = PX; p.T = T; p.act = act;
    int mny = dy[0], mxy = dy[0], mnx = dx[0], mxx = dx[0];
    for (int t = 0; t < T; ++t) {
        p.dy[t] = dy[t]; p.dx[t] = dx[t];
        mny = dy[t] < mny ? dy[t] : mny; mxy = dy[t] > mxy ? dy[t] : mxy;
        mnx = dx[t] < mnx ? dx[t] : mnx; mxx = dx[t] > mxx ? dx[t] : mxx;
    }
    p.min_dy = mny; p.min_dx = mnx;
    p.PH = (kTile - 1) * S + (mxy - mny) + 1;
    p.PW = (kTile - 1) * S + (mxx - mnx) + 1;
    const size_t smem = sizeof(float) * ((size_t)kCiT * p.PH * p.PW + (size_t)T * kCiT * kCoT);
    if (smem > 200 * 1024) return DRBA_E_UNSUPPORTED;
    const int co_tiles = (Cout + kCoT - 1) / kCoT;
    if ((long long)N * co_tiles > 65535) return DRBA_E_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(direct_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid((OW + kTile - 1) / kTile, (OH + kTile - 1) / kTile, N * co_tiles);
    direct_conv_kernel<<<grid, dim3(kTile, kTile), smem, as_stream(stream)>>>(p);
    DRBA_RETURN_IF_LAUNCH_FAILED();
    return DRBA_OK;
}

}  // extern "C"
