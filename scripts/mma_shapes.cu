// Isolated tcgen05.mma throughput on B200 (sm_100a): cycles per instruction for kind::f16, operands resident in
// shared memory (SS mode), as a function of
//   M in {64, 128} (cta_group::1) / 256 (cta_group::2), N in {16 ... 256},
//   the operand swizzle (128B / 64B / 32B: Kc = 64 / 32 / 16 fp16 per row),
//   and the descriptor geometry the conv engine's halo modes use (csrc/conv_tc.cu): 8-row core groups 16 rows apart
//   (SBO = 16 rows) and start addresses shifted by whole rows, which move the swizzle phase.
// Settles DESIGN.md 4.1's "an M = 128 MMA costs >= 115 cycles for any N <= 64" claim with a direct measurement
// (VERDICT r1, item 4).  One CTA (or CTA pair) per SM on every SM; the elected thread issues R back-to-back MMAs
// into one accumulator, commits, waits; cycles = (t(R2) - t(R1)) / (R2 - R1) removes the fixed latency.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_shapes mma_shapes.cu && ./build/mma_shapes
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

struct Params {
    int M, N;            // instruction shape (M = 256 only with CG = 2)
    int swz_bytes;       // 128 / 64 / 32  (row of the K-major operand tile = swz_bytes)
    int sbo_rows;        // stride between 8-row core groups, in rows (8 = dense tile, 16 = halo tile)
    int a_shift_rows;    // start address of A shifted by this many rows (tap offsets of the halo modes)
    int b_shift_rows;
    int ksteps;          // K = 16 steps per "iteration" (walks the swizzle span like the engine: +32 B each)
    int r1, r2;          // MMAs per timed burst
    long long* out;      // [grid][2] cycles of burst r1, r2
};

// K-major swizzled smem descriptor (see csrc/conv_tc.cu make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int swz_bytes, int sbo_bytes) {
    const uint64_t layout = swz_bytes == 128 ? 2ull : (swz_bytes == 64 ? 4ull : 6ull);
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_shapes_kernel(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (smem - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5;
    uint32_t cta_rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

    // operands: small fp16 values (any data gives the same timing; avoid NaN / denormal patterns)
    for (int i = threadIdx.x; i < 120 * 1024 / 2; i += blockDim.x)
        reinterpret_cast<__half*>(base)[i] = __float2half(0.001f * (float)((i * 7) % 13 - 6));
    asm volatile("fence.proxy.async;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    const uint32_t tmem_d = tmem_base_smem;

    const int rowb = p.swz_bytes;                       // bytes per operand row
    const uint32_t a_addr = smem + (uint32_t)(p.a_shift_rows * rowb);                  // A region: first 40 KB
    const uint32_t b_addr = smem + 40960u + (uint32_t)(p.b_shift_rows * rowb);         // B region: next 80 KB (256 rows, groups 16 rows apart, + shift)
    const uint64_t da0 = make_desc(a_addr, p.swz_bytes, p.sbo_rows * rowb);
    const uint64_t db0 = make_desc(b_addr, p.swz_bytes, p.sbo_rows * rowb);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);

    if (warp == 0 && (CG == 1 || cta_rank == 0)) {
        uint32_t phase = 0;
        for (int burst = 0; burst < 3; ++burst) {       // burst 0 = warm-up
            const int R = burst == 2 ? p.r2 : p.r1;
            __syncwarp();
            long long t0 = 0, t1 = 0;
            if (elect_one()) {
                t0 = clock64();
                uint32_t acc = 0;
                for (int r = 0; r < R; ++r) {
                    const int k = r & (p.ksteps - 1);          // ksteps is 1, 2 or 4
                    const uint64_t da = da0 + (uint64_t)(k * 2), db = db0 + (uint64_t)(k * 2);
                    if (CG == 1)
                        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    acc = 1;
                }
                if (CG == 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                else
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
            }
            __syncwarp();
            mbar_wait(&bar, phase);
            phase ^= 1u;
            t1 = clock64();
            t0 = __shfl_sync(0xffffffffu, t0, __ffs(__ballot_sync(0xffffffffu, t0 != 0)) - 1);
            if (burst > 0 && (threadIdx.x & 31) == 0) p.out[(size_t)blockIdx.x * 2 + (burst - 1)] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
    }
}

static double run(int cg, Params p, int grid, long long* dout, const char* label)
{
    const size_t smem = 120 * 1024 + 1024;
    static bool attr = false;
    if (!attr) {
        CK(cudaFuncSetAttribute(mma_shapes_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(mma_shapes_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    p.out = dout;
    CK(cudaMemset(dout, 0, sizeof(long long) * 2 * grid));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = cg == 2 ? 1 : 0;
    if (cg == 1) CK(cudaLaunchKernelEx(&cfg, mma_shapes_kernel<1>, p));
    else CK(cudaLaunchKernelEx(&cfg, mma_shapes_kernel<2>, p));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(2 * grid);
    CK(cudaMemcpy(h.data(), dout, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
    double sum = 0, mx = 0;
    int n = 0;
    for (int b = 0; b < grid; ++b) {
        if (cg == 2 && (b & 1)) continue;          // only the leader CTA of a pair times
        const double c = (double)(h[2 * b + 1] - h[2 * b]) / (double)(p.r2 - p.r1);
        sum += c; if (c > mx) mx = c; ++n;
    }
    const double avg = sum / n;
    // FLOP per instruction = 2 * M * N * 16
    const double flop_per_clk = 2.0 * p.M * p.N * 16.0 / avg / (cg == 2 ? 2.0 : 1.0);      // per SM
    printf("{\"case\": \"%s\", \"cta_group\": %d, \"M\": %d, \"N\": %d, \"swizzle\": %d, \"sbo_rows\": %d, \"a_shift_rows\": %d, \"b_shift_rows\": %d, "
           "\"ksteps\": %d, \"cycles_per_mma\": %.1f, \"cycles_per_mma_max_sm\": %.1f, \"flop_per_clk_per_sm\": %.0f, \"pixels_per_clk_if_M_is_pixels\": %.3f, "
           "\"pixels_per_clk_if_N_is_pixels\": %.3f}\n",
           label, cg, p.M, p.N, p.swz_bytes, p.sbo_rows, p.a_shift_rows, p.b_shift_rows, p.ksteps, avg, mx, flop_per_clk,
           (double)p.M / avg / (cg == 2 ? 2.0 : 1.0), (double)p.N / avg / (cg == 2 ? 2.0 : 1.0));
    fflush(stdout);
    return avg;
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long* dout = nullptr;
    CK(cudaMalloc(&dout, sizeof(long long) * 2 * 1024));
    const int grid = sms & ~1;
    Params p = {};
    p.r1 = 128; p.r2 = 640;
    // 1. dense K-major tiles, 128B swizzle (Kc = 64: 4 K steps), M x N sweep, one SM per instruction
    for (int M : {128, 64})
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
            p.M = M; p.N = N; p.swz_bytes = 128; p.sbo_rows = 8; p.a_shift_rows = 0; p.b_shift_rows = 0; p.ksteps = 4;
            run(1, p, grid, dout, "dense_swz128");
        }
    // 2. narrower rows: 64B swizzle (Kc = 32, 2 K steps), 32B swizzle (Kc = 16, 1 K step)
    for (int swz : {64, 32})
        for (int N : {16, 32, 64, 128, 256}) {
            p.M = 128; p.N = N; p.swz_bytes = swz; p.sbo_rows = 8; p.a_shift_rows = 0; p.b_shift_rows = 0; p.ksteps = swz / 32;
            run(1, p, grid, dout, swz == 64 ? "dense_swz64" : "dense_swz32");
        }
    // 3. the conv engine's halo geometry: core groups 16 rows apart, A start shifted by the tap offset
    for (int swz : {128, 64})        // (32B swizzle with SBO = 16 rows faults: the engine packs such layers into 128-byte lines instead)
        for (int N : {32, 64})
            for (int shift : {0, 1, 17, 33}) {
                p.M = 128; p.N = N; p.swz_bytes = swz; p.sbo_rows = 16; p.a_shift_rows = shift; p.b_shift_rows = 0; p.ksteps = swz / 32;
                run(1, p, grid, dout, "halo_A_pixels");
            }
    // 4. operand roles swapped: weights on M (64 or 128 rows), pixels on N (halo geometry on B)
    for (int M : {64, 128})
        for (int N : {128, 256})
            for (int shift : {0, 1, 17}) {
                p.M = M; p.N = N; p.swz_bytes = 128; p.sbo_rows = 16; p.a_shift_rows = 0; p.b_shift_rows = shift; p.ksteps = 4;
                run(1, p, grid, dout, "halo_B_pixels");
            }
    // 5. CTA pairs (cta_group::2): M = 256 (128 rows of A per CTA), each CTA supplies N / 2 rows of B
    for (int N : {32, 64, 128, 256}) {
        p.M = 256; p.N = N; p.swz_bytes = 128; p.sbo_rows = 8; p.a_shift_rows = 0; p.b_shift_rows = 0; p.ksteps = 4;
        run(2, p, grid, dout, "pair_dense_swz128");
    }
    for (int N : {32, 64, 128, 256}) {
        p.M = 128; p.N = N; p.swz_bytes = 128; p.sbo_rows = 8; p.a_shift_rows = 0; p.b_shift_rows = 0; p.ksteps = 4;
        run(2, p, grid, dout, "pair_M128_swz128");
    }
    CK(cudaFree(dout));
    return 0;
}
