// umma_rate_probe.cu — issue-rate experiment for tcgen05.mma (kind::f16, BF16, M = 128, K = 16) with the operand
// layouts the kernels of this repo use: K-major (forward convolution, csrc/conv_tc.cu) and MN-major (weight gradient,
// csrc/wgrad_tc.cu), N = 64 / 128 / 256.  Operands are static in shared memory (no TMA): every CTA (one per SM) issues
// `iters` groups of four K steps walking over four 48-KB stages exactly as the kernels' main loops do, then commits and
// waits.  Prints clocks per MMA and the aggregate TFLOP/s.     nvcc -arch=sm_100a -O2 -o umma_rate_probe umma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(c), "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_k(uint32_t addr)      // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo)   // MN-major, SWIZZLE_128B, 64-element blocks lbo apart
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

constexpr int kStage = 48 * 1024, kStages = 4;

struct Variant { int a_mn, b_mn, n; };

__global__ void __launch_bounds__(128, 1) k_rate(Variant v, int iters, long long* clocks)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + kStages * kStage);
    uint32_t* holder = (uint32_t*)(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kStages * kStage / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3F803F80u;   // bf16 1.0
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *holder;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)v.a_mn << 15) | ((uint32_t)v.b_mn << 16) | ((uint32_t)(v.n >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        const long long t0 = clock64();
        int stage = 0;
        for (int it = 0; it < iters; it++) {
            const uint32_t sa = base + stage * kStage, sb = sa + 16 * 1024;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint64_t da = v.a_mn ? desc_mn(sa + k * 2048, 8192) : desc_k(sa) + (uint64_t)(k * 2);
                const uint64_t db = v.b_mn ? desc_mn(sb + k * 2048, 8192) : desc_k(sb) + (uint64_t)(k * 2);
                mma(tmem, da, db, idesc, (it | k) != 0);
            }
            if (++stage == kStages) stage = 0;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        mbar_wait(bar, 0);
        clocks[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main()
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int smem = kStages * kStage + 2048;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d_clk;
    cudaMalloc(&d_clk, sizeof(long long) * sms);
    long long* h = (long long*)malloc(sizeof(long long) * sms);
    const Variant vs[] = {{0, 0, 256}, {1, 1, 256}, {0, 0, 128}, {1, 1, 128}, {0, 0, 64}, {1, 1, 64}, {1, 0, 256}, {0, 1, 256}};
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (const Variant& v : vs) {
        k_rate<<<sms, 128, smem>>>(v, 200, d_clk);      // warm-up
        cudaEventRecord(e0);
        k_rate<<<sms, 128, smem>>>(v, iters, d_clk);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant a_mn=%d b_mn=%d n=%d: %s\n", v.a_mn, v.b_mn, v.n, cudaGetErrorString(e)); return 1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, d_clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1LL << 62;
        for (int i = 0; i < sms; i++) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
        const double nmma = 4.0 * iters, flop = nmma * 2.0 * 128 * v.n * 16 * sms;
        printf("A %s  B %s  N %3d : %7.1f clk/MMA (min %.1f)   ideal %5.1f   %7.1f TFLOP/s over %d SMs (%.3f ms)\n", v.a_mn ? "MN" : "K ", v.b_mn ? "MN" : "K ",
               v.n, mx / nmma, mn / nmma, 128.0 * v.n * 16 * 2 / 8192.0, flop / (ms * 1e-3) / 1e12, sms, ms);
    }
    return 0;
}
