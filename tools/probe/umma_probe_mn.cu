// umma_probe_mn.cu — experiment for the weight-gradient kernel: tcgen05.mma with BOTH operands MN-major (the reduction
// dimension = pixels is the ROW index of the NHWC activation tiles as TMA delivers them: rows of 64 channels = 128 B,
// SWIZZLE_128B).  D[m][n] = sum_p A[p][m] * B[p + shift][n], m < 128 (two 64-channel blocks, LBO apart), n < 64,
// p over 64 pixels (four K = 16 steps, +2048 B each).  Tries descriptor variants (LBO / SBO roles) and pixel shifts of
// the B operand (start address not 1024-B aligned along K).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(c), "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(b), "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kRows = 80;                 // pixels loaded per block
constexpr int kBlk = kRows * 128;         // bytes per 64-channel block (10 KB, 1024-B multiple)

__global__ void __launch_bounds__(128, 1)
k_probe_mn(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out /*[2 variants][9 shifts][128][64]*/)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                   // 2 blocks
    uint8_t* sB = smem + 2 * kBlk;        // 1 block
    uint64_t* bar = (uint64_t*)(smem + 3 * kBlk);
    uint64_t* mbar = bar + 1;
    uint32_t* holder = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *holder;
    if (threadIdx.x == 0) {
        mbar_expect(bar, 3 * kBlk);
        tma2d(sA, &mapA, bar, 0, 0);
        tma2d(sA + kBlk, &mapA, bar, 64, 0);
        tma2d(sB, &mapB, bar, 0, 0);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // c F32, a = b = BF16, a_major = b_major = MN (bits 15, 16), N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t par = 0;
    for (int var = 0; var < 1; var++)   // variant 1 (LBO / SBO swapped) strides 10 KB per 8 rows and reads past the shared-memory window: faults
        for (int s = 0; s < 9; s++) {
            if (threadIdx.x == 0) {
                // variant 0: LBO = distance between the 64-channel blocks, SBO = 1024 B (8 pixel rows); variant 1: swapped
                const uint32_t lbo = var == 0 ? kBlk : 1024, sbo = var == 0 ? 1024 : kBlk;
                for (int k = 0; k < 4; k++) {
                    uint64_t da = make_desc_mn(smem_u32(sA) + k * 2048, lbo, sbo);
                    uint64_t db = make_desc_mn(smem_u32(sB) + s * 128 + k * 2048, lbo, sbo);
                    uint32_t acc = k != 0;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
            }
            mbar_wait(mbar, par);
            par ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int half = 0; half < 2; half++) {
                uint32_t r[32];
                uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(ta) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* o = out + (((size_t)(var * 9 + s) * 128) + warp * 32 + lane) * 64 + half * 32;
                for (int i = 0; i < 32; i++) o[i] = __uint_as_float(r[i]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int umma_probe_mn(const void* A /*[80][128] bf16*/, const void* B /*[88][64] bf16*/, float* out)
{
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess) return -1;
    EncFn enc = (EncFn)fp;
    CUtensorMap ma, mb;
    cuuint32_t es[2] = {1, 1};
    cuuint64_t da[2] = {128, kRows}, sa[1] = {256}; cuuint32_t ba[2] = {64, kRows};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)A, da, sa, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -2;
    cuuint64_t db[2] = {64, kRows}, sb[1] = {128}; cuuint32_t bb[2] = {64, kRows};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)B, db, sb, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -3;
    cudaFuncSetAttribute(k_probe_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    k_probe_mn<<<1, 128, 3 * kBlk + 4096>>>(ma, mb, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe_mn: %s\n", cudaGetErrorString(e)); return -4; }
    return 0;
}
