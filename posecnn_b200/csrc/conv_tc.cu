// conv_tc.cu — the VGG16 convolution stack on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// Behavioural spec: Network.conv, lib/networks/network.py:159-188 (NHWC x HWIO, SAME padding,
// stride 1, bias add, optional ReLU) as wired by lib/networks/vgg16_convs.py:80-97, 128-163.
// The reference runs tf.nn.conv2d -> cuDNN in fp32; here the convolution is an implicit GEMM
//
//      D[m, n] = sum_{tap, c} A_tap[m, c] * Wt[n, tap*Cin + c],      m = pixel of an 8x16 tile
//
// with BF16 operands and FP32 accumulation (precision is stated with every number):
//   * im2col is never materialised: for every filter tap the A operand of a tile is ONE 4-D TMA
//     box {64 ch, 16 w, 8 h, 1 n} of the NHWC activation tensor at the tap's (dy, dx) offset;
//     out-of-image coordinates are zero-filled by the TMA unit = SAME padding;
//   * the box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, which is
//     exactly the canonical K-major UMMA operand layout, so tcgen05.mma reads it in place;
//   * accumulators live in TMEM (two stages of BN fp32 columns), the epilogue (bias, ReLU,
//     bf16 pack) runs out of TMEM while the next tile's MMAs are already being issued;
//   * warp roles: warps 0-3 epilogue (TMEM lanes 32w..32w+31), warp 4 TMA producer, warp 5
//     MMA issuer; persistent CTAs, one per SM, static round-robin tile schedule.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace pcnn {
namespace convtc {

constexpr int kTileH = 8, kTileW = 16;                            // 128 output pixels per tile (kTileM, tc_common.cuh)
static_assert(kTileH * kTileW == kTileM, "pixel tile = UMMA M");
constexpr int kABytes = kTileM * kKC * 2;                         // 16 KB
constexpr int kThreadsConv = 192;
constexpr int kStageBytes = 16 * 1024;                            // epilogue staging: one 64-channel group

struct ConvParams {
    int B, H, W, Cin, Cout;
    int taps, ksize;          // 9 / 3 or 1 / 1
    int tiles_h, tiles_w, n_tiles_n, total_tiles;
    int tile_h, tile_w;       // k_conv_tc: pixel tile (tile_h * tile_w = 128; 8 x 16, or 16 x 8 when that wastes fewer pixels)
    int relu;
    int pool;                 // 1: the epilogue applies the 2x2 / stride-2 max pool and stores ONLY the pooled tensor
    const float* bias;
};

template <int BN>
struct SmemPlan {
    static constexpr int kBBytes = BN * kKC * 2;
    static constexpr int kStage = kABytes + kBBytes;
    static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kOutBufs = 2;
    static constexpr int kBarOff = kStages * kStage + kOutBufs * kStageBytes;
    static constexpr int kTotal = kBarOff + 256 + 1024;  // barriers + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kThreadsConv, 1)
k_conv_tc(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w,
          const __grid_constant__ CUtensorMap map_out, const ConvParams p)  // map_out: box {64,16,8,1}, or {64,8,4,1} of the pooled tensor
{
    using Plan = SmemPlan<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* out_stage = smem + Plan::kStages * Plan::kStage;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Plan::kBarOff);
    uint64_t* empty = full + Plan::kStages;
    uint64_t* tfull = empty + Plan::kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = p.Cin / kKC;
    const int ksteps = p.taps * kchunks;
    constexpr uint32_t kTmemCols = 2 * BN >= 32 ? 2 * BN : 32;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_in) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    }
    if (warp == 5 && lane == 0) {
        for (int s = 0; s < Plan::kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_holder, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles_n;
                int rest = tile / p.n_tiles_n;
                const int tw = rest % p.tiles_w; rest /= p.tiles_w;
                const int th = rest % p.tiles_h;
                const int img = rest / p.tiles_h;
                const int h0 = th * p.tile_h, w0 = tw * p.tile_w, n0 = nt * BN;
                const int pad = p.ksize / 2;
                for (int ks = 0; ks < ksteps; ks++) {
                    const int tap = ks / kchunks, c0 = (ks % kchunks) * kKC;
                    const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Plan::kStage;
                    mbar_arrive_expect_tx(&full[stage], Plan::kStage);
                    tma_load_4d(sa, &map_in, &full[stage], c0, w0 + dx, h0 + dy, img);
                    tma_load_2d(sa + kABytes, &map_w, &full[stage], tap * p.Cin + c0, n0);
                    if (++stage == Plan::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = 0; ks < ksteps; ks++) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Plan::kStage);
                    const uint64_t da = make_desc(sa), db = make_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kKC / 16; k++)
                        umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks | k) != 0);
                    umma_commit(&empty[stage]);  // smem slot free once these MMAs have read it
                    if (++stage == Plan::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);  // accumulator complete
            }
        }
    } else {
        // ===================== epilogue warps 0..3 =====================
        int it = 0;
        int obuf = 0;
        const int row = warp * 32 + lane;  // pixel of the tile = TMEM lane
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
            const int nt = tile % p.n_tiles_n;
            int rest = tile / p.n_tiles_n;
            const int tw = rest % p.tiles_w; rest /= p.tiles_w;
            const int th = rest % p.tiles_h;
            const int img = rest / p.tiles_h;
            const int h0 = th * p.tile_h, w0 = tw * p.tile_w, n0 = nt * BN;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int g = 0; g < BN / 64; g++) {
                // staging buffer obuf must have been read out by the TMA store issued two groups ago
                if (threadIdx.x == 0) tma_store_wait_read<1>();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                uint8_t* ob = out_stage + obuf * kStageBytes;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + g * 64 + half * 32, r);
                    tmem_ld_wait();
                    const float* bias = p.bias + n0 + g * 64 + half * 32;
#pragma unroll
                    for (int j = 0; j < 4; j++) {  // four 16-byte pieces = 8 channels each
                        uint32_t packed[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            float v0 = __uint_as_float(r[j * 8 + q * 2]) + __ldg(bias + j * 8 + q * 2);
                            float v1 = __uint_as_float(r[j * 8 + q * 2 + 1]) + __ldg(bias + j * 8 + q * 2 + 1);
                            if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                            packed[q] = *reinterpret_cast<uint32_t*>(&b2);
                        }
                        const int piece = half * 4 + j;  // 16-byte piece index within the 128-byte row
                        if (!p.pool) {
                            uint4* dst = reinterpret_cast<uint4*>(ob + row * 128 + ((piece ^ (row & 7)) << 4));
                            *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        } else {
                            // fused max_pool 2x2/2 (network.py:303-310): tile row m = 16 h + w, a warp holds two image
                            // rows; the 2x2 window of (h even, w even) lives in lanes l, l^1, l^16, l^17.  max on the
                            // bf16 values = bf16 of the fp32 max (rounding is monotone).
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&packed[q]);
                                uint32_t o1 = __shfl_xor_sync(0xffffffffu, packed[q], 1);
                                v = __hmax2(v, *reinterpret_cast<__nv_bfloat162*>(&o1));
                                uint32_t vv = *reinterpret_cast<uint32_t*>(&v);
                                uint32_t o2 = __shfl_xor_sync(0xffffffffu, vv, 16);
                                v = __hmax2(v, *reinterpret_cast<__nv_bfloat162*>(&o2));
                                packed[q] = *reinterpret_cast<uint32_t*>(&v);
                            }
                            if ((lane & 17) == 0) {
                                const int prow = warp * 8 + (lane >> 1);  // pooled tile: 4 rows x 8 cols, row = ph * 8 + pw
                                uint4* dst = reinterpret_cast<uint4*>(ob + prow * 128 + ((piece ^ (prow & 7)) << 4));
                                *dst = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                            }
                        }
                    }
                }
                fence_proxy_async();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (threadIdx.x == 0) {
                    if (!p.pool) tma_store_4d(&map_out, ob, n0 + g * 64, w0, h0, img);
                    else tma_store_4d(&map_out, ob, n0 + g * 64, w0 >> 1, h0 >> 1, img);
                    tma_store_commit();
                }
                obuf ^= 1;
            }
            // all TMEM reads of this accumulator stage are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (threadIdx.x == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// Row mode for the K-small layers (conv1_2, conv2_x: Cin, Cout <= 128), which are bound by L2 -> SM operand
// traffic in the tile kernel (every tap re-fetches its A tile; profiles/r01_conv_tile_kernel_ncu_full_early.txt).
// A work item is TWO output rows x 128 pixels of one image:
//   * per 64-channel chunk ONE TMA box {64 ch, 130 px, 4 rows} (halo included) is loaded; the A operand of
//     tap (r, s) for output row j is the 128 consecutive patch rows starting at ((r + j) * 130 + s): a
//     row-shifted view of the same shared-memory patch (the UMMA swizzle is a function of the shared-memory
//     address, so a start address that is 128-B but not 1024-B aligned needs no descriptor change — checked on
//     hardware with tools/probe/umma_probe.cu);
//   * every weight stage feeds both output rows (two accumulators), halving the weight traffic as well;
//   * the two rows of a pair are exactly the rows a 2x2 max pool combines, so the pool stays fused: vertical
//     max in registers (both rows live in the same TMEM lane), horizontal max by shuffle.
// Operand bytes per 128 output pixels drop from 9 * (16 + BN/8) KB to (32.5 + 4.5 * BN/8) KB per chunk.
// ---------------------------------------------------------------------------------------------
constexpr int kRowPx = 128, kPatchW = 130, kPatchH = 4;
constexpr int kPatchBytes = kPatchW * kPatchH * 128;  // 66,560 = 65 * 1024
constexpr int kRowAStages = 2;

// RESB (Cin = 64, BN = 64 only): the nine 8-KB weight slices of the CTA's N tile stay resident in shared memory for
// the whole kernel (a 48-KB ring cannot keep enough weight bytes in flight to hide the L2 latency); every CTA keeps
// one N tile (nt = blockIdx.x % n_tiles_n).  The epilogue staging shrinks to 16 KB to make room.
template <int BN, bool RESB>
struct RowPlan {
    static constexpr int kBBytes = BN * 128;
    static constexpr int kBStages = RESB ? 9 : (BN == 64 ? 6 : 3);
    static constexpr int kBOff = kRowAStages * kPatchBytes;
    static constexpr int kOutOff = kBOff + kBStages * kBBytes;
    static constexpr int kOutBytes = RESB ? kStageBytes : 2 * kStageBytes;
    static constexpr int kBarOff = kOutOff + kOutBytes;
    static constexpr int kTotal = kBarOff + 256 + 1024;
};

template <int BN, bool RESB>
__global__ void __launch_bounds__(kThreadsConv, 1)
k_conv_row2(const __grid_constant__ CUtensorMap map_in /*box {64,130,4,1}*/, const __grid_constant__ CUtensorMap map_w,
            const __grid_constant__ CUtensorMap map_out /*box {64,128,1,1}, or {64,64,1,1} of the pooled tensor*/,
            const ConvParams p)
{
    using Plan = RowPlan<BN, RESB>;
    static_assert(!RESB || BN == 64, "resident weights: BN = 64 only");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem + Plan::kBOff;
    uint8_t* out_stage = smem + Plan::kOutOff;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + Plan::kBarOff);
    uint64_t* emptyA = fullA + kRowAStages;
    uint64_t* fullB = emptyA + kRowAStages;
    uint64_t* emptyB = fullB + Plan::kBStages;
    uint64_t* tfull = emptyB + Plan::kBStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wbar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = p.Cin / kKC;
    constexpr uint32_t kTmemCols = 4 * BN;
    // work items of this CTA: RESB -> fixed N tile, spatial tiles strided; otherwise all (spatial, N) tiles strided
    const int tile0 = RESB ? blockIdx.x / p.n_tiles_n : blockIdx.x;
    const int tstep = RESB ? gridDim.x / p.n_tiles_n : gridDim.x;
    const int tcount = RESB ? p.total_tiles / p.n_tiles_n : p.total_tiles;
    const int nt_fixed = blockIdx.x % p.n_tiles_n;  // 2 accumulator stages x 2 output rows
    // tiles_h = row pairs, tiles_w = 128-pixel segments
    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_in) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    }
    if (warp == 5 && lane == 0) {
        for (int s = 0; s < kRowAStages; s++) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
        for (int s = 0; s < Plan::kBStages; s++) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_holder, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            if (RESB) {
                mbar_arrive_expect_tx(wbar, 9 * Plan::kBBytes);
                // resident slices in (s, r) order: the slices of taps (r - 1, s) and (r, s) are adjacent and form ONE
                // 128-row B tile for the merged two-row MMA below
                for (int tap = 0; tap < 9; tap++)
                    tma_load_2d(sB + ((tap % 3) * 3 + tap / 3) * Plan::kBBytes, &map_w, wbar, tap * p.Cin, nt_fixed * BN);
            }
            for (int tile = tile0; tile < tcount; tile += tstep) {
                const int nt = RESB ? nt_fixed : tile % p.n_tiles_n;
                int rest = RESB ? tile : tile / p.n_tiles_n;
                const int tw = rest % p.tiles_w; rest /= p.tiles_w;
                const int yp = rest % p.tiles_h;
                const int img = rest / p.tiles_h;
                const int y0 = 2 * yp, x0 = tw * kRowPx, n0 = nt * BN;
                for (int c = 0; c < kchunks; c++) {
                    mbar_wait(&emptyA[sa], pa ^ 1);
                    mbar_arrive_expect_tx(&fullA[sa], kPatchBytes);
                    tma_load_4d(smem + sa * kPatchBytes, &map_in, &fullA[sa], c * kKC, x0 - 1, y0 - 1, img);
                    if (++sa == kRowAStages) { sa = 0; pa ^= 1; }
                    if (!RESB) {
                        for (int tap = 0; tap < 9; tap++) {
                            mbar_wait(&emptyB[sb], pb ^ 1);
                            mbar_arrive_expect_tx(&fullB[sb], Plan::kBBytes);
                            tma_load_2d(sB + sb * Plan::kBBytes, &map_w, &fullB[sb], tap * p.Cin + c * kKC, n0);
                            if (++sb == Plan::kBStages) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(BN);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int it = 0;
            if (RESB) mbar_wait(wbar, 0);
            for (int tile = tile0; tile < tcount; tile += tstep, it++) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                for (int c = 0; c < kchunks; c++) {
                    mbar_wait(&fullA[sa], pa);
                    const uint32_t a_base = smem_u32(smem + sa * kPatchBytes);
                    if (RESB) {
                        // The A operand of tap (r, s) for output row j is patch row rho = r + j: rows 0 and 1 of the pair
                        // read the SAME A tile for rho = 1, 2 (taps r = rho and r = rho - 1).  One N = 2 BN MMA against
                        // the two adjacent weight slices [W(rho-1, s) | W(rho, s)] writes [acc(row 1) | acc(row 0)]:
                        // A is fetched once instead of twice (with N = 64 the MMA is bound by the 128 B/clk shared-memory
                        // operand fetch, ncu: 65 clk per MMA for 33 clk of math).  rho = 1 goes first so that the first
                        // MMA initialises both accumulators.
                        tc_fence_after();
                        constexpr uint32_t idesc2 = make_idesc(2 * BN);
                        const uint32_t d_pair = tmem_base + (acc * 2) * BN;      // [row 1 | row 0]
#pragma unroll
                        for (int o = 0; o < 4; o++) {
                            const int rho = o == 0 ? 1 : (o == 1 ? 2 : (o == 2 ? 0 : 3));
#pragma unroll
                            for (int s3 = 0; s3 < 3; s3++) {
                                const uint64_t da = make_desc(a_base + (uint32_t)((rho * kPatchW + s3) * 128));
                                if (rho == 1 || rho == 2) {
                                    const uint64_t db = make_desc(smem_u32(sB + (s3 * 3 + rho - 1) * Plan::kBBytes));
#pragma unroll
                                    for (int k = 0; k < kKC / 16; k++)
                                        umma_bf16(d_pair, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (c | o | s3 | k) != 0);
                                } else {
                                    // rho = 0: row 0 only, tap r = 0;  rho = 3: row 1 only, tap r = 2
                                    const uint64_t db = make_desc(smem_u32(sB + (s3 * 3 + (rho == 0 ? 0 : 2)) * Plan::kBBytes));
                                    const uint32_t d1 = d_pair + (rho == 0 ? BN : 0);
#pragma unroll
                                    for (int k = 0; k < kKC / 16; k++) umma_bf16(d1, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1);
                                }
                            }
                        }
                    }
                    for (int tap = 0; tap < (RESB ? 0 : 9); tap++) {
                        if (!RESB) mbar_wait(&fullB[sb], pb);
                        tc_fence_after();
                        const int r = tap / 3, s = tap - 3 * r;
                        const uint64_t db = make_desc(smem_u32(sB + (RESB ? tap : sb) * Plan::kBBytes));
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const uint64_t da = make_desc(a_base + (uint32_t)(((r + j) * kPatchW + s) * 128));
                            const uint32_t d_tmem = tmem_base + (acc * 2 + j) * BN;
#pragma unroll
                            for (int k = 0; k < kKC / 16; k++)
                                umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (c | tap | k) != 0);
                        }
                        if (!RESB) {
                            umma_commit(&emptyB[sb]);
                            if (++sb == Plan::kBStages) { sb = 0; pb ^= 1; }
                        }
                    }
                    umma_commit(&emptyA[sa]);
                    if (++sa == kRowAStages) { sa = 0; pa ^= 1; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ===================== epilogue warps 0..3: thread = pixel x0 + row of both output rows =====================
        int it = 0, obuf = 0;
        const int row = warp * 32 + lane;
        for (int tile = tile0; tile < tcount; tile += tstep, it++) {
            const int nt = RESB ? nt_fixed : tile % p.n_tiles_n;
            int rest = RESB ? tile : tile / p.n_tiles_n;
            const int tw = rest % p.tiles_w; rest /= p.tiles_w;
            const int yp = rest % p.tiles_h;
            const int img = rest / p.tiles_h;
            const int y0 = 2 * yp, x0 = tw * kRowPx, n0 = nt * BN;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
            if (!p.pool) {
#pragma unroll 1
                for (int jg = 0; jg < 2 * (BN / 64); jg++) {
                    const int j = jg / (BN / 64), g = jg % (BN / 64);
                    if (threadIdx.x == 0) { if (RESB) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    uint8_t* ob = out_stage + (RESB ? 0 : obuf * kStageBytes);
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        uint32_t rr[32];
                        tmem_ld_32x32(t_lane + (acc * 2 + (RESB ? 1 - j : j)) * BN + g * 64 + half * 32, rr);   // RESB: [row 1 | row 0]
                        tmem_ld_wait();
                        const float* bias = p.bias + n0 + g * 64 + half * 32;
#pragma unroll
                        for (int q4 = 0; q4 < 4; q4++) {
                            uint32_t packed[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                float v0 = __uint_as_float(rr[q4 * 8 + q * 2]) + __ldg(bias + q4 * 8 + q * 2);
                                float v1 = __uint_as_float(rr[q4 * 8 + q * 2 + 1]) + __ldg(bias + q4 * 8 + q * 2 + 1);
                                if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                                packed[q] = *reinterpret_cast<uint32_t*>(&b2);
                            }
                            const int piece = half * 4 + q4;
                            *reinterpret_cast<uint4*>(ob + row * 128 + ((piece ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        }
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (threadIdx.x == 0) {
                        tma_store_4d(&map_out, ob, n0 + g * 64, x0, y0 + j, img);
                        tma_store_commit();
                    }
                    obuf ^= 1;
                }
            } else {
#pragma unroll 1
                for (int g = 0; g < BN / 64; g++) {
                    if (threadIdx.x == 0) tma_store_wait_read<1>();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    uint8_t* ob = out_stage + obuf * (RESB ? kStageBytes / 2 : kStageBytes);  // pooled tile: 64 rows = 8 KB
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        uint32_t r0[32], r1[32];
                        tmem_ld_32x32(t_lane + (acc * 2 + 0) * BN + g * 64 + half * 32, r0);
                        tmem_ld_32x32(t_lane + (acc * 2 + 1) * BN + g * 64 + half * 32, r1);
                        tmem_ld_wait();
                        const float* bias = p.bias + n0 + g * 64 + half * 32;
#pragma unroll
                        for (int q4 = 0; q4 < 4; q4++) {
                            uint32_t packed[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const float b0 = __ldg(bias + q4 * 8 + q * 2), b1 = __ldg(bias + q4 * 8 + q * 2 + 1);
                                // vertical max of the pair (same bias, monotone ReLU / rounding: order is irrelevant)
                                float v0 = fmaxf(__uint_as_float(r0[q4 * 8 + q * 2]), __uint_as_float(r1[q4 * 8 + q * 2])) + b0;
                                float v1 = fmaxf(__uint_as_float(r0[q4 * 8 + q * 2 + 1]), __uint_as_float(r1[q4 * 8 + q * 2 + 1])) + b1;
                                if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(v0, v1);
                                uint32_t pk = *reinterpret_cast<uint32_t*>(&b2);
                                uint32_t o1 = __shfl_xor_sync(0xffffffffu, pk, 1);  // horizontal neighbour x ^ 1
                                b2 = __hmax2(b2, *reinterpret_cast<__nv_bfloat162*>(&o1));
                                packed[q] = *reinterpret_cast<uint32_t*>(&b2);
                            }
                            if ((lane & 1) == 0) {
                                const int prow = row >> 1;  // pooled pixel within the 64-wide pooled segment
                                const int piece = half * 4 + q4;
                                *reinterpret_cast<uint4*>(ob + prow * 128 + ((piece ^ (prow & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                            }
                        }
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (threadIdx.x == 0) {
                        tma_store_4d(&map_out, ob, n0 + g * 64, x0 >> 1, y0 >> 1, img);
                        tma_store_commit();
                    }
                    obuf ^= 1;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (threadIdx.x == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// First layer (conv1_1, Cin = 3, K = 27 -> 32) on the tensor cores with the im2col done in shared memory:
// warps 8-23 (four groups, round-robin over tiles) build the A tile of an 8x16 pixel tile straight from the uint8 / f32
// image (pre-processing `BGR - PIXEL_MEANS` fused, zero outside the image = SAME padding) in the swizzled K-major
// layout, warp 24 issues two K = 16 UMMAs per tile against the resident 64 x 64 weight tile, warps 0-7 are two
// epilogue groups, one per TMEM accumulator.  No im2col tensor ever exists in HBM.
// The layer is a 1.26 GB write with almost no math: per tile the builder (27 dependent-free byte loads, ~1 us) and
// the epilogue (TMEM load -> pack -> TMA store) are latency chains, so the kernel keeps four builds and two
// epilogues in flight per SM (ncu on the 2-builder / 1-epilogue version: issue slots 44 % busy, 0.9 us per tile).
// ---------------------------------------------------------------------------------------------
constexpr int kC1Stages = 6;
constexpr int kC1EpiGroups = 2, kC1BuildGroups = 4, kC1Accs = 4;
constexpr int kC1Threads = 128 * (kC1EpiGroups + kC1BuildGroups) + 32;  // 8 epilogue warps, 16 A-builder warps, 1 MMA warp
constexpr int kC1MmaWarp = 4 * (kC1EpiGroups + kC1BuildGroups);
constexpr int kC1PatchWords = 10 * 16;   // raw uint8 input patch of a tile: 10 rows x 14 words (+2 pad), per builder group x 2 buffers
constexpr int kC1PatchOff = kC1Stages * kABytes + 64 * 128 + kC1EpiGroups * 2 * kStageBytes;
constexpr int kC1BarOff = kC1PatchOff + kC1BuildGroups * 2 * kC1PatchWords * 4;
constexpr int kC1Smem = kC1BarOff + 256 + 1024;

// Input element types: unsigned char / float = [B,H,W,3] colour image (BGR); DepthIn = [B,H,W] raw depth image (one
// float per pixel, sensor units): the `_p` trunk's input blob clip(d / 2000, 0, 1) * 255 tiled x3 - PIXEL_MEANS
// (lib/fcn/test.py:70-76) is formed on the fly in float32 exactly as numpy forms it.
struct DepthIn { float d; };

template <typename TIn>
__global__ void __launch_bounds__(kC1Threads, 1)
k_conv1_tc(const TIn* __restrict__ in, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
           const ConvParams p, float m0, float m1, float m2)
{
    constexpr int BN = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sb = smem + kC1Stages * kABytes;          // weights: 64 rows x 128 B, swizzled (TMA)
    uint8_t* out_stage = sb + 64 * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kC1BarOff);
    uint64_t* empty = full + kC1Stages;
    uint64_t* tfull = empty + kC1Stages;
    uint64_t* tempty = tfull + kC1Accs;
    uint64_t* wbar = tempty + kC1Accs;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wbar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = kC1Accs * BN;

    if (warp == kC1MmaWarp && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        for (int s = 0; s < kC1Stages; s++) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
        for (int a = 0; a < kC1Accs; a++) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == kC1MmaWarp) tmem_alloc(tmem_holder, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp >= 4 * kC1EpiGroups && warp < kC1MmaWarp) {
        // ===================== A builders: one tile row (pixel) per thread; the groups take tiles round-robin =====================
        const int group = (threadIdx.x - 128 * kC1EpiGroups) >> 7;
        const int r = (threadIdx.x - 128 * kC1EpiGroups) & 127;
        const int hl = r >> 4, wl = r & 15;
        uint32_t* patch_base = reinterpret_cast<uint32_t*>(smem + kC1PatchOff);
        // aligned-word staging needs word-aligned image rows
        const bool fast_u8 = sizeof(TIn) == 1 && (p.W & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0;
        int it = 0, lit = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
            if (it % kC1BuildGroups != group) continue;
            const int stage = it % kC1Stages;
            const uint32_t phase = (it / kC1Stages) & 1;
            const int tw = tile % p.tiles_w;
            const int rest = tile / p.tiles_w;
            const int th = rest % p.tiles_h, img = rest / p.tiles_h;
            const int y = th * kTileH + hl, x = tw * kTileW + wl;
            constexpr int kInCh = std::is_same<TIn, DepthIn>::value ? 1 : 3;
            const TIn* base = in + (size_t)img * p.H * p.W * kInCh;
            float v[32];
            // K = 27, 28 carry 1.0: the matching weight rows hold the bias (bf16 hi + lo parts, patched into the
            // resident weight tile by the MMA warp), so the bias add happens inside the MMA
            v[27] = 1.f; v[28] = 1.f;
#pragma unroll
            for (int k = 29; k < 32; k++) v[k] = 0.f;
            bool done = false;
            if constexpr (sizeof(TIn) == 1) {
                if (fast_u8) {
                    // uint8 fast path: the tile's raw 10 x 18 px patch is staged once in shared memory with aligned
                    // 32-bit loads -- a patch row starts at byte (16 tw - 1) * 3 = 48 tw - 3 of the image row, so the
                    // 14-word window from byte 48 tw - 4 is word aligned and the patch sits at byte offset 1 in it --
                    // and every thread then cuts its 3 x 9 bytes out of it (3 words + 2 byte-permutes per row) instead
                    // of issuing 27 byte loads with their own address arithmetic and bounds tests.
                    uint32_t* pw = patch_base + (group * 2 + (lit & 1)) * kC1PatchWords;
                    const int y0 = th * kTileH - 1, wq0 = 12 * tw - 1, row_words = (p.W * 3) >> 2;
                    const uint32_t* img32 = reinterpret_cast<const uint32_t*>(base);
                    for (int idx = r; idx < 140; idx += 128) {
                        const int prow = idx / 14, wi = idx - prow * 14;
                        const int yy = y0 + prow, wq = wq0 + wi;
                        uint32_t word = 0;
                        if (yy >= 0 && yy < p.H && wq >= 0 && wq < row_words) word = __ldg(img32 + (size_t)yy * row_words + wq);
                        pw[prow * 16 + wi] = word;
                    }
                    if (group == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
                    else if (group == 1) asm volatile("bar.sync 4, 128;" ::: "memory");
                    else if (group == 2) asm volatile("bar.sync 5, 128;" ::: "memory");
                    else asm volatile("bar.sync 6, 128;" ::: "memory");
                    const int boff = 1 + 3 * wl, w0 = boff >> 2, o = boff & 3;
                    const uint32_t sel = (uint32_t)(o | ((o + 1) << 4) | ((o + 2) << 8) | ((o + 3) << 12));
                    const bool border = th == 0 || th * kTileH + kTileH + 1 > p.H || tw == 0 || tw * kTileW + kTileW + 1 > p.W;
#pragma unroll
                    for (int dy = 0; dy < 3; dy++) {
                        const uint32_t* rowp = pw + (hl + dy) * 16 + w0;
                        const uint32_t a0 = rowp[0], a1 = rowp[1], a2 = rowp[2];
                        const uint32_t b03 = __byte_perm(a0, a1, sel), b47 = __byte_perm(a1, a2, sel), b8 = (a2 >> (8 * o)) & 0xffu;
#pragma unroll
                        for (int k = 0; k < 9; k++) {
                            const uint32_t byte = k < 4 ? (b03 >> (8 * k)) & 0xffu : (k < 8 ? (b47 >> (8 * (k - 4))) & 0xffu : b8);
                            const int c = k % 3;
                            v[dy * 9 + k] = (float)byte - (c == 0 ? m0 : (c == 1 ? m1 : m2));
                        }
                    }
                    if (border) {   // SAME padding pads the mean-subtracted image with zeros
#pragma unroll
                        for (int dy = 0; dy < 3; dy++) {
                            const int yy = y + dy - 1;
#pragma unroll
                            for (int dx = 0; dx < 3; dx++) {
                                const int xx = x + dx - 1;
                                if (!(yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)) {
                                    v[(dy * 3 + dx) * 3 + 0] = 0.f; v[(dy * 3 + dx) * 3 + 1] = 0.f; v[(dy * 3 + dx) * 3 + 2] = 0.f;
                                }
                            }
                        }
                    }
                    lit++;
                    done = true;
                }
            }
            if (!done) {
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    const int yy = y + dy - 1;
                    const bool rowok = yy >= 0 && yy < p.H;
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) {
                        const int xx = x + dx - 1;
                        const bool ok = rowok && xx >= 0 && xx < p.W;
                        if constexpr (std::is_same<TIn, DepthIn>::value) {
                            float g = 0.f;
                            if (ok) g = __fmul_rn(fminf(fmaxf(__fdiv_rn(base[(size_t)yy * p.W + xx].d, 2000.f), 0.f), 1.f), 255.f);
                            v[(dy * 3 + dx) * 3 + 0] = ok ? __fsub_rn(g, m0) : 0.f;
                            v[(dy * 3 + dx) * 3 + 1] = ok ? __fsub_rn(g, m1) : 0.f;
                            v[(dy * 3 + dx) * 3 + 2] = ok ? __fsub_rn(g, m2) : 0.f;
                        } else {
                            const TIn* px = base + ((size_t)yy * p.W + xx) * 3;
                            v[(dy * 3 + dx) * 3 + 0] = ok ? (float)px[0] - m0 : 0.f;
                            v[(dy * 3 + dx) * 3 + 1] = ok ? (float)px[1] - m1 : 0.f;
                            v[(dy * 3 + dx) * 3 + 2] = ok ? (float)px[2] - m2 : 0.f;
                        }
                    }
                }
            }
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kABytes + r * 128;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t pk[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[j * 8 + q * 2], v[j * 8 + q * 2 + 1]);
                    pk[q] = *reinterpret_cast<uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(sa + ((j ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
        }
    } else if (warp == kC1MmaWarp) {
        // ===================== MMA issuer (+ one-time weight load) =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(wbar, 64 * 128);
            tma_load_2d(sb, &map_w, wbar, 0, 0);
        }
        __syncwarp();
        mbar_wait(wbar, 0);   // every lane observes the completed TMA before it patches the tile
        // bias -> weight rows K = 27 (bf16 of the bias) and K = 28 (bf16 of the remainder), swizzled K-major layout:
        // element k of output channel n sits at n * 128 + ((k / 8) ^ (n & 7)) * 16 + (k % 8) * 2
        __syncwarp();
        for (int n = lane; n < BN; n += 32) {
            const float bv = p.bias[n];
            const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
            const __nv_bfloat16 lo = __float2bfloat16_rn(bv - __bfloat162float(hi));
            uint8_t* rowp = sb + n * 128 + ((3 ^ (n & 7)) << 4);
            *reinterpret_cast<__nv_bfloat16*>(rowp + 6) = hi;    // k = 27
            *reinterpret_cast<__nv_bfloat16*>(rowp + 8) = lo;    // k = 28
        }
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(BN);
            const uint64_t db = make_desc(smem_u32(sb));
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
                const int acc = it % kC1Accs;
                const uint32_t acc_phase = (it / kC1Accs) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint64_t da = make_desc(smem_u32(smem + stage * kABytes));
                umma_bf16(tmem_base + acc * BN, da, db, idesc, 0);            // K 0..15
                umma_bf16(tmem_base + acc * BN, da + 2, db + 2, idesc, 1);    // K 16..31 (27..31 are zero)
                // ONE commit per tile: the epilogue that observes tfull also releases the A stage (software arrive on
                // empty[stage]) -- a second tcgen05.commit per 2-MMA tile costs more than the tile's math
                umma_commit(&tfull[acc]);
                if (++stage == kC1Stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: group g = warps 4g..4g+3 owns TMEM accumulator g (tiles with it & 1 == g) =====================
        int it = 0, obuf = 0;
        const int eg = warp >> 2, wq = warp & 3;
        const int row = wq * 32 + lane;
        const bool issuer = (threadIdx.x & 127) == 0;
        uint8_t* my_stage = out_stage + eg * 2 * kStageBytes;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, it++) {
            if ((it & 1) != eg) continue;
            const int tw = tile % p.tiles_w;
            const int rest = tile / p.tiles_w;
            const int th = rest % p.tiles_h, img = rest / p.tiles_h;
            const int acc = it % kC1Accs;                 // four TMEM accumulators: the MMA warp runs up to four tiles ahead
            const uint32_t acc_phase = (it / kC1Accs) & 1;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            if (wq == 0 && lane == 0) mbar_arrive(&empty[it % kC1Stages]);   // the tile's MMAs are done reading its A stage
            const uint32_t t_addr = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BN;
            // both halves of the accumulator into registers, then hand the TMEM stage back before any math
            uint32_t rr[2][32];
            tmem_ld_32x32(t_addr, rr[0]);
            tmem_ld_32x32(t_addr + 32, rr[1]);
            if (issuer) tma_store_wait_read<1>();         // overlaps the TMEM read latency
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            uint8_t* ob = my_stage + obuf * kStageBytes;
            const __nv_bfloat162 floor2 = __floats2bfloat162_rn(p.relu ? 0.f : -INFINITY, p.relu ? 0.f : -INFINITY);
#pragma unroll
            for (int half = 0; half < 2; half++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    uint32_t packed[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        // bias already inside the accumulator; ReLU on the packed pair (rounding is monotone, 0 is exact)
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(rr[half][j * 8 + q * 2]), __uint_as_float(rr[half][j * 8 + q * 2 + 1]));
                        b2 = __hmax2(b2, floor2);
                        packed[q] = *reinterpret_cast<uint32_t*>(&b2);
                    }
                    const int piece = half * 4 + j;
                    *reinterpret_cast<uint4*>(ob + row * 128 + ((piece ^ (row & 7)) << 4)) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                }
            }
            fence_proxy_async();
            if (eg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if (issuer) {
                tma_store_4d(&map_out, ob, 0, tw * kTileW, th * kTileH, img);
                tma_store_commit();
            }
            obuf ^= 1;
        }
        if (issuer) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kC1MmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// conv1_1 (Cin = 3, K = 27: below any tensor-core tile) on the CUDA cores, fused with the input
// pre-processing-free path: fp32 NHWC in, bf16 NHWC out, bias + ReLU.  One thread = one pixel x 16
// output channels; the 27 x Cout weights sit in shared memory.
// ---------------------------------------------------------------------------------------------
template <int CO_PER_THREAD>
__global__ void __launch_bounds__(256)
k_conv_small_cin(const float* __restrict__ in, const float* __restrict__ w /*[3][3][Cin][Cout]*/,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int H, int W, int Cin, int Cout,
                 int relu)
{
    extern __shared__ float sw[];  // [9*Cin][Cout] + bias[Cout]
    const int K = 9 * Cin;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[K * Cout + i] = bias[i];
    __syncthreads();
    const int groups = Cout / CO_PER_THREAD;
    const size_t total = (size_t)B * H * W * groups;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(idx % groups);
        const size_t pix = idx / groups;
        const int x = (int)(pix % W), y = (int)((pix / W) % H);
        const size_t n = pix / ((size_t)W * H);
        float acc[CO_PER_THREAD];
#pragma unroll
        for (int k = 0; k < CO_PER_THREAD; k++) acc[k] = sw[K * Cout + g * CO_PER_THREAD + k];
        for (int r = 0; r < 3; r++) {
            const int yy = y + r - 1;
            if (yy < 0 || yy >= H) continue;
            for (int s = 0; s < 3; s++) {
                const int xx = x + s - 1;
                if (xx < 0 || xx >= W) continue;
                const float* ip = in + ((n * H + yy) * W + xx) * Cin;
                for (int c = 0; c < Cin; c++) {
                    const float v = __ldg(ip + c);
                    const float* wp = sw + ((r * 3 + s) * Cin + c) * Cout + g * CO_PER_THREAD;
#pragma unroll
                    for (int k = 0; k < CO_PER_THREAD; k++) acc[k] = fmaf(v, wp[k], acc[k]);
                }
            }
        }
        __nv_bfloat16* op = out + pix * Cout + g * CO_PER_THREAD;
#pragma unroll
        for (int k = 0; k < CO_PER_THREAD; k += 2) {
            float v0 = acc[k], v1 = acc[k + 1];
            if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            *reinterpret_cast<__nv_bfloat162*>(op + k) = __floats2bfloat162_rn(v0, v1);
        }
    }
}

// im2col for the first layer only (Cin = 3, K = 27 < one UMMA K step): [B,H,W,3] -> [B,H,W,64] bf16 with
// K index = tap * 3 + c (zero beyond 27), so conv1_1 runs on the tensor cores as a 1x1 convolution.
// The input pre-processing of the caller (lib/fcn/test.py:37-110: BGR - PIXEL_MEANS) is fused for uint8
// input: value = (float)u8 - mean[c].
template <typename TIn>
__global__ void __launch_bounds__(256)
k_im2col_c3(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out, int H, int W, float m0, float m1, float m2)
{
    // grid = (ceil(W / 128), H, B): a CTA builds 128 pixels x 64 K-values of one image row in four 32-pixel passes (614 k tiny
    // CTAs at batch 64 were launch-bound).  The 3 x 130 x 3 input patch (mean subtracted, zero outside the image = SAME padding)
    // is staged in shared memory once; thread (x, j) then packs the 8 K-values k = 8j .. 8j+7 (K order tap*3 + c) into one 16-byte store.
    constexpr int kSeg = 128;
    __shared__ float patch[3][(kSeg + 2) * 3];
    const int y = blockIdx.y, n = blockIdx.z, x0 = blockIdx.x * kSeg, t = threadIdx.x;
    const TIn* img = in + (size_t)n * H * W * 3;
    constexpr int kRow = (kSeg + 2) * 3;
    for (int i = t; i < 3 * kRow; i += 256) {
        const int r = i / kRow, rem = i - r * kRow, px = rem / 3, c = rem - px * 3;
        const int yy = y + r - 1, xx = x0 + px - 1;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = (float)img[(yy * W + xx) * 3 + c] - (c == 0 ? m0 : (c == 1 ? m1 : m2));
        patch[r][rem] = v;
    }
    __syncthreads();
    const int j = t & 7;
#pragma unroll
    for (int pass = 0; pass < kSeg / 32; pass++) {
        const int xl = pass * 32 + (t >> 3), x = x0 + xl;
        if (x >= W) break;
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int k = j * 8 + e;  // tap = k / 3 (dy = tap / 3, dx = tap % 3), c = k % 3 -> patch[dy][(xl + dx) * 3 + c]
            float val = 0.f;
            if (k < 27) {
                const int dy = k / 9, rem = k - dy * 9;  // rem = dx * 3 + c
                val = patch[dy][xl * 3 + rem];
            }
            v[e] = __float2bfloat16_rn(val);
        }
        *reinterpret_cast<uint4*>(out + (((size_t)n * H + y) * W + x) * 64 + j * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

// 2x2 / stride 2 max pool, NHWC bf16 (Network.max_pool, network.py:303-310; H, W even here)
__global__ void __launch_bounds__(256)
k_maxpool2x2_bf16(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H, int W, int C)
{
    const int Ho = H / 2, Wo = W / 2, cg = C / 8;
    const size_t total = (size_t)B * Ho * Wo * cg;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int g = (int)(idx % cg);
        size_t r = idx / cg;
        const int xo = (int)(r % Wo); r /= Wo;
        const int yo = (int)(r % Ho);
        const size_t n = r / Ho;
        const __nv_bfloat16* p0 = in + ((n * H + 2 * yo) * W + 2 * xo) * C + g * 8;
        uint4 a = __ldg(reinterpret_cast<const uint4*>(p0));
        uint4 b = __ldg(reinterpret_cast<const uint4*>(p0 + C));
        uint4 c = __ldg(reinterpret_cast<const uint4*>(p0 + (size_t)W * C));
        uint4 d = __ldg(reinterpret_cast<const uint4*>(p0 + (size_t)W * C + C));
        uint4 o;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
        const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&c);
        const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&d);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; k++) po[k] = __hmax2(__hmax2(pa[k], pb[k]), __hmax2(pc[k], pd[k]));
        *reinterpret_cast<uint4*>(out + ((n * Ho + yo) * Wo + xo) * C + g * 8) = o;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int make_map_nhwc(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int box_c, int box_w = kTileW, int box_h = kTileH)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return PCNN_E_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(NHWC %dx%dx%dx%d) failed: %d", B, H, W, C, (int)r); return PCNN_E_CUDA; }
    return PCNN_OK;
}

template <int BN>
static int launch_conv(const CUtensorMap& mi, const CUtensorMap& mw, const CUtensorMap& mo, const ConvParams& p, int num_sms,
                       cudaStream_t st)
{
    using Plan = SmemPlan<BN>;
    PCNN_SMEM_OPTIN(k_conv_tc<BN>, Plan::kTotal, "conv_tc");
    int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    k_conv_tc<BN><<<grid, kThreadsConv, Plan::kTotal, st>>>(mi, mw, mo, p);
    return check_launch("conv_tc");
}

template <int BN, bool RESB>
static int launch_conv_row2(const CUtensorMap& mi, const CUtensorMap& mw, const CUtensorMap& mo, const ConvParams& p, int num_sms,
                            cudaStream_t st)
{
    using Plan = RowPlan<BN, RESB>;
    PCNN_SMEM_OPTIN((k_conv_row2<BN, RESB>), Plan::kTotal, "conv_row2");
    int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    if (RESB) grid = grid / p.n_tiles_n * p.n_tiles_n;  // every CTA owns one N tile
    k_conv_row2<BN, RESB><<<grid, kThreadsConv, Plan::kTotal, st>>>(mi, mw, mo, p);
    return check_launch("conv_row2");
}

}  // namespace convtc
}  // namespace pcnn

using namespace pcnn;
using namespace pcnn::convtc;

// in [B,H,W,Cin] bf16, weights [Cout][ksize*ksize*Cin] bf16 (tap-major, channel-minor), bias [Cout] f32,
// out [B,H,W,Cout] bf16.  Cin % 64 == 0, Cout % 64 == 0, ksize in {1, 3}.
static int conv_bf16_tc_impl(const void* in, const void* weights, const float* bias, void* out, int B, int H, int W, int Cin,
                            int Cout, int ksize, int relu, int block_n, int pool, void* stream);

extern "C" int pcnn_conv_bf16_tc(const void* in, const void* weights, const float* bias, void* out, int B, int H, int W,
                                 int Cin, int Cout, int ksize, int relu, int block_n, void* stream)
{
    return conv_bf16_tc_impl(in, weights, bias, out, B, H, W, Cin, Cout, ksize, relu, block_n, 0, stream);
}

// conv + bias + ReLU + 2x2/2 max pool fused: out is the POOLED tensor [B,H/2,W/2,Cout] bf16 (H, W even)
extern "C" int pcnn_conv_pool_bf16_tc(const void* in, const void* weights, const float* bias, void* out_pooled, int B, int H,
                                      int W, int Cin, int Cout, int ksize, int relu, int block_n, void* stream)
{
    PCNN_REQUIRE(H % 2 == 0 && W % 2 == 0, "conv_pool: needs even H, W (got %d x %d)", H, W);
    return conv_bf16_tc_impl(in, weights, bias, out_pooled, B, H, W, Cin, Cout, ksize, relu, block_n, 1, stream);
}

static int conv_bf16_tc_impl(const void* in, const void* weights, const float* bias, void* out, int B, int H, int W, int Cin,
                            int Cout, int ksize, int relu, int block_n, int pool, void* stream)
{
    PCNN_REQUIRE(in && weights && bias && out, "conv: NULL tensor pointer");
    PCNN_REQUIRE(ksize == 1 || ksize == 3, "conv: ksize must be 1 or 3 (got %d)", ksize);
    PCNN_REQUIRE(Cin % 64 == 0 && Cin >= 64, "conv: Cin must be a multiple of 64 (got %d)", Cin);
    PCNN_REQUIRE(Cout % 64 == 0 && Cout >= 64, "conv: Cout must be a multiple of 64 (got %d)", Cout);
    PCNN_REQUIRE(B >= 1 && H >= 1 && W >= 1, "conv: bad shape");
    int bn = block_n;
    if (bn == 0) bn = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    PCNN_REQUIRE((bn == 64 || bn == 128 || bn == 256) && Cout % bn == 0, "conv: block_n %d does not divide Cout %d", bn, Cout);
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap mi, mw, mo;
    int rc;
    // row mode (two output rows x 128 px per work item, A patch reused by all nine taps) for the K-small layers
    static const bool row_mode_on = getenv("PCNN_CONV_ROWMODE") == nullptr || atoi(getenv("PCNN_CONV_ROWMODE")) != 0;
    if (row_mode_on && block_n == 0 && ksize == 3 && Cin <= 128 && Cout <= 128 && H % 2 == 0 && W >= kRowPx) {
        // conv1_2 shape (Cin = Cout = 64): the nine 8-KB weight slices stay resident in shared memory (measured:
        // 0.70 -> 0.65 ms); with two N tiles (conv2_1) the doubled A traffic costs more than the weight ring saves
        const bool resb = Cin == 64 && Cout == 64;
        bn = Cout;  // 64 or 128
        rc = make_map_nhwc(&mi, in, B, H, W, Cin, kKC, kPatchW, kPatchH);
        if (rc) return rc;
        rc = make_map_weights(&mw, weights, 9 * Cin, Cout, bn);
        if (rc) return rc;
        rc = pool ? make_map_nhwc(&mo, out, B, H / 2, W / 2, Cout, 64, kRowPx / 2, 1) : make_map_nhwc(&mo, out, B, H, W, Cout, 64, kRowPx, 1);
        if (rc) return rc;
        ConvParams p;
        p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = 3; p.taps = 9;
        p.tile_h = 2; p.tile_w = kRowPx;
        p.tiles_h = H / 2;
        p.tiles_w = (W + kRowPx - 1) / kRowPx;
        p.n_tiles_n = Cout / bn;
        p.total_tiles = B * p.tiles_h * p.tiles_w * p.n_tiles_n;
        p.relu = relu; p.pool = pool; p.bias = bias;
        if (resb && p.total_tiles >= p.n_tiles_n) return launch_conv_row2<64, true>(mi, mw, mo, p, sms, st);
        return bn == 128 ? launch_conv_row2<128, false>(mi, mw, mo, p, sms, st) : launch_conv_row2<64, false>(mi, mw, mo, p, sms, st);
    }
    // pixel tile: 8 x 16, or 16 x 8 when that covers the map with fewer tiles (conv5: 30 x 40 -> 10 tiles instead of 12).
    // The tile's pixel order is whatever the TMA box says (row = h * tile_w + w for load and store alike), so only
    // the box shape and the tile origin change; the fused pool epilogue is written for 8 x 16.
    int tile_h = kTileH, tile_w = kTileW;
    if (!pool && ((H + 15) / 16) * ((W + 7) / 8) < ((H + 7) / 8) * ((W + 15) / 16)) { tile_h = 16; tile_w = 8; }
    rc = make_map_nhwc(&mi, in, B, H, W, Cin, kKC, tile_w, tile_h);
    if (rc) return rc;
    rc = make_map_weights(&mw, weights, ksize * ksize * Cin, Cout, bn);
    if (rc) return rc;
    rc = pool ? make_map_nhwc(&mo, out, B, H / 2, W / 2, Cout, 64, kTileW / 2, kTileH / 2)
              : make_map_nhwc(&mo, out, B, H, W, Cout, 64, tile_w, tile_h);
    if (rc) return rc;
    ConvParams p;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.ksize = ksize; p.taps = ksize * ksize;
    p.tile_h = tile_h; p.tile_w = tile_w;
    p.tiles_h = (H + tile_h - 1) / tile_h;
    p.tiles_w = (W + tile_w - 1) / tile_w;
    p.n_tiles_n = Cout / bn;
    p.total_tiles = B * p.tiles_h * p.tiles_w * p.n_tiles_n;
    p.relu = relu;
    p.pool = pool;
    p.bias = bias;
    if (bn == 256) return launch_conv<256>(mi, mw, mo, p, sms, st);
    if (bn == 128) return launch_conv<128>(mi, mw, mo, p, sms, st);
    return launch_conv<64>(mi, mw, mo, p, sms, st);
}

// conv with tiny Cin (conv1_1): in [B,H,W,Cin] f32, weights HWIO [3,3,Cin,Cout] f32, out [B,H,W,Cout] bf16
extern "C" int pcnn_conv3x3_small_cin(const float* in, const float* weights_hwio, const float* bias, void* out, int B, int H,
                                      int W, int Cin, int Cout, int relu, void* stream)
{
    PCNN_REQUIRE(in && weights_hwio && bias && out, "conv_small: NULL tensor pointer");
    PCNN_REQUIRE(Cin >= 1 && Cin <= 8 && Cout % 16 == 0 && Cout <= 128, "conv_small: needs Cin <= 8, Cout %% 16 == 0, Cout <= 128");
    size_t smem = sizeof(float) * (size_t)(9 * Cin + 1) * Cout;
    size_t total = (size_t)B * H * W * (Cout / 16);
    int blocks = (int)((total + 255) / 256 < (size_t)kNumSMs * 8 ? (total + 255) / 256 : (size_t)kNumSMs * 8);
    k_conv_small_cin<16><<<blocks, 256, smem, (cudaStream_t)stream>>>(in, weights_hwio, bias, (__nv_bfloat16*)out, B, H, W, Cin,
                                                                      Cout, relu);
    return check_launch("conv_small_cin");
}

extern "C" int pcnn_maxpool2x2_bf16(const void* in, void* out, int B, int H, int W, int C, void* stream)
{
    PCNN_REQUIRE(in && out, "maxpool: NULL tensor pointer");
    PCNN_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool: needs even H, W and C %% 8 == 0 (got %d,%d,%d)", H, W, C);
    size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    int blocks = (int)((total + 255) / 256 < (size_t)kNumSMs * 16 ? (total + 255) / 256 : (size_t)kNumSMs * 16);
    k_maxpool2x2_bf16<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, H, W, C);
    return check_launch("maxpool2x2");
}

// first-layer im2col: in [B,H,W,3] (f32, or u8 with the per-channel mean subtracted) -> out [B,H,W,64] bf16
extern "C" int pcnn_im2col_c3(const void* in, int in_is_u8, const float* mean3_host, void* out_bf16, int B, int H, int W,
                              void* stream)
{
    PCNN_REQUIRE(in && out_bf16, "im2col: NULL tensor pointer");
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    if (mean3_host) { m0 = mean3_host[0]; m1 = mean3_host[1]; m2 = mean3_host[2]; }
    PCNN_REQUIRE(H <= 65535 && B <= 65535, "im2col: image too tall for the launch grid");
    dim3 grid((W + 127) / 128, H, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_is_u8)
        k_im2col_c3<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)in, (__nv_bfloat16*)out_bf16, H, W, m0, m1, m2);
    else
        k_im2col_c3<float><<<grid, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out_bf16, H, W, m0, m1, m2);
    return check_launch("im2col_c3");
}

// conv1_1 fused: in [B,H,W,3] (u8 minus mean, or f32), weights [64][64] bf16 in im2col K order (tap*3 + c, zero padded),
// bias [64] f32 -> out [B,H,W,64] bf16.  Replaces pcnn_im2col_c3 + a 1x1 pcnn_conv_bf16_tc without the HBM round trip.
static int conv1_fused_impl(const void* in, int in_kind, const float* mean3_host, const void* weights_bf16,
                            const float* bias, void* out_bf16, int B, int H, int W, int relu, void* stream);

extern "C" int pcnn_conv1_fused_tc(const void* in, int in_is_u8, const float* mean3_host, const void* weights_bf16,
                                   const float* bias, void* out_bf16, int B, int H, int W, int relu, void* stream)
{
    return conv1_fused_impl(in, in_is_u8 ? 1 : 0, mean3_host, weights_bf16, bias, out_bf16, B, H, W, relu, stream);
}

// conv1_1_p on a raw depth image [B,H,W] f32 (sensor units): the depth blob of lib/fcn/test.py:70-76 fused into the loader
extern "C" int pcnn_conv1_depth_fused_tc(const float* depth, const float* mean3_host, const void* weights_bf16,
                                         const float* bias, void* out_bf16, int B, int H, int W, int relu, void* stream)
{
    return conv1_fused_impl(depth, 2, mean3_host, weights_bf16, bias, out_bf16, B, H, W, relu, stream);
}

static int conv1_fused_impl(const void* in, int in_kind, const float* mean3_host, const void* weights_bf16,
                                   const float* bias, void* out_bf16, int B, int H, int W, int relu, void* stream)
{
    PCNN_REQUIRE(in && weights_bf16 && bias && out_bf16, "conv1: NULL tensor pointer");
    PCNN_REQUIRE(B >= 1 && H >= 1 && W >= 1, "conv1: bad shape");
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    if (mean3_host) { m0 = mean3_host[0]; m1 = mean3_host[1]; m2 = mean3_host[2]; }
    CUtensorMap mw, mo;
    int rc = make_map_weights(&mw, weights_bf16, 64, 64, 64);
    if (rc) return rc;
    rc = make_map_nhwc(&mo, out_bf16, B, H, W, 64, 64);
    if (rc) return rc;
    ConvParams p;
    p.B = B; p.H = H; p.W = W; p.Cin = 3; p.Cout = 64; p.ksize = 3; p.taps = 9;
    p.tile_h = kTileH; p.tile_w = kTileW;
    p.tiles_h = (H + kTileH - 1) / kTileH;
    p.tiles_w = (W + kTileW - 1) / kTileW;
    p.n_tiles_n = 1;
    p.total_tiles = B * p.tiles_h * p.tiles_w;
    p.relu = relu; p.pool = 0; p.bias = bias;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.total_tiles < sms ? p.total_tiles : sms;
    cudaStream_t st = (cudaStream_t)stream;
    PCNN_SMEM_OPTIN(k_conv1_tc<unsigned char>, kC1Smem, "conv1_tc<u8>");
    PCNN_SMEM_OPTIN(k_conv1_tc<float>, kC1Smem, "conv1_tc<f32>");
    PCNN_SMEM_OPTIN(k_conv1_tc<DepthIn>, kC1Smem, "conv1_tc<depth>");
    if (in_kind == 1) k_conv1_tc<unsigned char><<<grid, kC1Threads, kC1Smem, st>>>((const unsigned char*)in, mw, mo, p, m0, m1, m2);
    else if (in_kind == 2) k_conv1_tc<DepthIn><<<grid, kC1Threads, kC1Smem, st>>>((const DepthIn*)in, mw, mo, p, m0, m1, m2);
    else k_conv1_tc<float><<<grid, kC1Threads, kC1Smem, st>>>((const float*)in, mw, mo, p, m0, m1, m2);
    return check_launch("conv1_fused_tc");
}
