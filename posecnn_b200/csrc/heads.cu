// heads.cu — the FCN heads of vgg16_convs after the 1x1 convolutions on conv4_3 / conv5_3:
// semantic-label head and vertex head (lib/networks/vgg16_convs.py:128-163).
//
// The reference runs, at FULL resolution, two dense conv2d_transpose ops whose filters are fixed
// diagonal bilinear kernels (make_deconv_filter, lib/networks/network.py:141-157, 207-222) followed by
// 1x1 convolutions (`score` 64->C with ReLU, `vertex_pred` 128->3C): 50 GFLOP/frame of multiplications
// by zero plus 6 GFLOP of 1x1 work on 307 200 pixels.  Both stages are linear and act on different
// axes (bilinear: per channel over space; 1x1: per pixel over channels), so they commute:
//
//      conv1x1(up8(x)) + b  ==  up8(conv1x1_nobias(x)) + b
//
// (the bias must stay outside: near the image border the transposed convolution's weights do not sum
// to one).  Hence:
//   k_lowres_heads  at 1/8 resolution: add_score = score_conv4 + up2(score_conv5) (both heads), then
//                   the two 1x1 matrices -> [B,h,w, C + 3C] fp32
//   k_up8_heads     per output pixel: bilinear x8 (exact conv2d_transpose weights), + bias, ReLU /
//                   softmax / arg-max for the label head, + bias for the vertex head; one streaming
//                   pass that writes label_2d (int32), vertex_pred (fp32) and optionally the
//                   normalised probabilities.
#include <cuda_bf16.h>
#include <float.h>

#include <algorithm>

#include "common.cuh"

namespace pcnn {

// make_deconv_filter (network.py:141-157): f = ceil(k/2), c = (2f - 1 - f%2) / (2f), W[x] = 1 - |x/f - c|
__host__ __device__ inline float deconv_w(int x, int k)
{
    // k = 16: f = 8, c = 15/16; k = 4: f = 2, c = 3/4 (exact in binary floating point)
    if (k == 16) return 1.f - fabsf((float)x * 0.125f - 0.9375f);
    if (k == 4) return 1.f - fabsf((float)x * 0.5f - 0.75f);
    int f = (k + 1) / 2;
    float c = (2.f * f - 1.f - (float)(f % 2)) / (2.f * f);
    return 1.f - fabsf((float)x / (float)f - c);
}

// ---------------------------------------------------------------------------------------------
// k_lowres_heads: one warp per pair of low-resolution pixels.  Lanes first build the two "add" vectors
// (conv4 branch + 4x4/2 transposed convolution of the conv5 branch, coalesced bf16 reads) in the warp's
// shared-memory slot, then every lane owns up to three of the 4C outputs and runs the two small matrix
// products out of shared memory (weights conflict-free over lanes, activations broadcast).
// ---------------------------------------------------------------------------------------------
constexpr int kLrWarps = 8;

__global__ void __launch_bounds__(256)
k_lowres_heads(const __nv_bfloat16* __restrict__ s4 /*[B,h,w,Cs]*/, const __nv_bfloat16* __restrict__ s5 /*[B,h/2,w/2,Cs]*/,
               const __nv_bfloat16* __restrict__ v4 /*[B,h,w,Cv]*/, const __nv_bfloat16* __restrict__ v5,
               const float* __restrict__ Ws /*[Cs][C]*/, const float* __restrict__ Wv /*[Cv][3C]*/, int B, int h, int w,
               int Cs, int Cv, int C, float* __restrict__ out /*[B,h,w,4C]*/)
{
    extern __shared__ float sm[];
    const int Ct = Cs + Cv;        // channels of the two "add" tensors
    const int No = 4 * C;          // outputs per pixel
    float* sW = sm;                // [Cs*C + Cv*3C]
    float* sx_all = sm + Cs * C + Cv * 3 * C;  // [kLrWarps][2][Ct]
    for (int i = threadIdx.x; i < Cs * C; i += blockDim.x) sW[i] = Ws[i];
    for (int i = threadIdx.x; i < Cv * 3 * C; i += blockDim.x) sW[Cs * C + i] = Wv[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sx = sx_all + warp * 2 * Ct;
    const int npix = B * h * w;
    const int h5 = h / 2, w5 = w / 2;
    const int npairs = (npix + 1) / 2;
    for (int pr = blockIdx.x * kLrWarps + warp; pr < npairs; pr += gridDim.x * kLrWarps) {
        // ---- add = conv4 branch + up2(conv5 branch): conv2d_transpose 4x4 / stride 2, SAME (pad 1)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int p = 2 * pr + q;
            if (p >= npix) { for (int ch = lane; ch < Ct; ch += 32) sx[q * Ct + ch] = 0.f; continue; }
            const int x = p % w, y = (p / w) % h, n = p / (w * h);
            // out[o] += in[i] * W[o - 2 i + 1], 0 <= o - 2i + 1 <= 3: two source rows / columns
            const int iy0 = ((y + 1) >> 1) - 1, ix0 = ((x + 1) >> 1) - 1;
            float wgt[4];
            size_t off5[4];
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const int iy = iy0 + (d >> 1), ix = ix0 + (d & 1);
                const int ky = y - 2 * iy + 1, kx = x - 2 * ix + 1;
                const bool ok = iy >= 0 && iy < h5 && ix >= 0 && ix < w5;
                wgt[d] = ok ? deconv_w(ky, 4) * deconv_w(kx, 4) : 0.f;
                off5[d] = ((size_t)(n * h5 + min(max(iy, 0), h5 - 1)) * w5 + min(max(ix, 0), w5 - 1));
            }
            for (int ch = lane; ch < Ct; ch += 32) {
                const bool vert = ch >= Cs;
                const int cc = vert ? ch - Cs : ch, Cn = vert ? Cv : Cs;
                const __nv_bfloat16* a = vert ? v4 : s4;
                const __nv_bfloat16* b5 = vert ? v5 : s5;
                float up = wgt[0] * __bfloat162float(b5[off5[0] * Cn + cc]);
                up = fmaf(wgt[1], __bfloat162float(b5[off5[1] * Cn + cc]), up);
                up = fmaf(wgt[2], __bfloat162float(b5[off5[2] * Cn + cc]), up);
                up = fmaf(wgt[3], __bfloat162float(b5[off5[3] * Cn + cc]), up);
                sx[q * Ct + ch] = __bfloat162float(a[(size_t)p * Cn + cc]) + up;
            }
        }
        __syncwarp();
        // ---- outputs o = lane, lane + 32, lane + 64, ... for both pixels
        for (int o = lane; o < No; o += 32) {
            float a0 = 0.f, a1 = 0.f;
            if (o < C) {
                const float* wp = sW + o;
                for (int k = 0; k < Cs; k++) {
                    const float wk = wp[k * C];
                    a0 = fmaf(sx[k], wk, a0);
                    a1 = fmaf(sx[Ct + k], wk, a1);
                }
            } else {
                const float* wp = sW + Cs * C + (o - C);
                const int ld = 3 * C;
                for (int k = 0; k < Cv; k++) {
                    const float wk = wp[k * ld];
                    a0 = fmaf(sx[Cs + k], wk, a0);
                    a1 = fmaf(sx[Ct + Cs + k], wk, a1);
                }
            }
            const int p0 = 2 * pr;
            out[(size_t)p0 * No + o] = a0;
            if (p0 + 1 < npix) out[(size_t)(p0 + 1) * No + o] = a1;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// k_up8_heads: one CTA per output row (y, image).  The two contributing low-resolution rows are
// combined vertically into shared memory once; every output value is then a 2-tap horizontal blend.
// Class scores of the row are kept in shared memory for the per-pixel arg-max / softmax.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_up8_heads(const float* __restrict__ lr /*[B,h,w,4C]*/, const float* __restrict__ bias_s /*[C]*/,
            const float* __restrict__ bias_v /*[3C]*/, int h, int w, int C, int seg_cells, int* __restrict__ label /*[B,8h,8w]*/,
            float* __restrict__ vertex /*[B,8h,8w,3C]*/, float* __restrict__ prob /*[B,8h,8w,C] or null*/,
            float* __restrict__ score_out /*[B,8h,8w,C] or null*/)
{
    // C even: every channel pair is one 8-byte vector (vertex rows are 3C floats = 8-byte aligned).
    // A CTA produces the output pixels of `seg_cells` low-resolution cells of one output row (small shared-memory
    // footprint -> many resident CTAs to hide the streaming-store latency).
    extern __shared__ float smem_f[];
    const int No = 4 * C, W = 8 * w, H = 8 * h, N2 = No / 2, C2 = C / 2;
    const int c_lo = blockIdx.z * seg_cells, c_hi = min(c_lo + seg_cells, w);   // cells [c_lo, c_hi)
    const int s_lo = max(c_lo - 1, 0), s_hi = min(c_hi + 1, w);                 // source cells incl. halo
    float2* rowi = reinterpret_cast<float2*>(smem_f) - (size_t)s_lo * N2;       // [s_lo, s_hi) x N2, indexed by absolute cell
    float* sc = smem_f + (size_t)(seg_cells + 2) * No - (size_t)8 * c_lo * C;   // [8 seg_cells][C], indexed by absolute x
    const int y = blockIdx.x, n = blockIdx.y, t = threadIdx.x;
    // conv2d_transpose 16x16 / stride 8, SAME (pad 4): out[o] = sum_i in[i] * W[o - 8i + 4]
    const int my = y >> 3, ty = y & 7;
    const int iy0 = ty < 4 ? my - 1 : my, iy1 = iy0 + 1;
    const float wy0 = (iy0 >= 0 && iy0 < h) ? deconv_w(y - 8 * iy0 + 4, 16) : 0.f;
    const float wy1 = (iy1 >= 0 && iy1 < h) ? deconv_w(y - 8 * iy1 + 4, 16) : 0.f;
    const float2* r0 = reinterpret_cast<const float2*>(lr + ((size_t)n * h + min(max(iy0, 0), h - 1)) * w * No);
    const float2* r1 = reinterpret_cast<const float2*>(lr + ((size_t)n * h + min(max(iy1, 0), h - 1)) * w * No);
    for (int i = s_lo * N2 + t; i < s_hi * N2; i += 256) {
        const float2 a = __ldg(r0 + i), b = __ldg(r1 + i);
        rowi[i] = make_float2(fmaf(wy1, b.x, wy0 * a.x), fmaf(wy1, b.y, wy0 * a.y));
    }
    __syncthreads();
    const size_t rowbase = ((size_t)n * H + y) * W;
    // thread = (cell phase g, channel pair c2): a fixed channel pair, strided over the low-resolution cells of the
    // row; the 8 output pixels of a cell blend the same three source values with compile-time weights.  Consecutive
    // lanes hold consecutive channel pairs, so the vertex stores of a pixel are contiguous.
    const int groups = 256 / N2;
    const int g = t / N2, c2 = t - g * N2;
    if (g < groups) {
        const float2 bb = c2 < C2 ? make_float2(bias_s[2 * c2], bias_s[2 * c2 + 1])
                                  : make_float2(bias_v[2 * c2 - C], bias_v[2 * c2 + 1 - C]);
        const bool is_score = c2 < C2;
        float* vbase = vertex + rowbase * 3 * C + 2 * (c2 - C2);
        for (int mx = c_lo + g; mx < c_hi; mx += groups) {
            const float2 zero = make_float2(0.f, 0.f);
            const float2 vl = mx > 0 ? rowi[(mx - 1) * N2 + c2] : zero;
            const float2 vc = rowi[mx * N2 + c2];
            const float2 vr = mx + 1 < w ? rowi[(mx + 1) * N2 + c2] : zero;
#pragma unroll
            for (int tx = 0; tx < 8; tx++) {
                // x = 8 mx + tx: sources (mx-1, mx) with taps (tx+12, tx+4) for tx < 4, (mx, mx+1) with (tx+4, tx-4) otherwise
                const float wa = deconv_w(tx < 4 ? tx + 12 : tx + 4, 16), wb = deconv_w(tx < 4 ? tx + 4 : tx - 4, 16);
                const float2 a = tx < 4 ? vl : vc, b = tx < 4 ? vc : vr;
                float v0 = fmaf(wb, b.x, wa * a.x) + bb.x;
                float v1 = fmaf(wb, b.y, wa * a.y) + bb.y;
                const int x = 8 * mx + tx;
                if (is_score) {
                    v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f);  // `score` has a ReLU (vgg16_convs.py:141, network.py:160)
                    *reinterpret_cast<float2*>(sc + x * C + 2 * c2) = make_float2(v0, v1);
                    if (score_out) *reinterpret_cast<float2*>(score_out + (rowbase + x) * C + 2 * c2) = make_float2(v0, v1);
                } else {
                    *reinterpret_cast<float2*>(vbase + (size_t)x * 3 * C) = make_float2(v0, v1);
                }
            }
        }
    }
    __syncthreads();
    // arg-max over classes, lowest index wins ties (tf.argmax); softmax for prob_normalized (network.py:474-488)
    for (int x = 8 * c_lo + t; x < 8 * c_hi; x += 256) {
        const float* s = sc + x * C;
        float best = s[0];
        int bi = 0;
        for (int c = 1; c < C; c++)
            if (s[c] > best) { best = s[c]; bi = c; }
        label[rowbase + x] = bi;
        if (prob) {
            float sum = 0.f;
            for (int c = 0; c < C; c++) sum += expf(s[c] - best);
            float* pr = prob + (rowbase + x) * C;
            for (int c = 0; c < C; c++) pr[c] = expf(s[c] - best) / sum;
        }
    }
}

// generic bilinear transposed convolution (depthwise, diagonal filter): out[B,s*h,s*w,C] f32 from in[B,h,w,C] f32.
// Only used by tests / the un-fused reference path.
__global__ void __launch_bounds__(256)
k_deconv_bilinear(const float* __restrict__ in, float* __restrict__ out, int B, int h, int w, int C, int k, int s)
{
    const int H = h * s, W = w * s, pad = (k - s) / 2;
    const size_t total = (size_t)B * H * W * C;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        size_t r = idx / C;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const size_t n = r / H;
        float acc = 0.f;
        for (int iy = 0; iy < h; iy++) {
            int ky = y - s * iy + pad;
            if (ky < 0 || ky >= k) continue;
            for (int ix = 0; ix < w; ix++) {
                int kx = x - s * ix + pad;
                if (kx < 0 || kx >= k) continue;
                acc += deconv_w(ky, k) * deconv_w(kx, k) * in[((n * h + iy) * w + ix) * C + c];
            }
        }
        out[idx] = acc;
    }
}

}  // namespace pcnn

using namespace pcnn;

extern "C" int pcnn_lowres_heads(const void* score4, const void* score5, const void* vert4, const void* vert5,
                                 const float* w_score, const float* w_vertex, int B, int h, int w, int Cs, int Cv, int C,
                                 float* out, void* stream)
{
    PCNN_REQUIRE(score4 && score5 && vert4 && vert5 && w_score && w_vertex && out, "lowres_heads: NULL tensor pointer");
    PCNN_REQUIRE(h % 2 == 0 && w % 2 == 0, "lowres_heads: conv4 resolution must be even (got %d x %d)", h, w);
    size_t smem = sizeof(float) * ((size_t)Cs * C + (size_t)Cv * 3 * C + (size_t)kLrWarps * 2 * (Cs + Cv));
    PCNN_REQUIRE(smem <= 200 * 1024, "lowres_heads: weights do not fit shared memory");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_lowres_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    PCNN_REQUIRE((long long)B * h * w < 0x7fffffffLL, "lowres_heads: too many pixels");
    size_t npix = (size_t)B * h * w;
    int blocks = (int)std::min<size_t>((npix + 2 * kLrWarps - 1) / (2 * kLrWarps), (size_t)kNumSMs * 4);
    k_lowres_heads<<<blocks, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)score4, (const __nv_bfloat16*)score5,
                                                               (const __nv_bfloat16*)vert4, (const __nv_bfloat16*)vert5, w_score,
                                                               w_vertex, B, h, w, Cs, Cv, C, out);
    return check_launch("lowres_heads");
}

extern "C" int pcnn_up8_heads(const float* lowres, const float* bias_score, const float* bias_vertex, int B, int h, int w, int C,
                              int32_t* label, float* vertex, float* prob, float* score, void* stream)
{
    PCNN_REQUIRE(lowres && bias_score && bias_vertex && label && vertex, "up8_heads: NULL tensor pointer");
    PCNN_REQUIRE(C >= 1 && B >= 1 && h >= 1 && w >= 1, "up8_heads: bad shape");
    PCNN_REQUIRE(8 * h <= 65535 * 1 && B <= 65535, "up8_heads: image too tall for the launch grid");
    PCNN_REQUIRE(C % 2 == 0 && 2 * C <= 256, "up8_heads: num_classes must be even and <= 128 (got %d)", C);
    int seg_cells = w <= 20 ? w : 20;  // 160 output pixels per CTA
    size_t smem = sizeof(float) * ((size_t)(seg_cells + 2) * 4 * C + (size_t)8 * seg_cells * C);
    PCNN_REQUIRE(smem <= 200 * 1024, "up8_heads: segment does not fit shared memory (C = %d)", C);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_up8_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
    dim3 grid(8 * h, B, (w + seg_cells - 1) / seg_cells);
    k_up8_heads<<<grid, 256, smem, (cudaStream_t)stream>>>(lowres, bias_score, bias_vertex, h, w, C, seg_cells, label, vertex, prob,
                                                           score);
    return check_launch("up8_heads");
}

extern "C" int pcnn_deconv_bilinear(const float* in, float* out, int B, int h, int w, int C, int k, int s, void* stream)
{
    PCNN_REQUIRE(in && out && k >= s && (k - s) % 2 == 0, "deconv_bilinear: bad arguments");
    size_t total = (size_t)B * h * s * w * s * C;
    int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)kNumSMs * 16);
    k_deconv_bilinear<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, B, h, w, C, k, s);
    return check_launch("deconv_bilinear");
}
