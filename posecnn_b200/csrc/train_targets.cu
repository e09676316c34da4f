// train_targets.cu — training-side target generation and fused losses.
//
//   pcnn_vertex_targets_fwd      lib/gt_synthesize_layer/minibatch.py:543-602 (_generate_vertex_targets, the
//                                single-instance branch :578-599): per labelled pixel of a present class
//                                (dx, dy) / (|(dx, dy)| + 1e-10) toward the projected centre and log z; weights = W_INSIDE
//   pcnn_loss_cls_hard_fwd       lib/fcn/train.py:455-465 (loss_cross_entropy_single_frame) applied to the Hardlabel mask
//                                (hard_label_op_gpu.cu.cc:16-29) WITHOUT materialising the [B,H,W,C] mask
//   pcnn_smooth_l1_vertex_fwd    lib/fcn/train.py:564-573 (smooth_l1_loss_vertex)
// Losses: fixed 592-CTA grid, per-CTA partial sums in double, last CTA to finish reduces them in index order
// (run-to-run deterministic), optional gradient pass w.r.t. the first input.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "heads_common.cuh"

namespace pcnn {

constexpr int kLossBlocks = kNumSMs * 4;
constexpr int kLossThreads = 256;

// ---------------------------------------------------------------------------------------------
// deterministic two-value reduction: per-CTA partials, last CTA sums them in index order
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_reduce2(double& a, double& b, double* sh /*[2 * warps]*/)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { sh[2 * w] = a; sh[2 * w + 1] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0, sb = 0;
        for (int k = 0; k < nw; k++) { sa += sh[2 * k]; sb += sh[2 * k + 1]; }
        a = sa; b = sb;
    }
}

__device__ __forceinline__ bool finish_partials(double a, double b, double* partial /*[2 * grid]*/, unsigned* ticket, double& ta, double& tb)
{
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return false;
    if (threadIdx.x == 0) {
        __threadfence();
        const volatile double* vp = partial;
        double sa = 0, sb = 0;
        for (unsigned k = 0; k < gridDim.x; k++) { sa += vp[2 * k]; sb += vp[2 * k + 1]; }
        ta = sa; tb = sb;
        *ticket = 0;                                       // ready for the next launch
    }
    return threadIdx.x == 0;
}

// cross entropy over the pixels Hardlabel selects: gt != -1 and (gt > 0 or prob[gt] < threshold)
__global__ void __launch_bounds__(kLossThreads)
k_loss_cls_hard(const float* __restrict__ score /*log-softmax*/, const float* __restrict__ prob, const int* __restrict__ gt, unsigned npix,
                int C, float threshold, double* __restrict__ partial, unsigned* __restrict__ ticket, float* __restrict__ out /*[2]: loss, count*/)
{
    __shared__ double sh[2 * kLossThreads / 32];
    double s = 0, n = 0;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int g = __ldg(gt + p);
        if (g < 0 || g >= C) continue;                     // -1 = ignore (hard_label_op_gpu.cu.cc:24); out-of-range labels ignored
        if (g > 0 || __ldg(prob + (size_t)p * C + g) < threshold) { s -= (double)__ldg(score + (size_t)p * C + g); n += 1.0; }
    }
    block_reduce2(s, n, sh);
    double ts, tn;
    if (finish_partials(s, n, partial, ticket, ts, tn)) {
        out[0] = (float)(ts / (tn + 1e-10));               // train.py:463
        out[1] = (float)tn;
    }
}

// same loss from the RAW scores (the `score` layer output): log-softmax of the labelled class computed per selected pixel
// (network.py:491-506: x - max - log sum exp(x - max)), so the [B,H,W,C] log-probability tensor is never materialised
__global__ void __launch_bounds__(kLossThreads)
k_loss_cls_hard_raw(const float* __restrict__ score_raw, const float* __restrict__ prob, const int* __restrict__ gt, unsigned npix, int C,
                    float threshold, double* __restrict__ partial, unsigned* __restrict__ ticket, float* __restrict__ out)
{
    __shared__ double sh[2 * kLossThreads / 32];
    double s = 0, n = 0;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int g = __ldg(gt + p);
        if (g < 0 || g >= C) continue;
        if (g > 0 || __ldg(prob + (size_t)p * C + g) < threshold) {
            const float* sp = score_raw + (size_t)p * C;
            float m = __ldg(sp);
            for (int c = 1; c < C; c++) m = fmaxf(m, __ldg(sp + c));
            float se = 0.f;
            for (int c = 0; c < C; c++) se += expf(__ldg(sp + c) - m);
            s -= (double)(__ldg(sp + g) - m - logf(se));
            n += 1.0;
        }
    }
    block_reduce2(s, n, sh);
    double ts, tn;
    if (finish_partials(s, n, partial, ticket, ts, tn)) {
        out[0] = (float)(ts / (tn + 1e-10));
        out[1] = (float)tn;
    }
}

__global__ void __launch_bounds__(256)
k_loss_cls_hard_grad(const float* __restrict__ prob, const int* __restrict__ gt, unsigned npix, int C, float threshold,
                     const float* __restrict__ loss_out, float upstream, float* __restrict__ grad /*[npix, C]*/)
{
    const float scale = -upstream / (loss_out[1] + 1e-10f);
    const size_t total = (size_t)npix * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned p = (unsigned)(i / C);
        const int c = (int)(i - (size_t)p * C);
        const int g = __ldg(gt + p);
        const bool sel = g == c && g >= 0 && (g > 0 || __ldg(prob + (size_t)p * C + g) < threshold);
        grad[i] = sel ? scale : 0.f;
    }
}

// smooth L1 on weighted differences (train.py:564-573)
__device__ __forceinline__ float sl1_term(float pred, float targ, float wgt, float sigma2, float& dterm)
{
    const float diff = wgt * (pred - targ);
    const float ad = fabsf(diff);
    const bool quad = ad < 1.f / sigma2;
    dterm = quad ? diff * sigma2 : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
    return quad ? diff * diff * (sigma2 * 0.5f) : ad - 0.5f / sigma2;
}

__global__ void __launch_bounds__(kLossThreads)
k_smooth_l1_vertex(const float* __restrict__ pred, const float* __restrict__ targ, const float* __restrict__ wgt, size_t n4 /*float4 count*/,
                   size_t n, float sigma2, double* __restrict__ partial, unsigned* __restrict__ ticket, float* __restrict__ out /*[2]: loss, sum w*/)
{
    __shared__ double sh[2 * kLossThreads / 32];
    double s = 0, sw = 0;
    float d;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 w4 = ld_stream_f4(reinterpret_cast<const float4*>(wgt) + i);
        sw += (double)w4.x + (double)w4.y + (double)w4.z + (double)w4.w;
        if (w4.x == 0.f && w4.y == 0.f && w4.z == 0.f && w4.w == 0.f) continue;   // zero weight -> zero term: skip the other loads
        const float4 p4 = ld_stream_f4(reinterpret_cast<const float4*>(pred) + i);
        const float4 t4 = ld_stream_f4(reinterpret_cast<const float4*>(targ) + i);
        s += (double)sl1_term(p4.x, t4.x, w4.x, sigma2, d) + (double)sl1_term(p4.y, t4.y, w4.y, sigma2, d) +
             (double)sl1_term(p4.z, t4.z, w4.z, sigma2, d) + (double)sl1_term(p4.w, t4.w, w4.w, sigma2, d);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = 4 * n4; i < n; i++) { sw += (double)wgt[i]; s += (double)sl1_term(pred[i], targ[i], wgt[i], sigma2, d); }
    block_reduce2(s, sw, sh);
    double ts, tw;
    if (finish_partials(s, sw, partial, ticket, ts, tw)) {
        out[0] = (float)(ts / (tw + 1e-10));               // train.py:572
        out[1] = (float)tw;
    }
}

__global__ void __launch_bounds__(256)
k_smooth_l1_vertex_grad(const float* __restrict__ pred, const float* __restrict__ targ, const float* __restrict__ wgt, size_t n, float sigma2,
                        const float* __restrict__ loss_out, float upstream, float* __restrict__ grad)
{
    const float scale = upstream / (loss_out[1] + 1e-10f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float w = wgt[i];
        float d = 0.f;
        if (w != 0.f) sl1_term(pred[i], targ[i], w, sigma2, d);
        grad[i] = w * d * scale;                           // d in_loss / d pred = w * g(diff); smoothL1_sign carries no gradient
    }
}

// ---------------------------------------------------------------------------------------------
// fused vertex loss: smooth_l1_loss_vertex(vertex_pred, targets(label, centers), weights(label, centers)) without
// the two [B,H,W,3C] target / weight tensors (5.2 GB at batch 32): only the three channels of a labelled pixel's own
// class carry weight, so the kernel reads 12 bytes of vertex_pred per foreground pixel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pixel_targets(const int* __restrict__ label, const float* __restrict__ centers, unsigned pix, int HW,
                                              int W, int C, int& cls, float t[3])
{
    const int l = __ldg(label + pix);
    if (l <= 0 || l >= C) return false;
    const int b = pix / HW, p = pix - b * HW;
    const float* cen = centers + ((size_t)b * C + l) * 3;
    const float z = cen[2];
    if (!(z > 0.f)) return false;
    const double dx = (double)cen[0] - (double)(p % W), dy = (double)cen[1] - (double)(p / W);
    const double nrm = sqrt(dx * dx + dy * dy) + 1e-10;
    t[0] = (float)(dx / nrm); t[1] = (float)(dy / nrm); t[2] = (float)log((double)z);
    cls = l;
    return true;
}

// materialised targets / weights (drop-in for the data layer's blobs): the tensors are >97 % zeros, so they are
// cleared with two memset nodes and only the three channels of each labelled pixel's own class are written
__global__ void __launch_bounds__(256)
k_vertex_targets_sparse(const int* __restrict__ label, const float* __restrict__ centers, unsigned npix, int HW, int W, int C, float w_inside,
                        float* __restrict__ targets, float* __restrict__ weights)
{
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        int cls;
        float t[3];
        if (!pixel_targets(label, centers, pix, HW, W, C, cls, t)) continue;
        const size_t o = (size_t)pix * 3 * C + 3 * cls;
#pragma unroll
        for (int k = 0; k < 3; k++) { targets[o + k] = t[k]; weights[o + k] = w_inside; }
    }
}

// multi-instance branch of _generate_vertex_targets (minibatch.py:549-573): several instances of one class in an image
// are told apart by an instance-mask image; instance i = (cls, mask id = cls_indexes_old[i] + 1, projected centre, z)
// owns the pixels with mask == id AND label == cls.  The reference loops the instances in order and overwrites, so the
// LAST matching instance wins.  instances [B, I, 5] = (cls, mask_id, cx, cy, z); z <= 0 marks an unused slot.
__global__ void __launch_bounds__(256)
k_vertex_targets_instances(const int* __restrict__ label, const int* __restrict__ mask, const float* __restrict__ inst, unsigned npix,
                           int HW, int W, int C, int I, float w_inside, float* __restrict__ targets, float* __restrict__ weights)
{
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        const int l = __ldg(label + pix);
        if (l <= 0 || l >= C) continue;
        const int m = __ldg(mask + pix);
        const int b = pix / HW, p = pix - b * HW;
        const float* rows = inst + (size_t)b * I * 5;
        int hit = -1;
        for (int i = 0; i < I; i++)
            if (rows[5 * i + 4] > 0.f && (int)rows[5 * i] == l && (int)rows[5 * i + 1] == m) hit = i;
        if (hit < 0) continue;
        const float* r = rows + 5 * hit;
        const double dx = (double)r[2] - (double)(p % W), dy = (double)r[3] - (double)(p / W);
        const double nrm = sqrt(dx * dx + dy * dy) + 1e-10;
        const size_t o = (size_t)pix * 3 * C + 3 * l;
        targets[o] = (float)(dx / nrm); targets[o + 1] = (float)(dy / nrm); targets[o + 2] = (float)log((double)r[4]);
        weights[o] = w_inside; weights[o + 1] = w_inside; weights[o + 2] = w_inside;
    }
}

// ---------------------------------------------------------------------------------------------
// pose blob and meta_data packing of the data layer (minibatch.py:440-451, 474-492):
//   pose_blob rows [image, cls, 0, 0, 0, 0, mat2quat(R) (w, x, y, z), T] for every listed instance, images in order;
//   meta_data[48]: K * im_scale with K[2][2] = 1 in [0:9], its (pseudo-)inverse in [9:18], zeros elsewhere, FLIP_X signs.
// mat2quat is transforms3d's (Bar-Itzhack): eigenvector of the largest eigenvalue of the symmetric 4x4 K matrix, here
// by cyclic Jacobi rotations in double, w made non-negative.
// ---------------------------------------------------------------------------------------------
__device__ void mat2quat_d(const float* __restrict__ rt /*3x4 row-major*/, float q[4])
{
    // transforms3d: `Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = M.flat` (row-major flat order: Qyx = M[0][1], Qxy = M[1][0], ...)
    const double Qxx = rt[0], Qyx = rt[1], Qzx = rt[2], Qxy = rt[4], Qyy = rt[5], Qzy = rt[6], Qxz = rt[8], Qyz = rt[9], Qzz = rt[10];
    double A[4][4] = {{Qxx - Qyy - Qzz, Qyx + Qxy, Qzx + Qxz, Qyz - Qzy},
                      {Qyx + Qxy, Qyy - Qxx - Qzz, Qzy + Qyz, Qzx - Qxz},
                      {Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, Qxy - Qyx},
                      {Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz}};
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) A[i][j] /= 3.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0;
        for (int i = 0; i < 4; i++)
            for (int j = i + 1; j < 4; j++) off += A[i][j] * A[i][j];
        if (off < 1e-30) break;
        for (int pI = 0; pI < 3; pI++)
            for (int qI = pI + 1; qI < 4; qI++) {
                if (fabs(A[pI][qI]) < 1e-300) continue;
                const double theta = (A[qI][qI] - A[pI][pI]) / (2.0 * A[pI][qI]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 4; k++) {
                    const double akp = A[k][pI], akq = A[k][qI];
                    A[k][pI] = c * akp - sn * akq; A[k][qI] = sn * akp + c * akq;
                }
                for (int k = 0; k < 4; k++) {
                    const double apk = A[pI][k], aqk = A[qI][k];
                    A[pI][k] = c * apk - sn * aqk; A[qI][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 4; k++) {
                    const double vkp = V[k][pI], vkq = V[k][qI];
                    V[k][pI] = c * vkp - sn * vkq; V[k][qI] = sn * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int k = 1; k < 4; k++)
        if (A[k][k] > A[best][best]) best = k;
    double w = V[3][best], x = V[0][best], y = V[1][best], z = V[2][best];   // vecs[[3, 0, 1, 2], argmax]
    if (w < 0) { w = -w; x = -x; y = -y; z = -z; }
    q[0] = (float)w; q[1] = (float)x; q[2] = (float)y; q[3] = (float)z;
}

__global__ void __launch_bounds__(256)
k_pack_pose_meta(const float* __restrict__ poses /*[B,I,12]*/, const int* __restrict__ cls /*[B,I], < 0 = unused*/,
                 const float* __restrict__ intr /*[B,9]*/, int B, int I, float im_scale, int flip_x, float* __restrict__ pose_blob /*[B*I,13]*/,
                 int* __restrict__ num_rows, float* __restrict__ meta /*[B,48]*/)
{
    __shared__ int s_off[1025];
    const int t = threadIdx.x, n = B * I;
    if (t == 0) {
        int run = 0;
        for (int k = 0; k < n; k++) { s_off[k] = run; run += cls[k] >= 0 ? 1 : 0; }
        s_off[n] = run;
        *num_rows = run;
    }
    __syncthreads();
    for (int k = t; k < n; k += blockDim.x) {
        float* row = pose_blob + (size_t)k * 13;
        if (k >= s_off[n])
            for (int j = 0; j < 13; j++) row[j] = 0.f;      // rows beyond the count are zero (capacity buffer)
    }
    __syncthreads();
    for (int k = t; k < n; k += blockDim.x) {
        if (cls[k] < 0) continue;
        const float* rt = poses + (size_t)k * 12;
        float* row = pose_blob + (size_t)s_off[k] * 13;
        row[0] = (float)(k / I); row[1] = (float)cls[k];
        row[2] = row[3] = row[4] = row[5] = 0.f;            // box: "fill later" (minibatch.py:447)
        mat2quat_d(rt, row + 6);
        row[10] = rt[3]; row[11] = rt[7]; row[12] = rt[11];
    }
    for (int b = t; b < B; b += blockDim.x) {
        float* m = meta + (size_t)b * 48;
        for (int j = 0; j < 48; j++) m[j] = 0.f;
        double K[9];
        for (int j = 0; j < 9; j++) K[j] = (double)(float)(intr[b * 9 + j]) * (double)im_scale;
        K[8] = 1.0;
        // inverse by cofactors (np.linalg.pinv of the non-singular 3x3; agreement 1e-6 relative after the float32 cast)
        const double c00 = K[4] * K[8] - K[5] * K[7], c01 = K[5] * K[6] - K[3] * K[8], c02 = K[3] * K[7] - K[4] * K[6];
        const double det = K[0] * c00 + K[1] * c01 + K[2] * c02;
        double Ki[9] = {c00 / det, (K[2] * K[7] - K[1] * K[8]) / det, (K[1] * K[5] - K[2] * K[4]) / det,
                        c01 / det, (K[0] * K[8] - K[2] * K[6]) / det, (K[2] * K[3] - K[0] * K[5]) / det,
                        c02 / det, (K[1] * K[6] - K[0] * K[7]) / det, (K[0] * K[4] - K[1] * K[3]) / det};
        for (int j = 0; j < 9; j++) { m[j] = (float)K[j]; m[9 + j] = (float)Ki[j]; }
        if (flip_x) { m[0] = -m[0]; m[9] = -m[9]; m[11] = -m[11]; }   // minibatch.py:488-491
    }
}

__global__ void __launch_bounds__(kLossThreads)
k_vertex_loss_fused(const float* __restrict__ pred, const float* __restrict__ lowres, const float* __restrict__ bias_v,
                    const int* __restrict__ label, const float* __restrict__ centers, unsigned npix, int HW,
                    int W, int C, float w_inside, float sigma2, double* __restrict__ partial, unsigned* __restrict__ ticket,
                    float* __restrict__ out)
{
    // pred == NULL: the labelled pixels' three vertex values are formed on demand from the 1/8-resolution head tensor with
    // k_up8_heads' own operation sequence (heads_common.cuh) — bit-identical to reading the dense vertex_pred
    __shared__ double sh[2 * kLossThreads / 32];
    double s = 0, sw = 0;
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        int cls;
        float t[3], d;
        if (!pixel_targets(label, centers, pix, HW, W, C, cls, t)) continue;
        if (pred) {
            const float* pp = pred + (size_t)pix * 3 * C + 3 * cls;
#pragma unroll
            for (int k = 0; k < 3; k++) s += (double)sl1_term(__ldg(pp + k), t[k], w_inside, sigma2, d);
        } else {
            const int b = pix / HW, pp = pix - b * HW, y = pp / W, x = pp - y * W;
#pragma unroll
            for (int k = 0; k < 3; k++)
                s += (double)sl1_term(up8_value(lowres, b, (HW / W) >> 3, W >> 3, 4 * C, C + 3 * cls + k, y, x, __ldg(bias_v + 3 * cls + k)), t[k],
                                      w_inside, sigma2, d);
        }
        sw += 3.0 * (double)w_inside;
    }
    block_reduce2(s, sw, sh);
    double ts, tw;
    if (finish_partials(s, sw, partial, ticket, ts, tw)) {
        out[0] = (float)(ts / (tw + 1e-10));
        out[1] = (float)tw;
    }
}

__global__ void __launch_bounds__(256)
k_vertex_loss_fused_grad(const float* __restrict__ pred, const int* __restrict__ label, const float* __restrict__ centers, unsigned npix,
                         int HW, int W, int C, float w_inside, float sigma2, const float* __restrict__ loss_out, float upstream,
                         float* __restrict__ grad /*zero-filled*/)
{
    const float scale = upstream / (loss_out[1] + 1e-10f);
    for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += gridDim.x * blockDim.x) {
        int cls;
        float t[3], d;
        if (!pixel_targets(label, centers, pix, HW, W, C, cls, t)) continue;
        const size_t o = (size_t)pix * 3 * C + 3 * cls;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            sl1_term(__ldg(pred + o + k), t[k], w_inside, sigma2, d);
            grad[o + k] = w_inside * d * scale;
        }
    }
}

}  // namespace pcnn

using namespace pcnn;

extern "C" int pcnn_vertex_loss_fused_fwd(const float* pred, const int32_t* label, const float* centers, int B, int H, int W, int C,
                                          float w_inside, float sigma, float* loss_out, float upstream, float* grad_pred,
                                          void* workspace, size_t workspace_bytes, void* stream)
{
    PCNN_REQUIRE(pred && label && centers && loss_out && workspace, "vertex_loss_fused: NULL tensor pointer");
    PCNN_REQUIRE(sigma > 0.f && B >= 1 && H >= 1 && W >= 1 && C >= 1, "vertex_loss_fused: bad arguments");
    PCNN_REQUIRE((unsigned long long)B * H * W < 0xffffffffULL, "vertex_loss_fused: too many pixels");
    size_t need = 0;
    pcnn_train_loss_workspace_bytes(&need);
    PCNN_REQUIRE(workspace_bytes >= need, "vertex_loss_fused: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)(partial + 2 * kLossBlocks);
    const unsigned npix = (unsigned)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    k_vertex_loss_fused<<<kLossBlocks, kLossThreads, 0, st>>>(pred, nullptr, nullptr, label, centers, npix, H * W, W, C, w_inside, sigma * sigma,
                                                               partial, ticket, loss_out);
    if (grad_pred) {
        cudaMemsetAsync(grad_pred, 0, sizeof(float) * (size_t)npix * 3 * C, st);
        k_vertex_loss_fused_grad<<<kNumSMs * 16, 256, 0, st>>>(pred, label, centers, npix, H * W, W, C, w_inside, sigma * sigma, loss_out,
                                                                upstream, grad_pred);
    }
    return check_launch("vertex_loss_fused");
}

// the same loss with the vertex head given as the 1/8-resolution head tensor `lowres` [B,H/8,W/8,4C] (channels C.. = vertex) + the
// vertex_pred bias [3C]: no dense vertex_pred tensor is needed anywhere in the training step
extern "C" int pcnn_vertex_loss_fused_lowres_fwd(const float* lowres, const float* bias_vertex, const int32_t* label, const float* centers, int B,
                                                 int H, int W, int C, float w_inside, float sigma, float* loss_out, void* workspace,
                                                 size_t workspace_bytes, void* stream)
{
    PCNN_REQUIRE(lowres && bias_vertex && label && centers && loss_out && workspace, "vertex_loss_fused_lowres: NULL tensor pointer");
    PCNN_REQUIRE(sigma > 0.f && B >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0 && C >= 1, "vertex_loss_fused_lowres: bad arguments");
    PCNN_REQUIRE((unsigned long long)B * H * W < 0xffffffffULL, "vertex_loss_fused_lowres: too many pixels");
    size_t need = 0;
    pcnn_train_loss_workspace_bytes(&need);
    PCNN_REQUIRE(workspace_bytes >= need, "vertex_loss_fused_lowres: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)(partial + 2 * kLossBlocks);
    k_vertex_loss_fused<<<kLossBlocks, kLossThreads, 0, (cudaStream_t)stream>>>(nullptr, lowres, bias_vertex, label, centers, (unsigned)B * H * W, H * W,
                                                                                 W, C, w_inside, sigma * sigma, partial, ticket, loss_out);
    return check_launch("vertex_loss_fused_lowres");
}

extern "C" int pcnn_train_loss_workspace_bytes(size_t* bytes)
{
    PCNN_REQUIRE(bytes, "train_loss_workspace_bytes: NULL pointer");
    *bytes = sizeof(double) * 2 * kLossBlocks + 16;        // per-CTA partials + ticket (must be zero on first use)
    return PCNN_OK;
}

extern "C" int pcnn_vertex_targets_fwd(const int32_t* label, const float* centers, int B, int H, int W, int C, float w_inside,
                                       float* targets, float* weights, void* stream)
{
    PCNN_REQUIRE(label && centers && targets && weights, "vertex_targets: NULL tensor pointer");
    PCNN_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1, "vertex_targets: bad shape");
    PCNN_REQUIRE((unsigned long long)B * H * W < 0xffffffffULL, "vertex_targets: too many pixels");
    const unsigned npix = (unsigned)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(targets, 0, sizeof(float) * (size_t)npix * 3 * C, st);
    cudaMemsetAsync(weights, 0, sizeof(float) * (size_t)npix * 3 * C, st);
    k_vertex_targets_sparse<<<kNumSMs * 16, 256, 0, st>>>(label, centers, npix, H * W, W, C, w_inside, targets, weights);
    return check_launch("vertex_targets");
}

extern "C" int pcnn_loss_cls_hard_fwd(const float* score, const float* prob, const int32_t* gt, int B, int H, int W, int C, float threshold,
                                      float* loss_out, float upstream, float* grad_score, void* workspace, size_t workspace_bytes,
                                      void* stream)
{
    PCNN_REQUIRE(score && prob && gt && loss_out && workspace, "loss_cls_hard: NULL tensor pointer");
    PCNN_REQUIRE((unsigned long long)B * H * W < 0xffffffffULL, "loss_cls_hard: too many pixels");
    size_t need = 0;
    pcnn_train_loss_workspace_bytes(&need);
    PCNN_REQUIRE(workspace_bytes >= need, "loss_cls_hard: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)(partial + 2 * kLossBlocks);
    const unsigned npix = (unsigned)B * H * W;
    k_loss_cls_hard<<<kLossBlocks, kLossThreads, 0, (cudaStream_t)stream>>>(score, prob, gt, npix, C, threshold, partial, ticket, loss_out);
    if (grad_score)
        k_loss_cls_hard_grad<<<kNumSMs * 16, 256, 0, (cudaStream_t)stream>>>(prob, gt, npix, C, threshold, loss_out, upstream, grad_score);
    return check_launch("loss_cls_hard");
}

extern "C" int pcnn_smooth_l1_vertex_fwd(const float* pred, const float* targets, const float* weights, size_t n, float sigma,
                                         float* loss_out, float upstream, float* grad_pred, void* workspace, size_t workspace_bytes,
                                         void* stream)
{
    PCNN_REQUIRE(pred && targets && weights && loss_out && workspace, "smooth_l1_vertex: NULL tensor pointer");
    PCNN_REQUIRE(sigma > 0.f, "smooth_l1_vertex: sigma must be positive");
    PCNN_REQUIRE(((uintptr_t)pred | (uintptr_t)targets | (uintptr_t)weights) % 16 == 0, "smooth_l1_vertex: tensors must be 16-byte aligned");
    size_t need = 0;
    pcnn_train_loss_workspace_bytes(&need);
    PCNN_REQUIRE(workspace_bytes >= need, "smooth_l1_vertex: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)(partial + 2 * kLossBlocks);
    k_smooth_l1_vertex<<<kLossBlocks, kLossThreads, 0, (cudaStream_t)stream>>>(pred, targets, weights, n / 4, n, sigma * sigma, partial, ticket,
                                                                                loss_out);
    if (grad_pred)
        k_smooth_l1_vertex_grad<<<kNumSMs * 16, 256, 0, (cudaStream_t)stream>>>(pred, targets, weights, n, sigma * sigma, loss_out, upstream,
                                                                                  grad_pred);
    return check_launch("smooth_l1_vertex");
}

extern "C" int pcnn_vertex_targets_instances_fwd(const int32_t* label, const int32_t* mask, const float* instances, int B, int H, int W,
                                                 int C, int I, float w_inside, float* targets, float* weights, void* stream)
{
    PCNN_REQUIRE(label && mask && instances && targets && weights, "vertex_targets_instances: NULL tensor pointer");
    PCNN_REQUIRE(B >= 1 && H >= 1 && W >= 1 && C >= 1 && I >= 1, "vertex_targets_instances: bad shape");
    PCNN_REQUIRE((unsigned long long)B * H * W < 0xffffffffULL, "vertex_targets_instances: too many pixels");
    const unsigned npix = (unsigned)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(targets, 0, sizeof(float) * (size_t)npix * 3 * C, st);
    cudaMemsetAsync(weights, 0, sizeof(float) * (size_t)npix * 3 * C, st);
    k_vertex_targets_instances<<<kNumSMs * 16, 256, 0, st>>>(label, mask, instances, npix, H * W, W, C, I, w_inside, targets, weights);
    return check_launch("vertex_targets_instances");
}

extern "C" int pcnn_pack_pose_meta_fwd(const float* poses, const int32_t* cls, const float* intrinsics, int B, int I, float im_scale,
                                       int flip_x, float* pose_blob, int32_t* num_rows, float* meta, void* stream)
{
    PCNN_REQUIRE(poses && cls && intrinsics && pose_blob && num_rows && meta, "pack_pose_meta: NULL tensor pointer");
    PCNN_REQUIRE(B >= 1 && I >= 1 && B * I <= 1024, "pack_pose_meta: at most 1024 instance slots per batch (got %d x %d)", B, I);
    k_pack_pose_meta<<<1, 256, 0, (cudaStream_t)stream>>>(poses, cls, intrinsics, B, I, im_scale, flip_x, pose_blob, num_rows, meta);
    return check_launch("pack_pose_meta");
}

extern "C" int pcnn_loss_cls_hard_raw_fwd(const float* score_raw, const float* prob, const int32_t* gt, int B, int H, int W, int C,
                                          float threshold, float* loss_out, void* workspace, size_t workspace_bytes, void* stream)
{
    PCNN_REQUIRE(score_raw && prob && gt && loss_out && workspace, "loss_cls_hard_raw: NULL tensor pointer");
    PCNN_REQUIRE((unsigned long long)B * H * W < 0xffffffffULL, "loss_cls_hard_raw: too many pixels");
    size_t need = 0;
    pcnn_train_loss_workspace_bytes(&need);
    PCNN_REQUIRE(workspace_bytes >= need, "loss_cls_hard_raw: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* partial = (double*)workspace;
    unsigned* ticket = (unsigned*)(partial + 2 * kLossBlocks);
    k_loss_cls_hard_raw<<<kLossBlocks, kLossThreads, 0, (cudaStream_t)stream>>>(score_raw, prob, gt, (unsigned)B * H * W, C, threshold, partial, ticket,
                                                                                loss_out);
    return check_launch("loss_cls_hard_raw");
}
