// nms_pose.cu — device-side per-class non-maximum suppression + pose assembly.
//
// Replaces the host step that follows the network at test time:
//   lib/utils/nms.py:3-32        greedy NMS over the Hough ROIs [batch, cls, x1, y1, x2, y2, score], a box is
//                                dropped when IoU(+1 convention) > thresh with an already kept box OF THE SAME CLASS
//   lib/fcn/test.py:197-211      rois / poses_init / poses_pred <- [keep]; poses[i, :4] = poses_pred[i, 4c : 4c + 4]
// so that ROIs never leave the device between the Hough op and the final [roi | pose] records (the only D2H read
// left per frame is the record buffer itself, and the all-gather payload of posecnn_b200/parallel.py is final).
//
// Canonical order (the reference's argsort()[::-1] is not stable across numpy versions when scores tie):
// descending score, ties by DESCENDING row index — what a stable ascending argsort, reversed, yields.
// The reference ignores the batch column (it only ever runs batch 1): per_image = 0 reproduces that, per_image = 1
// suppresses within (image, class) only — identical for batch 1.
//
// One CTA; N <= 1152 rows (PCNN_HOUGH_MAX_ROWS).  fp32 arithmetic in numpy's operation order, no FMA contraction.
#include <cuda_runtime.h>

#include "common.cuh"

namespace pcnn {

constexpr int kNmsThreads = 1024;

struct NmsBox {
    float x1, y1, x2, y2, score, area;
    int cls, img;
};

__device__ __forceinline__ bool nms_overlaps(const NmsBox& a, const NmsBox& b, float thresh)
{
    // nms.py:20-27, float32 like numpy: w = max(0, xx2 - xx1 + 1), inter / (area_i + area_j - inter) > thresh
    const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
    const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
    const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
    const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
    return ovr > thresh;
}

// before(i, j): box i is processed before box j (score descending, ties: larger row index first)
__device__ __forceinline__ bool nms_before(float si, int i, float sj, int j) { return si > sj || (si == sj && i > j); }

__global__ void __launch_bounds__(kNmsThreads)
k_nms_pose(const float* __restrict__ rois /*[cap,7]*/, const float* __restrict__ poses_init /*[cap,7]*/,
           const float* __restrict__ poses_pred /*[cap,4C] or null*/, const int* __restrict__ num_rois_dev, int n_rows, int cap,
           int C, float thresh, int per_image, int* __restrict__ keep /*[cap]*/, float* __restrict__ out_rois /*[cap,7]*/,
           float* __restrict__ out_poses /*[cap,7]*/, int* __restrict__ num_keep)
{
    extern __shared__ unsigned char nms_smem[];
    // the reference always sees at least one row (the all-zero dummy row of hough_voting_gpu_op.cc:381-383)
    int N = num_rois_dev ? max(*num_rois_dev, 1) : n_rows;
    N = min(N, cap);
    NmsBox* box = reinterpret_cast<NmsBox*>(nms_smem);
    int* grouped = reinterpret_cast<int*>(box + cap);   // row ids ordered by (image, processing order)
    int* rank = grouped + cap;                           // processing order of row i
    int* gstart = rank + cap;                            // [cap + 1] segment start of group g (a group = image, or everything)
    unsigned char* dead = reinterpret_cast<unsigned char*>(gstart + cap + 1);
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;

    for (int i = t; i < N; i += kNmsThreads) {
        const float* r = rois + (size_t)i * 7;
        NmsBox b;
        b.img = per_image ? (int)r[0] : 0;
        b.cls = (int)r[1];
        b.x1 = r[2]; b.y1 = r[3]; b.x2 = r[4]; b.y2 = r[5]; b.score = r[6];
        b.area = __fmul_rn(__fadd_rn(__fsub_rn(b.x2, b.x1), 1.f), __fadd_rn(__fsub_rn(b.y2, b.y1), 1.f));   // nms.py:12
        box[i] = b;
        dead[i] = 0;
    }
    __syncthreads();
    // processing rank and grouped rank by counting (N <= 1152: ~1.3 M comparisons for the CTA)
    int ngroups = 1;
    for (int i = t; i < N; i += kNmsThreads) {
        const float si = box[i].score;
        const int gi = box[i].img;
        int r = 0, gr = 0;
        for (int j = 0; j < N; j++) {
            const bool bef = nms_before(box[j].score, j, si, i);
            r += bef;
            const int gj = box[j].img;
            gr += (gj < gi) || (gj == gi && bef);
        }
        // output position key: processing order inside the row's group; with per_image the groups (images) come in
        // ascending order, like the reference op loops the batch (hough_voting_gpu_op.cc:369-377) — so the rows of
        // image shards, concatenated in rank order, are the rows of the whole batch
        rank[i] = gr;
        grouped[gr] = i;
        (void)r;
    }
    __syncthreads();
    // group boundaries in the grouped order
    for (int p = t; p < N; p += kNmsThreads) {
        const int g = box[grouped[p]].img;
        if (p == 0 || box[grouped[p - 1]].img != g) gstart[p] = 1; else gstart[p] = 0;
    }
    __syncthreads();
    // greedy suppression: one warp per group (image); groups are independent, inside a group the scan is sequential in
    // processing order and the lanes sweep the boxes behind the current one
    for (int p0 = 0; p0 < N; p0++) {
        if (!gstart[p0]) continue;                       // warp-uniform scan for group heads
        if ((ngroups++ - 1) % (kNmsThreads / 32) != w) continue;
        int p1 = p0 + 1;
        while (p1 < N && !gstart[p1]) p1++;
        for (int p = p0; p < p1; p++) {
            const int i = grouped[p];
            __syncwarp();
            if (dead[i]) continue;
            const NmsBox bi = box[i];
            for (int q = p + 1 + lane; q < p1; q += 32) {
                const int j = grouped[q];
                if (!dead[j] && box[j].cls == bi.cls && nms_overlaps(bi, box[j], thresh)) dead[j] = 1;
            }
        }
    }
    __syncthreads();
    // compaction in (image, processing order) + pose assembly (test.py:204-211); one group = plain processing order
    for (int i = t; i < N; i += kNmsThreads) {
        if (dead[i]) continue;
        const int ri = rank[i];
        int pos = 0;
        for (int j = 0; j < N; j++) pos += (!dead[j]) && rank[j] < ri;
        keep[pos] = i;
        const float* r = rois + (size_t)i * 7;
        const float* pi = poses_init + (size_t)i * 7;
        float* orow = out_rois + (size_t)pos * 7;
        float* prow = out_poses + (size_t)pos * 7;
#pragma unroll
        for (int k = 0; k < 7; k++) { orow[k] = r[k]; prow[k] = pi[k]; }
        const int c = box[i].cls;
        if (poses_pred && c >= 0 && c < C) {
#pragma unroll
            for (int k = 0; k < 4; k++) prow[k] = poses_pred[(size_t)i * 4 * C + 4 * c + k];
        }
    }
    if (t == 0) {
        int nk = 0;
        for (int j = 0; j < N; j++) nk += !dead[j];
        *num_keep = nk;
    }
    // rows beyond the kept count: zero (fixed-size records downstream)
    __syncthreads();
    {
        int nk = 0;
        for (int j = 0; j < N; j++) nk += !dead[j];
        for (int i = nk * 7 + t; i < cap * 7; i += kNmsThreads) { out_rois[i] = 0.f; out_poses[i] = 0.f; }
        for (int i = nk + t; i < cap; i += kNmsThreads) keep[i] = -1;
    }
}

}  // namespace pcnn

using namespace pcnn;

extern "C" int pcnn_nms_pose_fwd(const float* rois, const float* poses_init, const float* poses_pred, const int* num_rois_dev,
                                 int num_rows, int capacity, int num_classes, float thresh, int per_image, int* keep,
                                 float* out_rois, float* out_poses, int* num_keep, void* stream)
{
    PCNN_REQUIRE(rois && poses_init && keep && out_rois && out_poses && num_keep, "nms_pose: NULL tensor pointer");
    PCNN_REQUIRE(capacity >= 1 && capacity <= PCNN_HOUGH_MAX_ROWS, "nms_pose: capacity %d outside [1, %d]", capacity,
                 PCNN_HOUGH_MAX_ROWS);
    PCNN_REQUIRE(num_rois_dev || (num_rows >= 0 && num_rows <= capacity), "nms_pose: num_rows %d outside [0, capacity]", num_rows);
    PCNN_REQUIRE(!poses_pred || num_classes >= 1, "nms_pose: poses_pred needs num_classes >= 1");
    size_t smem = (size_t)capacity * (sizeof(NmsBox) + 2 * sizeof(int) + 1) + (size_t)(capacity + 1) * sizeof(int) + 16;
    PCNN_SMEM_OPTIN(k_nms_pose, 100 * 1024, "nms_pose");
    PCNN_REQUIRE(smem <= 100 * 1024, "nms_pose: capacity %d does not fit shared memory", capacity);
    k_nms_pose<<<1, kNmsThreads, smem, (cudaStream_t)stream>>>(rois, poses_init, poses_pred, num_rois_dev, num_rows, capacity,
                                                               num_classes, thresh, per_image, keep, out_rois, out_poses, num_keep);
    return check_launch("nms_pose");
}
