// cpu_hough_ransac.cpp — standalone restatement of the reference's CPU `Houghvoting` op
// (lib/hough_voting_layer): pre-emptive RANSAC over 2-pixel direction-line intersections.
//
// TEST INFRASTRUCTURE / CPU BASELINE ONLY (BASELINE.json configs[0], bench.py cpu_baseline and
// --impl reference).  The original cannot be built here or on the GPU box (TensorFlow 1.x,
// OpenCV C++, nlopt), so it is restated with:
//   cv::solve(DECOMP_SVD) on an n x 2 system  -> 2-unknown least squares (normal equations, double)
//   cv::projectPoints(rvec = 0, no distortion) -> pinhole formula in double
//   cv::norm                                   -> sqrt of the double dot product
// libstdc++ <random> is reused, so the RNG streams (mt19937 seed 1305 + tid, thread_rand.cpp:46-67;
// default-seeded mt19937 + negative_binomial in countInliers2D, hough_voting_op.cc:420-421) match.
//
// Restated functions (paths relative to /root/reference/lib/hough_voting_layer):
//   getLabels hough_voting_op.cc:287-308        getWorkingQueue :364-386
//   countInliers2D :408-448                     compute_width_height :451-480
//   updateHyp2D / filterInliers2D :483-513      estimateCenter :516-857
//   TransHyp ransac.h:40-142                    Hypothesis::calcCenter Hypothesis.cpp:96-118
//   HoughvotingOp::Compute hough_voting_op.cc:104-236 (batch loop, dummy row cls = -1)
// Known quirks kept on purpose: mean LOG depth without exp (ransac.h:105-116);
// OpenMP is OFF in the reference build (lib/make.sh:54-60), so threads = 1 is "as shipped".
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <random>
#include <vector>

namespace {

struct P2 {
    double x, y;
};

struct TransHyp {  // ransac.h:40-142
    int objID = 0;
    P2 center{0, 0};
    float width_ = 0, height_ = 0;
    int bb_w = 0, bb_h = 0;
    std::vector<std::pair<P2, P2>> inlierPts2D;  // (direction, pixel)
    int maxPixels = 0, effPixels = 0, inliers = 0, refSteps = 0;
    bool operator<(const TransHyp& o) const { return (float)inliers > (float)o.inliers; }
};

struct Rng {  // thread_rand.cpp:40-67
    std::vector<std::mt19937> gens;
    explicit Rng(int nthreads, unsigned seed = 1305)
    {
        for (int i = 0; i < nthreads; i++) {
            gens.emplace_back();
            gens[i].seed(i + seed);
        }
    }
    int irand(int incMin, int excMax)
    {
        std::uniform_int_distribution<int> dist(incMin, excMax - 1);
        return dist(gens[omp_get_thread_num() % gens.size()]);
    }
};

// Hypothesis::calcCenter: least squares of {m_i x + n_i y = m_i a_i + n_i b_i}, m = -dir.y, n = dir.x
P2 calc_center(const std::vector<std::pair<P2, P2>>& pts)
{
    double a11 = 0, a12 = 0, a22 = 0, b1 = 0, b2 = 0;
    for (auto& p : pts) {
        double m = -p.first.y, n = p.first.x;
        double r = m * p.second.x + n * p.second.y;
        a11 += m * m; a12 += m * n; a22 += n * n; b1 += m * r; b2 += n * r;
    }
    double det = a11 * a22 - a12 * a12;
    double tr = a11 + a22;
    if (std::fabs(det) <= 1e-12 * tr * tr) {
        // rank deficient: minimum-norm solution, as the SVD solve returns
        if (tr <= 0) return P2{0, 0};
        // A^T A = lambda v v^T  ->  x = v (v . b) / lambda
        double vx = a11 >= a22 ? a11 : a12, vy = a11 >= a22 ? a12 : a22;
        double nv = std::sqrt(vx * vx + vy * vy);
        if (nv == 0) return P2{0, 0};
        vx /= nv; vy /= nv;
        double k = (vx * b1 + vy * b2) / tr;
        return P2{vx * k, vy * k};
    }
    return P2{(b1 * a22 - b2 * a12) / det, (a11 * b2 - a12 * b1) / det};
}

inline float point2line(P2 x, float nx, float ny, float px, float py)
{
    float n1 = -ny, n2 = nx;
    float x1 = (float)x.x, x2 = (float)x.y;
    return std::fabs(n1 * (x1 - px) + n2 * (x2 - py)) / std::sqrt(n1 * n1 + n2 * n2);
}

struct Ctx {
    const int* labelmap;
    const float* vertmap;
    int H, W, C;
    std::vector<std::vector<int>> labels;
};

inline void mode2d(const Ctx& c, int objID, int x, int y, float& u, float& v, float& dist)
{
    size_t off = (size_t)3 * objID + (size_t)3 * c.C * ((size_t)y * c.W + x);
    u = c.vertmap[off]; v = c.vertmap[off + 1]; dist = c.vertmap[off + 2];
}

inline bool is_inlier(const Ctx& c, const TransHyp& h, int px, int py, float u, float v, float thr)
{
    double ddx = h.center.x - px, ddy = h.center.y - py;
    float d = (float)std::sqrt(ddx * ddx + ddy * ddy);
    float ang = u * ((float)h.center.x - (float)px) + v * ((float)h.center.y - (float)py);
    return point2line(h.center, u, v, (float)px, (float)py) < thr && ang > 0 && d < (float)std::max(h.bb_w, h.bb_h);
}

void count_inliers(const Ctx& c, TransHyp& hyp, float thr, int pixelBatch)
{
    hyp.inlierPts2D.clear();
    hyp.inliers = 0;
    hyp.effPixels = 0;
    hyp.maxPixels += pixelBatch;
    const std::vector<int>& L = c.labels[hyp.objID];
    int maxPt = (int)L.size();
    float successRate = hyp.maxPixels / (float)maxPt;
    std::mt19937 generator;
    std::negative_binomial_distribution<int> distribution(1, successRate);
    for (unsigned ptIdx = 0; ptIdx < (unsigned)maxPt;) {
        int index = L[ptIdx];
        int x = index % c.W, y = index / c.W;
        hyp.effPixels++;
        float u, v, dist;
        mode2d(c, hyp.objID, x, y, u, v, dist);
        if (is_inlier(c, hyp, x, y, u, v, thr)) {
            hyp.inlierPts2D.push_back({P2{u, v}, P2{(double)x, (double)y}});
            hyp.inliers++;
        }
        if (successRate < 1) ptIdx += std::max(1, distribution(generator));
        else ptIdx++;
    }
}

std::vector<TransHyp*> working_queue(std::map<int, std::vector<TransHyp>>& hm, int maxIt, int is_train)
{
    std::vector<TransHyp*> q;
    for (auto& kv : hm)
        for (auto& h : kv.second)
            if (is_train ? (h.refSteps < maxIt) : (kv.second.size() > 1 || h.refSteps < maxIt)) q.push_back(&h);
    return q;
}

void estimate_center(const Ctx& c0, const float* extents, int batch, int is_train, float fx, float fy, float px, float py,
                     int nthreads, Rng& rng, std::vector<std::vector<float>>& outputs)
{
    Ctx c = c0;
    const int W = c.W, H = c.H, C = c.C;
    const float minArea = 400, minDist2D = 10, inlierThreshold = 0.5f;
    const int ransacIterations = 256, preemptiveBatch = 100, maxPixels = 1000, refIt = is_train ? 4 : 8;
    const long maxIterations = 10000000;
    // getLabels: column-major scan (x outer, y inner), hough_voting_op.cc:293-299
    c.labels.assign(C, {});
    for (int x = 0; x < W; x++)
        for (int y = 0; y < H; y++) {
            int l = c.labelmap[y * W + x];
            if (l >= 0 && l < C) c.labels[l].push_back(y * W + x);
        }
    std::vector<int> object_ids;
    for (int i = 1; i < C; i++)
        if ((float)c.labels[i].size() > minArea) object_ids.push_back(i);
    if (object_ids.empty()) return;

    std::map<int, std::vector<TransHyp>> hypMap;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic) if (nthreads > 1)
    for (int h = 0; h < ransacIterations; h++)
        for (long i = 0; i < maxIterations; i++) {
            int objID = object_ids[rng.irand(0, (int)object_ids.size())];
            if (objID == 0) continue;
            const std::vector<int>& L = c.labels[objID];
            int index = L[rng.irand(0, (int)L.size())];
            float p1x = (float)(index % W), p1y = (float)(index / W);
            float u1, v1, d1;
            mode2d(c, objID, (int)p1x, (int)p1y, u1, v1, d1);
            if (d1 < 0) continue;  // samplePoint2D, :338-352
            index = L[rng.irand(0, (int)L.size())];
            float p2x = (float)(index % W), p2y = (float)(index / W);
            {
                double ex = (double)(p1x - p2x), ey = (double)(p1y - p2y);
                if (std::sqrt(ex * ex + ey * ey) < minDist2D) continue;
            }
            float u2, v2, d2;
            mode2d(c, objID, (int)p2x, (int)p2y, u2, v2, d2);
            if (d2 < 0) continue;
            std::vector<std::pair<P2, P2>> pts = {{P2{u1, v1}, P2{p1x, p1y}}, {P2{u2, v2}, P2{p2x, p2y}}};
            float distance = (d1 + d2) / 2;
            P2 center = calc_center(pts);
            int cx = (int)center.x, cy = (int)center.y;
            if (C > 2 && cx >= 0 && cx < W && cy >= 0 && cy < H && c.labelmap[cy * W + cx] == 0) continue;
            TransHyp hyp;
            hyp.objID = objID;
            hyp.center = center;
            // projectPoints of the extent box at depth `distance` (rvec = 0), :621-642
            int minX = 10000000, maxX = -10000000, minY = 10000000, maxY = -10000000;
            float xh = extents[objID * 3] * 0.5f, yh = extents[objID * 3 + 1] * 0.5f, zh = extents[objID * 3 + 2] * 0.5f;
            for (int k = 0; k < 8; k++) {
                double X = (k & 1) ? -xh : xh, Y = (k & 2) ? -yh : yh, Z = ((k & 4) ? -zh : zh) + (double)distance;
                float bx = (float)((double)fx * (X / Z) + (double)px), by = (float)((double)fy * (Y / Z) + (double)py);
                minX = (int)std::min((float)minX, bx); minY = (int)std::min((float)minY, by);
                maxX = (int)std::max((float)maxX, bx); maxY = (int)std::max((float)maxY, by);
            }
            hyp.bb_w = maxX - minX + 1; hyp.bb_h = maxY - minY + 1;
            float cxf = (float)center.x, cyf = (float)center.y;
            double n1 = std::sqrt((double)(p1x - cxf) * (p1x - cxf) + (double)(p1y - cyf) * (p1y - cyf));
            double n2 = std::sqrt((double)(p2x - cxf) * (p2x - cxf) + (double)(p2y - cyf) * (p2y - cyf));
            int mx = std::max(hyp.bb_w, hyp.bb_h);
            if (n1 > mx || n2 > mx) continue;
#pragma omp critical
            hypMap[objID].push_back(hyp);
            break;
        }

    std::vector<int> objList;
    for (auto& kv : hypMap) objList.push_back(kv.first);
    auto queue = working_queue(hypMap, refIt, is_train);
    while (!queue.empty()) {
#pragma omp parallel for num_threads(nthreads) schedule(dynamic) if (nthreads > 1)
        for (int h = 0; h < (int)queue.size(); h++) count_inliers(c, *queue[h], inlierThreshold, preemptiveBatch);
        for (int o : objList) {
            auto& v = hypMap[o];
            if (v.size() > 1) {
                std::sort(v.begin(), v.end());
                v.erase(v.begin() + v.size() / 2, v.end());
            }
        }
        queue = working_queue(hypMap, refIt, is_train);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic) if (nthreads > 1)
        for (int h = 0; h < (int)queue.size(); h++) {
            TransHyp& hyp = *queue[h];
            if (hyp.inlierPts2D.size() >= 4) {  // updateHyp2D
                if ((int)hyp.inlierPts2D.size() >= maxPixels) {  // filterInliers2D
                    std::vector<std::pair<P2, P2>> keep;
                    for (int k = 0; k < maxPixels; k++) keep.push_back(hyp.inlierPts2D[rng.irand(0, (int)hyp.inlierPts2D.size())]);
                    hyp.inlierPts2D = keep;
                }
                hyp.center = calc_center(hyp.inlierPts2D);
            }
            hyp.refSteps++;
        }
        queue = working_queue(hypMap, refIt, is_train);
    }

    for (auto& kv : hypMap)
        for (auto& hyp : kv.second) {
            std::vector<float> roi(13, 0.f);
            roi[0] = (float)batch;
            roi[1] = (float)hyp.objID;
            P2 center = hyp.center;
            float rx = (float)((center.x - px) / fx), ry = (float)((center.y - py) / fy);
            // TransHyp::compute_distance: mean of the third vertex channel over the stored inliers (no exp)
            float distance = 0;
            for (int i = 0; i < hyp.inliers && i < (int)hyp.inlierPts2D.size(); i++) {
                int x = (int)hyp.inlierPts2D[i].second.x, y = (int)hyp.inlierPts2D[i].second.y;
                distance += c.vertmap[(size_t)3 * hyp.objID + (size_t)3 * C * ((size_t)y * W + x) + 2];
            }
            distance = distance / hyp.inliers;
            // compute_width_height :451-480 over ALL pixels of the class
            float w = -1, hgt = -1;
            for (int index : c.labels[hyp.objID]) {
                int x = index % W, y = index / W;
                float u, v, dd;
                mode2d(c, hyp.objID, x, y, u, v, dd);
                if (is_inlier(c, hyp, x, y, u, v, inlierThreshold)) {
                    float ax = (float)std::fabs(x - center.x), ay = (float)std::fabs(y - center.y);
                    if (ax > w) w = ax;
                    if (ay > hgt) hgt = ay;
                }
            }
            hyp.width_ = 2 * w; hyp.height_ = 2 * hgt;
            float scale = 0.05f;
            roi[2] = (float)(center.x - hyp.width_ * (0.5 + scale));
            roi[3] = (float)(center.y - hyp.height_ * (0.5 + scale));
            roi[4] = (float)(center.x + hyp.width_ * (0.5 + scale));
            roi[5] = (float)(center.y + hyp.height_ * (0.5 + scale));
            roi[6] = 1; roi[7] = 0; roi[8] = 0; roi[9] = 0;  // rvec = 0 -> identity quaternion
            roi[10] = rx * distance; roi[11] = ry * distance; roi[12] = distance;
            outputs.push_back(roi);
            if (is_train) {
                float x1 = roi[2], y1 = roi[3], ww = roi[4] - roi[2], hh = roi[5] - roi[3];
                const int jx[8] = {-1, 1, -1, 1, 0, -1, 0, 1}, jy[8] = {-1, -1, 1, 1, -1, 0, 1, 0};
                for (int j = 0; j < 8; j++) {
                    roi[2] = jx[j] == 0 ? x1 : (float)(x1 + jx[j] * (0.05 * ww));
                    roi[3] = jy[j] == 0 ? y1 : (float)(y1 + jy[j] * (0.05 * hh));
                    roi[4] = roi[2] + ww; roi[5] = roi[3] + hh;
                    outputs.push_back(roi);
                }
            }
        }
}

}  // namespace

// HoughvotingOp::Compute.  top_box [cap,6], top_pose [cap,7]; returns the row count (>= 1: dummy row
// with cls = -1 when nothing was detected, hough_voting_op.cc:208-222), or -1 if cap is too small.
extern "C" int cpu_hough_voting(const int* label, const float* vertex, const float* extents, const float* meta, int B,
                                int H, int W, int C, int num_meta, int is_train, int nthreads, float* top_box,
                                float* top_pose, int cap)
{
    if (nthreads < 1) nthreads = 1;
    Rng rng(nthreads);
    std::vector<std::vector<float>> outputs;
    for (int n = 0; n < B; n++) {
        Ctx c{label + (size_t)n * H * W, vertex + (size_t)n * H * W * 3 * C, H, W, C, {}};
        const float* m = meta + (size_t)n * num_meta;
        estimate_center(c, extents, n, is_train, m[0], m[4], m[2], m[5], nthreads, rng, outputs);
    }
    if (outputs.empty()) {
        std::vector<float> roi(13, 0.f);
        roi[1] = -1; roi[4] = 1; roi[5] = 1; roi[6] = 1;
        outputs.push_back(roi);
    }
    if ((int)outputs.size() > cap) return -1;
    for (size_t i = 0; i < outputs.size(); i++) {
        for (int k = 0; k < 6; k++) top_box[i * 6 + k] = outputs[i][k];
        for (int k = 0; k < 7; k++) top_pose[i * 7 + k] = outputs[i][6 + k];
    }
    return (int)outputs.size();
}
