// f2: the evaluation path of the reference (Detect_OBB.py:456-648) - detection x ground-truth matching.
//
//   _match_dets_to_gts_pixel (:456-480)  per image: detections in list order, each takes the unused
//                                        same-class GT of largest IoU (first one on ties, IoU must be > 0)
//                                        and is a TP iff that IoU >= iou_thr.
//   compute_pr_for_class     (:512-565)  the same rule per class with the detections of all images in
//                                        stable score-descending order, GT "matched" flags per image.
//   evaluate_map             (:574-607)  compute_pr_for_class for 10 IoU thresholds x every class: the
//                                        IoU values do not depend on the threshold, so they are computed
//                                        once and the matching is replayed per threshold.
//   evaluate_center_hit      (:609-648)  a detection is a TP iff its centre lies strictly inside the first
//                                        unused valid GT polygon of its class (shapely contains()).
//
// A "segment" is one (image [, class]) group: detections [det_off[s], det_off[s+1]) in processing
// order against GTs [gt_off[s], gt_off[s+1]).  IoU is float64 (the reference compares Python floats);
// GT quads are general (labels need not be rectangles), so the float64 clip of geom.cuh is used.
// The sequential part (one detection after the other, because of the used/matched flags) is one CTA
// per (segment, threshold): the CTA walks the detections, its threads scan the GTs and reduce to the
// best candidate.
#include "gm_common.cuh"
#include "geom.cuh"

namespace {

constexpr int EV_THREADS = 256;

__device__ __forceinline__ int seg_of(const long long* __restrict__ off, int ns, long long p) {
    int lo = 0, hi = ns;                 // off[lo] <= p < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// iou[mat_off[s] + i * ng_s + j] = IoU(det i, gt j) of segment s; -1 where the classes differ (never a candidate).
__global__ void __launch_bounds__(EV_THREADS)
k_eval_iou(const double* __restrict__ det, const int* __restrict__ det_cls, const double* __restrict__ gt,
           const int* __restrict__ gt_cls, const long long* __restrict__ det_off, const long long* __restrict__ gt_off,
           const long long* __restrict__ mat_off, int ns, double* __restrict__ iou) {
    const long long p = (long long)blockIdx.x * EV_THREADS + threadIdx.x;
    if (p >= mat_off[ns]) return;
    const int s = seg_of(mat_off, ns, p);
    const long long ng = gt_off[s + 1] - gt_off[s];
    const long long q = p - mat_off[s];
    const long long i = det_off[s] + q / ng, j = gt_off[s] + q % ng;
    double v = -1.0;
    if (!det_cls || !gt_cls || det_cls[i] == gt_cls[j]) v = iou_f64_from_corners(det + i * 8, gt + j * 8);
    iou[p] = v;
}

// (value desc, index asc) arg-max over a CTA; returns the winner in every thread
__device__ __forceinline__ void block_best(double& v, int& j, double* sv, int* sj) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, d);
        const int oj = __shfl_xor_sync(0xffffffffu, j, d);
        if (ov > v || (ov == v && oj < j)) { v = ov; j = oj; }
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sv[warp] = v; sj[warp] = j; }
    __syncthreads();
    if (warp == 0) {
        v = (threadIdx.x < EV_THREADS / 32) ? sv[threadIdx.x] : -2.0;
        j = (threadIdx.x < EV_THREADS / 32) ? sj[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, d);
            const int oj = __shfl_xor_sync(0xffffffffu, j, d);
            if (ov > v || (ov == v && oj < j)) { v = ov; j = oj; }
        }
        if (threadIdx.x == 0) { sv[0] = v; sj[0] = j; }
    }
    __syncthreads();
    v = sv[0]; j = sj[0];
    __syncthreads();
}

// match[t * n_det + i] = segment-local index of the GT detection i takes at threshold thr[t], or -1.
// used: scratch uint8 [n_thr][n_gt].
__global__ void __launch_bounds__(EV_THREADS)
k_eval_match(const double* __restrict__ iou, const long long* __restrict__ det_off, const long long* __restrict__ gt_off,
             const long long* __restrict__ mat_off, const double* __restrict__ thr, long long n_det, long long n_gt,
             unsigned char* __restrict__ used_all, int* __restrict__ match) {
    __shared__ double sv[EV_THREADS / 32];
    __shared__ int sj[EV_THREADS / 32];
    const int s = blockIdx.x, t = blockIdx.y;
    const long long d0 = det_off[s], d1 = det_off[s + 1], g0 = gt_off[s];
    const int ng = (int)(gt_off[s + 1] - g0);
    unsigned char* used = used_all + (long long)t * n_gt + g0;
    for (int j = threadIdx.x; j < ng; j += EV_THREADS) used[j] = 0;
    __syncthreads();
    const double th = thr[t];
    const double* M = iou + mat_off[s];
    for (long long i = d0; i < d1; ++i) {
        const double* row = M + (i - d0) * ng;
        // the reference starts from best_iou = 0.0 and replaces it only on a strictly larger IoU
        double bv = 0.0;
        int bj = 0x7fffffff;
        for (int j = threadIdx.x; j < ng; j += EV_THREADS) {
            const double v = row[j];
            if (!used[j] && v > bv) { bv = v; bj = j; }       // ascending j per thread: first of equal values kept
        }
        block_best(bv, bj, sv, sj);
        const bool hit = bj != 0x7fffffff && bv >= th;
        if (threadIdx.x == 0) {
            match[(long long)t * n_det + i] = hit ? bj : -1;
            if (hit) used[bj] = 1;
        }
        __syncthreads();
    }
}

// shapely Polygon(pts).is_valid / contains(Point) as restated in oracle/geometry.py: valid = simple ring with
// non-zero area (either winding; convex or concave); contains = strictly inside.
__device__ bool quad_contains(const double* __restrict__ q, double px, double py, bool& valid) {
    double s = 0.0;
    bool pos = false, neg = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3, k = (i + 2) & 3;
        s = __dadd_rn(s, __dsub_rn(__dmul_rn(q[2 * i], q[2 * j + 1]), __dmul_rn(q[2 * j], q[2 * i + 1])));
        const double cr = __dsub_rn(__dmul_rn(q[2 * j] - q[2 * i], q[2 * k + 1] - q[2 * j + 1]),
                                    __dmul_rn(q[2 * j + 1] - q[2 * i + 1], q[2 * k] - q[2 * j]));
        pos |= cr > 0.0;
        neg |= cr < 0.0;
    }
    valid = (s != 0.0) && !(pos && neg);
    if (!valid) {
        // not convex: a concave SIMPLE quad is still valid for shapely (Detect_OBB.py:631-634)
        if (s == 0.0) return false;
        GQuad g;
        gquad_from_corners(q, g);
        valid = g.kind == 2;
        return valid && gquad_contains_concave(g, px, py);
    }
    bool inside = true;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        // ring made counter-clockwise by reversal (p[::-1]) when the signed area is negative
        const int a = s >= 0.0 ? e : 3 - e;
        const int b = s >= 0.0 ? ((e + 1) & 3) : ((3 - e + 3) & 3);
        const double ax = q[2 * a], ay = q[2 * a + 1], bx = q[2 * b], by = q[2 * b + 1];
        const double cr = __dsub_rn(__dmul_rn(bx - ax, py - ay), __dmul_rn(by - ay, px - ax));
        inside &= cr > 0.0;
    }
    return inside;
}

// match[i] = segment-local index of the first unused same-class valid GT whose polygon strictly contains the
// centre of detection i, or -1.
__global__ void __launch_bounds__(EV_THREADS)
k_eval_center_hit(const double* __restrict__ det, const int* __restrict__ det_cls, const double* __restrict__ gt,
                  const int* __restrict__ gt_cls, const long long* __restrict__ det_off,
                  const long long* __restrict__ gt_off, unsigned char* __restrict__ used_all, int* __restrict__ match) {
    __shared__ int s_best;
    const int s = blockIdx.x;
    const long long d0 = det_off[s], d1 = det_off[s + 1], g0 = gt_off[s];
    const int ng = (int)(gt_off[s + 1] - g0);
    unsigned char* used = used_all + g0;
    for (int j = threadIdx.x; j < ng; j += EV_THREADS) used[j] = 0;
    __syncthreads();
    for (long long i = d0; i < d1; ++i) {
        if (threadIdx.x == 0) s_best = 0x7fffffff;
        __syncthreads();
        const double* b = det + i * 8;
        // box_center_from_xyxyxyxy (Detect_OBB.py:159-165): (x1 + x2 + x3 + x4) / 4.0
        const double cx = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(b[0], b[2]), b[4]), b[6]), 4.0);
        const double cy = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(b[1], b[3]), b[5]), b[7]), 4.0);
        const int c = det_cls[i];
        int mine = 0x7fffffff;
        for (int j = threadIdx.x; j < ng && j < mine; j += EV_THREADS) {
            if (used[j] || gt_cls[g0 + j] != c) continue;
            bool valid;
            if (quad_contains(gt + (g0 + j) * 8, cx, cy, valid)) mine = j;
        }
        if (mine != 0x7fffffff) atomicMin(&s_best, mine);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int bj = s_best;
            match[i] = bj != 0x7fffffff ? bj : -1;
            if (bj != 0x7fffffff) used[bj] = 1;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int gm_eval_iou_segments(const double* det_dev, const int32_t* det_cls_dev, const double* gt_dev,
                                    const int32_t* gt_cls_dev, const int64_t* det_off_dev, const int64_t* gt_off_dev,
                                    const int64_t* mat_off_dev, int32_t n_segments, int64_t n_pairs,
                                    double* iou_dev, void* stream) {
    if (n_segments < 0 || n_pairs < 0) return GM_EINVAL;
    if (n_segments == 0 || n_pairs == 0) return GM_OK;
    if (!det_dev || !gt_dev || !det_off_dev || !gt_off_dev || !mat_off_dev || !iou_dev) return GM_EINVAL;
    const long long blocks = (n_pairs + EV_THREADS - 1) / EV_THREADS;
    if (blocks > 0x7fffffffLL) return GM_ERANGE;
    k_eval_iou<<<(unsigned)blocks, EV_THREADS, 0, gm_stream(stream)>>>(
        det_dev, det_cls_dev, gt_dev, gt_cls_dev, (const long long*)det_off_dev, (const long long*)gt_off_dev,
        (const long long*)mat_off_dev, n_segments, iou_dev);
    gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_eval_match_greedy(const double* iou_dev, const int64_t* det_off_dev, const int64_t* gt_off_dev,
                                    const int64_t* mat_off_dev, int32_t n_segments, int64_t n_det, int64_t n_gt,
                                    const double* thr_dev, int32_t n_thr, uint8_t* used_scratch_dev,
                                    int32_t* match_dev, void* stream) {
    if (n_segments < 0 || n_thr < 0 || n_det < 0 || n_gt < 0) return GM_EINVAL;
    if (n_segments == 0 || n_thr == 0 || n_det == 0) return GM_OK;
    if (!det_off_dev || !gt_off_dev || !mat_off_dev || !thr_dev || !match_dev) return GM_EINVAL;
    if (n_gt > 0 && (!iou_dev || !used_scratch_dev)) return GM_EINVAL;
    if (n_thr > 65535) return GM_ERANGE;
    dim3 grid((unsigned)n_segments, (unsigned)n_thr);
    k_eval_match<<<grid, EV_THREADS, 0, gm_stream(stream)>>>(
        iou_dev, (const long long*)det_off_dev, (const long long*)gt_off_dev, (const long long*)mat_off_dev, thr_dev,
        n_det, n_gt, used_scratch_dev, match_dev);
    gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_eval_center_hit(const double* det_dev, const int32_t* det_cls_dev, const double* gt_dev,
                                  const int32_t* gt_cls_dev, const int64_t* det_off_dev, const int64_t* gt_off_dev,
                                  int32_t n_segments, uint8_t* used_scratch_dev, int32_t* match_dev, void* stream) {
    if (n_segments < 0) return GM_EINVAL;
    if (n_segments == 0) return GM_OK;
    if (!det_dev || !det_cls_dev || !gt_dev || !gt_cls_dev || !det_off_dev || !gt_off_dev || !used_scratch_dev || !match_dev)
        return GM_EINVAL;
    k_eval_center_hit<<<(unsigned)n_segments, EV_THREADS, 0, gm_stream(stream)>>>(
        det_dev, det_cls_dev, gt_dev, gt_cls_dev, (const long long*)det_off_dev, (const long long*)gt_off_dev,
        used_scratch_dev, match_dev);
    gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
