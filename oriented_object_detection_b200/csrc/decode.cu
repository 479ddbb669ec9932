// a5: post-network decode of a batch of tiles (what `model(tile, conf)` does after the CNN).
//
// Reference call site: run_inference_on_crop / detect_symbols read `results[0].obb[i]
// .xyxyxyxy/.cls/.conf` (Detect_OBB.py:81-83, 228-231).  The arithmetic is Ultralytics
// 8.3.196 (requirements.txt:3; not in the reference tree, not installable offline): restated
// from its published behaviour in SURVEY.md Appendix B and oracle/decode.py - confidence
// filter on the best class, confidence-descending order, class-wise probiou "fast NMS"
// (a box is dropped if ANY earlier box of its class has probiou >= thr, suppressed or not),
// first max_det, regularize_rboxes, scale_boxes (undo the letterbox), xywhr2xyxyxyxy.
// PARITY UNPINNED (no reference test or runnable upstream package); checked against the
// numpy restatement only.
//
// Letterbox convention: every tile is resized by gain = min(S/h, S/w) and centred in an
// S x S network input (fixed-shape batches; Ultralytics' rect/auto mode would crop the
// padding to a multiple of 32 instead - same gain, different pad).
//
// One CTA per tile.  Candidates are ranked by counting in shared memory; the probiou scan is
// one thread per candidate walking the earlier ones (broadcast loads).
#include "gm_common.cuh"

namespace {

constexpr int DEC_THREADS = 256;
constexpr int DEC_FIELDS = 12;     // sorted: cx, cy, w, h, theta, A, B, C, conf, cls; unsorted: conf, cls

__device__ __forceinline__ unsigned int enc_desc(float f) {
    const unsigned int b = __float_as_uint(f);
    const unsigned int e = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~e;
}

__global__ void __launch_bounds__(DEC_THREADS)
k_decode(const float* __restrict__ head, int n_classes, int A, const gm_tile* __restrict__ tiles, int net_h, int net_w,
         float conf_thr, float iou_thr, int max_det, float* __restrict__ ws,
         float* __restrict__ out_boxes, int* __restrict__ out_cls, float* __restrict__ out_conf,
         int* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);       // A entries
    __shared__ int s_count, s_out;
    const int t = blockIdx.x;
    const gm_tile tl = tiles[t];
    const int C = 4 + n_classes + 1;
    const float* hd = head + (long long)t * C * A;
    float* cand = ws + (long long)t * DEC_FIELDS * A;           // SoA: field f at cand[f*A + k]
    unsigned char* dead = reinterpret_cast<unsigned char*>(keys + A);                 // A entries
    const int tid = threadIdx.x;
    if (tid == 0) { s_count = 0; s_out = 0; }
    __syncthreads();

    // 1. confidence filter on the best class
    for (int a = tid; a < A; a += DEC_THREADS) {
        float best = hd[(long long)4 * A + a];
        int bc = 0;
        for (int c = 1; c < n_classes; ++c) {
            const float v = hd[(long long)(4 + c) * A + a];
            if (v > best) { best = v; bc = c; }
        }
        if (best > conf_thr) {
            const int k = atomicAdd(&s_count, 1);
            // rank key: confidence descending, then anchor index ascending
            keys[k] = ((unsigned long long)enc_desc(best) << 32) | (unsigned int)a;
            cand[10 * A + k] = best;                // unsorted slot
            cand[11 * A + k] = __int_as_float(bc);
        }
    }
    __syncthreads();
    const int K = s_count;
    // 2. rank by counting, scatter the candidate fields in sorted order
    for (int k = tid; k < K; k += DEC_THREADS) {
        const unsigned long long me = keys[k];
        int r = 0;
        for (int j = 0; j < K; ++j) r += (keys[j] < me) ? 1 : 0;
        const int a = (int)(me & 0xffffffffull);
        const float cx = hd[a], cy = hd[(long long)A + a];
        const float w = hd[(long long)2 * A + a], h = hd[(long long)3 * A + a];
        const float th = hd[(long long)(4 + n_classes) * A + a];
        // _get_covariance_matrix: a = w^2/12, b = h^2/12
        const float ga = w * w / 12.f, gb = h * h / 12.f;
        const float c = cosf(th), s = sinf(th);
        const float c2 = c * c, s2 = s * s;
        cand[0 * A + r] = cx; cand[1 * A + r] = cy; cand[2 * A + r] = w; cand[3 * A + r] = h;
        cand[4 * A + r] = th;
        cand[5 * A + r] = ga * c2 + gb * s2;
        cand[6 * A + r] = ga * s2 + gb * c2;
        cand[7 * A + r] = (ga - gb) * c * s;
        cand[8 * A + r] = cand[10 * A + k];
        cand[9 * A + r] = cand[11 * A + k];
    }
    __syncthreads();
    // 3. class-wise probiou fast-NMS: dead[j] iff some i < j of the same class reaches the threshold
    const float eps = 1e-7f;
    for (int j = tid; j < K; j += DEC_THREADS) {
        const float x2 = cand[j], y2 = cand[A + j];
        const float a2 = cand[5 * A + j], b2 = cand[6 * A + j], c2 = cand[7 * A + j];
        const int cls2 = __float_as_int(cand[9 * A + j]);
        const float det2 = fmaxf(a2 * b2 - c2 * c2, 0.f);
        unsigned char d = 0;
        for (int i = 0; i < j; ++i) {
            if (__float_as_int(cand[9 * A + i]) != cls2) continue;
            const float x1 = cand[i], y1 = cand[A + i];
            const float a1 = cand[5 * A + i], b1 = cand[6 * A + i], c1 = cand[7 * A + i];
            const float sa = a1 + a2, sb = b1 + b2, sc = c1 + c2;
            const float den = sa * sb - sc * sc;
            const float dx = x1 - x2, dy = y1 - y2;
            const float t1 = ((sa * dy * dy + sb * dx * dx) / (den + eps)) * 0.25f;
            const float t2 = ((sc * (x2 - x1) * (y1 - y2)) / (den + eps)) * 0.5f;
            const float det1 = fmaxf(a1 * b1 - c1 * c1, 0.f);
            const float t3 = logf(den / (4.f * sqrtf(det1 * det2) + eps) + eps) * 0.5f;
            const float bd = fminf(fmaxf(t1 + t2 + t3, eps), 100.f);
            const float hd2 = sqrtf(1.f - expf(-bd) + eps);
            if (1.f - hd2 >= iou_thr) { d = 1; break; }
        }
        dead[j] = d;
    }
    __syncthreads();
    // 4. first max_det survivors, in order: serial prefix over K by one warp (K is small)
    if (tid < 32) {
        int base = 0;
        for (int j0 = 0; j0 < K && base < max_det; j0 += 32) {
            const int j = j0 + tid;
            const bool live = (j < K) && !dead[j];
            const unsigned int m = __ballot_sync(0xffffffffu, live);
            const int slot = base + __popc(m & ((1u << tid) - 1u));
            if (live && slot < max_det) {
                // 5. regularize_rboxes
                const float PI = 3.14159265358979323846f;
                float w = cand[2 * A + j], h = cand[3 * A + j], th = cand[4 * A + j];
                float tm = fmodf(th, PI); if (tm < 0.f) tm += PI;
                const bool swap = tm >= PI / 2.f;
                const float w_ = swap ? h : w, h_ = swap ? w : h;
                float tr = fmodf(tm, PI / 2.f);
                // 6. scale_boxes(xywh=True): remove the letterbox pad, divide by the gain
                //    (gain and pad are Python floats / round() upstream: float64, half-to-even)
                const double gd = fmin((double)net_h / (double)tl.h, (double)net_w / (double)tl.w);
                const float gain = (float)gd;
                const float padx = (float)rint(((double)net_w - (double)tl.w * gd) / 2.0 - 0.1);
                const float pady = (float)rint(((double)net_h - (double)tl.h * gd) / 2.0 - 0.1);
                const float cx = (cand[j] - padx) / gain, cy = (cand[A + j] - pady) / gain;
                const float bw = w_ / gain, bh = h_ / gain;
                // 7. xywhr2xyxyxyxy
                const float c = cosf(tr), s = sinf(tr);
                const float v1x = bw / 2.f * c, v1y = bw / 2.f * s;
                const float v2x = -bh / 2.f * s, v2y = bh / 2.f * c;
                const long long o = (long long)t * max_det + slot;
                float* ob = out_boxes + o * 8;
                ob[0] = cx + v1x + v2x; ob[1] = cy + v1y + v2y;
                ob[2] = cx + v1x - v2x; ob[3] = cy + v1y - v2y;
                ob[4] = cx - v1x - v2x; ob[5] = cy - v1y - v2y;
                ob[6] = cx - v1x + v2x; ob[7] = cy - v1y + v2y;
                out_cls[o] = __float_as_int(cand[9 * A + j]);
                out_conf[o] = cand[8 * A + j];
            }
            base += __popc(m);
        }
        if (tid == 0) out_count[t] = min(base, max_det);
    }
}

}  // namespace

extern "C" size_t gm_decode_workspace_bytes(int32_t n_tiles, int32_t n_anchors) {
    if (n_tiles < 0 || n_anchors < 0) return 0;
    return gm_align_up((size_t)n_tiles * DEC_FIELDS * (size_t)n_anchors * sizeof(float), 256) + 256;
}

extern "C" int gm_decode_tiles(const float* head_dev, int32_t n_tiles, int32_t n_classes, int32_t n_anchors,
                               const gm_tile* tiles_dev, int32_t net_h, int32_t net_w,
                               float conf_thr, float iou_probiou, int32_t max_det,
                               float* boxes_local_dev, int32_t* cls_dev, float* conf_dev, int32_t* count_dev,
                               void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (n_tiles < 0 || n_classes < 1 || n_anchors < 1 || net_h < 1 || net_w < 1 || max_det < 1) return GM_EINVAL;
    if (n_tiles == 0) return GM_OK;
    if (!head_dev || !tiles_dev || !boxes_local_dev || !cls_dev || !conf_dev || !count_dev || !workspace_dev)
        return GM_EINVAL;
    if (n_anchors > 8192) return GM_ERANGE;
    if (workspace_bytes < gm_decode_workspace_bytes(n_tiles, n_anchors)) return GM_ENOSPC;
    const size_t smem = (size_t)n_anchors * (sizeof(unsigned long long) + 1) + 16;
    GM_CUDA_TRY(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_decode<<<(unsigned)n_tiles, DEC_THREADS, smem, gm_stream(stream)>>>(
        head_dev, n_classes, n_anchors, tiles_dev, net_h, net_w, conf_thr, iou_probiou, max_det,
        static_cast<float*>(workspace_dev), boxes_local_dev, cls_dev, conf_dev, count_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
