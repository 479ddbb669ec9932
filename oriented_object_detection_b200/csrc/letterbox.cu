// f1: the Ultralytics predictor's pre-processing on the device, for a batch of equally sized tiles.
//
// Reference call site: run_inference_on_crop -> model(net_input, conf=...) (Detect_OBB.py:76-85); the
// arithmetic is ultralytics 8.3.196 LetterBox(imgsz, auto=True, stride=32, scaleup=True) followed by
// BasePredictor.preprocess (BGR->RGB for 3 channels, HWC->CHW, float32 / 255), restated in
// oracle/letterbox.py.  The reference runs it once per tile on the host; here
// the tiles of one shape are letterboxed straight out of the packed tile batch into the float32
// [n, C, out_h, out_w] tensor the network consumes - one launch, no host round trip.
//
// Ragged tiles are resized with cv2's 8-bit INTER_LINEAR arithmetic, bit for bit: 11-bit fixed-point
// coefficient pairs (rounded half to even from float32 fractions), a horizontal pass at scale 2^11 and
// the vertical blend (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// HBM bound: C bytes read and 4 C bytes written per output pixel.
#include <cmath>
#include "gm_common.cuh"

namespace {

constexpr int LB_THREADS = 256;
constexpr int LB_ROWS = 4;             // output rows per CTA
constexpr int LB_MAXW = GM_MAX_TILE;   // widest letterboxed row

struct LbGeom {
    int tile_h, tile_w, new_h, new_w, top, left, out_h, out_w, resize;
    double scale_x, scale_y;
};

// cv2 resize: source index and the (1 - f, f) * 2048 coefficient pair of destination index d
__device__ __forceinline__ void lin_coef(int d, double scale, int src, bool clamp_f, int& s0, int& s1, int& c0, int& c1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_f) {                     // horizontal tables clamp the fraction with the index (HResize xmin / xmax)
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    c0 = (int)rintf(__fmul_rn(1.f - f, 2048.f));
    c1 = (int)rintf(__fmul_rn(f, 2048.f));
    s0 = min(max(s, 0), src - 1);
    s1 = min(max(s + 1, 0), src - 1);
}

template <int C>
__global__ void __launch_bounds__(LB_THREADS)
k_letterbox(const uint8_t* __restrict__ packed, const gm_tile* __restrict__ tiles, const LbGeom g,
            float* __restrict__ out) {
    __shared__ int xs0[LB_MAXW], xs1[LB_MAXW];
    __shared__ short xa0[LB_MAXW], xa1[LB_MAXW];
    const gm_tile t = tiles[blockIdx.x];
    const uint8_t* src = packed + (long long)C * t.px_off;
    const int y_first = blockIdx.y * LB_ROWS;
    if (g.resize) {
        for (int x = threadIdx.x; x < g.new_w; x += LB_THREADS) {
            int s0, s1, c0, c1;
            lin_coef(x, g.scale_x, g.tile_w, true, s0, s1, c0, c1);
            xs0[x] = s0; xs1[x] = s1; xa0[x] = (short)c0; xa1[x] = (short)c1;
        }
        __syncthreads();
    }
    const long long plane = (long long)g.out_h * g.out_w;
    float* dst = out + (long long)blockIdx.x * C * plane;
    for (int i = threadIdx.x; i < LB_ROWS * g.out_w; i += LB_THREADS) {
        const int y = y_first + i / g.out_w;
        const int x = i - (i / g.out_w) * g.out_w;
        if (y >= g.out_h) break;
        const int dy = y - g.top, dx = x - g.left;
        unsigned int v[C];
        if (dy < 0 || dy >= g.new_h || dx < 0 || dx >= g.new_w) {
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = 114u;
        } else if (!g.resize) {
            const uint8_t* p = src + ((long long)dy * g.tile_w + dx) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = p[c];
        } else {
            int sy0, sy1, b0, b1;
            lin_coef(dy, g.scale_y, g.tile_h, false, sy0, sy1, b0, b1);
            const int a0 = xa0[dx], a1 = xa1[dx];
            const uint8_t* p00 = src + ((long long)sy0 * g.tile_w + xs0[dx]) * C;
            const uint8_t* p01 = src + ((long long)sy0 * g.tile_w + xs1[dx]) * C;
            const uint8_t* p10 = src + ((long long)sy1 * g.tile_w + xs0[dx]) * C;
            const uint8_t* p11 = src + ((long long)sy1 * g.tile_w + xs1[dx]) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int r0 = (int)p00[c] * a0 + (int)p01[c] * a1;
                const int r1 = (int)p10[c] * a0 + (int)p11[c] * a1;
                const int o = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
                v[c] = (unsigned int)min(max(o, 0), 255);
            }
        }
        const long long o = (long long)y * g.out_w + x;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int cc = (C == 3) ? 2 - c : c;            // BGR -> RGB only for 3 channels
            dst[(long long)cc * plane + o] = __fdiv_rn((float)v[c], 255.f);
        }
    }
}

// Python round(): half to even, like rint() in the default rounding mode
int py_round(double v) { return (int)nearbyint(v); }

}  // namespace

extern "C" int gm_letterbox_shape(int32_t tile_h, int32_t tile_w, int32_t net_size, int32_t stride, int32_t auto_rect,
                                  int32_t* new_h, int32_t* new_w, int32_t* top, int32_t* left,
                                  int32_t* out_h, int32_t* out_w) {
    if (tile_h <= 0 || tile_w <= 0 || net_size <= 0 || stride <= 0) return GM_EINVAL;
    const double r = fmin((double)net_size / tile_h, (double)net_size / tile_w);
    const int nw = py_round(tile_w * r), nh = py_round(tile_h * r);
    int dwi = net_size - nw, dhi = net_size - nh;
    if (auto_rect) { dwi = ((dwi % stride) + stride) % stride; dhi = ((dhi % stride) + stride) % stride; }
    const double dw = dwi / 2.0, dh = dhi / 2.0;
    const int t = py_round(dh - 0.1), b = py_round(dh + 0.1), l = py_round(dw - 0.1), rr = py_round(dw + 0.1);
    if (new_h) *new_h = nh;
    if (new_w) *new_w = nw;
    if (top) *top = t;
    if (left) *left = l;
    if (out_h) *out_h = nh + t + b;
    if (out_w) *out_w = nw + l + rr;
    return GM_OK;
}

extern "C" int gm_letterbox_tiles(const uint8_t* packed_dev, int32_t channels, const gm_tile* tiles_dev, int32_t n_tiles,
                                  int32_t tile_h, int32_t tile_w, int32_t net_size, int32_t stride, int32_t auto_rect,
                                  float* out_dev, void* stream) {
    if (n_tiles < 0 || (channels != 3 && channels != 4)) return GM_EINVAL;
    if (n_tiles == 0) return GM_OK;
    if (!packed_dev || !tiles_dev || !out_dev) return GM_EINVAL;
    LbGeom g;
    int st = gm_letterbox_shape(tile_h, tile_w, net_size, stride, auto_rect, &g.new_h, &g.new_w, &g.top, &g.left,
                                &g.out_h, &g.out_w);
    if (st != GM_OK) return st;
    if (g.out_w > LB_MAXW || g.new_w > LB_MAXW) return GM_ERANGE;
    g.tile_h = tile_h; g.tile_w = tile_w;
    g.resize = (g.new_h != tile_h || g.new_w != tile_w) ? 1 : 0;
    g.scale_x = (double)tile_w / g.new_w;
    g.scale_y = (double)tile_h / g.new_h;
    dim3 grid((unsigned)n_tiles, (unsigned)((g.out_h + LB_ROWS - 1) / LB_ROWS));
    if (channels == 3) k_letterbox<3><<<grid, LB_THREADS, 0, gm_stream(stream)>>>(packed_dev, tiles_dev, g, out_dev);
    else k_letterbox<4><<<grid, LB_THREADS, 0, gm_stream(stream)>>>(packed_dev, tiles_dev, g, out_dev);
    gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
