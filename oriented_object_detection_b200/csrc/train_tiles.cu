// f3: the label side of the reference's training tilers - enumerate_and_save_nonempty_tiles
// (Train_OBB.py:44-146) and crop_images_and_labels (Train_OBB.py:290-428).
//
// Both walk every FULL tile of the overlapped plan (ragged edge tiles are skipped, :90-91 / :336-337,
// unlike detection) and, per tile, filter the image's label table with pandas: a label belongs to a tile
// iff the midpoint of its corners 1 and 4 lies in [x, x+ts) x [y, y+ts) and at least
// `object_boundary_threshold` of its axis-aligned bounding box is inside the tile (_cov_frac); kept labels
// are shifted to the tile, clipped to [0, ts] and divided by ts.  That is O(tiles x labels) pandas work per
// image; here one thread takes one (label, candidate tile) pair - a label's anchor can only fall into the
// ceil(ts / stride)^2 tiles around it - in float64 with the reference's operation order.
#include "gm_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_train_label_tiles(const double* __restrict__ labels, long long n, int H, int W, int ts, int stride, double thr,
                    int span, int cols_full, int rows_full, unsigned char* __restrict__ flag,
                    int* __restrict__ tile_id, double* __restrict__ coords) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = span * span;
    if (p >= n * per) return;
    const long long i = p / per;
    const int a = (int)(p - i * per);
    const double* b = labels + i * 8;
    flag[p] = 0;
    tile_id[p] = -1;
    const double mx = __ddiv_rn(__dadd_rn(b[0], b[6]), 2.0);          // (x1 + x4) / 2
    const double my = __ddiv_rn(__dadd_rn(b[1], b[7]), 2.0);          // (y1 + y4) / 2
    if (!(mx >= 0.0) || !(my >= 0.0)) return;                          // NaN or left/above the image: no tile starts below 0
    const long long kx = (long long)floor(mx / (double)stride) - (a % span);
    const long long ky = (long long)floor(my / (double)stride) - (a / span);
    if (kx < 0 || ky < 0 || kx >= cols_full || ky >= rows_full) return;
    const double x = (double)(kx * stride), y = (double)(ky * stride), t = (double)ts;
    if (!(mx >= x && mx < x + t && my >= y && my < y + t)) return;
    // _cov_frac: share of the label's bounding box that lies inside the tile
    const double bx1 = fmin(fmin(b[0], b[2]), fmin(b[4], b[6])), bx2 = fmax(fmax(b[0], b[2]), fmax(b[4], b[6]));
    const double by1 = fmin(fmin(b[1], b[3]), fmin(b[5], b[7])), by2 = fmax(fmax(b[1], b[3]), fmax(b[5], b[7]));
    const double ax = fmax(0.0, __dsub_rn(fmin(bx2, x + t), fmax(bx1, x)));
    const double ay = fmax(0.0, __dsub_rn(fmin(by2, y + t), fmax(by1, y)));
    const double inter = __dmul_rn(ax, ay);
    const double area = fmax(1e-6, __dmul_rn(__dsub_rn(bx2, bx1), __dsub_rn(by2, by1)));
    if (!(__ddiv_rn(inter, area) >= thr)) return;
    flag[p] = 1;
    tile_id[p] = (int)(ky * cols_full + kx);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double o = (k & 1) ? y : x;
        double v = __dsub_rn(b[k], o);
        v = fmin(fmax(v, 0.0), t);                                     // DataFrame.clip(0, tile_size)
        coords[p * 8 + k] = __ddiv_rn(v, t);
    }
}

}  // namespace

extern "C" int64_t gm_train_tile_grid(int32_t H, int32_t W, int32_t tile_size, int32_t overlap, int32_t* rows, int32_t* cols,
                                      int32_t* span) {
    const int stride = tile_size - overlap;
    if (H <= 0 || W <= 0 || tile_size <= 0 || stride <= 0) return GM_EINVAL;
    const int r = H >= tile_size ? (H - tile_size) / stride + 1 : 0;
    const int c = W >= tile_size ? (W - tile_size) / stride + 1 : 0;
    if (rows) *rows = r;
    if (cols) *cols = c;
    if (span) *span = (tile_size + stride - 1) / stride;
    return (int64_t)r * c;
}

extern "C" int gm_train_label_tiles(const double* labels_dev, int64_t n, int32_t H, int32_t W, int32_t tile_size,
                                    int32_t overlap, double cov_threshold, uint8_t* flag_dev, int32_t* tile_id_dev,
                                    double* coords_dev, void* stream) {
    int32_t rows = 0, cols = 0, span = 0;
    const int64_t nt = gm_train_tile_grid(H, W, tile_size, overlap, &rows, &cols, &span);
    if (nt < 0 || n < 0) return GM_EINVAL;
    if (n == 0) return GM_OK;
    if (!labels_dev || !flag_dev || !tile_id_dev || !coords_dev) return GM_EINVAL;
    const long long total = (long long)n * span * span;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return GM_ERANGE;
    k_train_label_tiles<<<(unsigned)blocks, 256, 0, gm_stream(stream)>>>(labels_dev, n, H, W, tile_size, tile_size - overlap,
                                                                         cov_threshold, span, cols, rows, flag_dev,
                                                                         tile_id_dev, coords_dev);
    gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
