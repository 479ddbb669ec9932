// a1 + a2: overlapped tile plan and the 3-channel tile gather.
//
// Reference: detect_symbols' tile loop (Detect_OBB.py:210-223) and the 3-ch branch of
// build_multich (Detect_OBB.py:92-93, a contiguous copy of each BGR crop).
//
// The gather is pure HBM traffic: every map byte is read once (plus the overlap re-reads,
// which hit L2) and every tile byte is written once.  Each thread produces one 16-byte
// aligned vector of the packed tile batch; its 16 source bytes sit at an arbitrary byte
// offset of one map row, so they are assembled from five aligned 32-bit loads with funnel
// shifts.  Vectors that straddle a tile-row boundary or the ends of a tile take a per-byte
// path (about 1 in 78 for 416-px tiles).
#include "gm_common.cuh"

extern "C" int64_t gm_tile_plan_count(int32_t H, int32_t W, int32_t tile_size, int32_t overlap,
                                      int32_t* rows, int32_t* cols, int64_t* total_px) {
    if (H <= 0 || W <= 0 || tile_size <= 0 || overlap < 0) return GM_EINVAL;
    const int64_t step = (tile_size - overlap) > 1 ? (tile_size - overlap) : 1;
    const int64_t nr = (H + step - 1) / step;
    const int64_t nc = (W + step - 1) / step;
    if (rows) *rows = (int32_t)nr;
    if (cols) *cols = (int32_t)nc;
    if (total_px) {
        int64_t sh = 0, sw = 0;
        for (int64_t r = 0; r < nr; ++r) {
            int64_t y0 = r * step;
            sh += ((y0 + tile_size < H) ? (y0 + tile_size) : H) - y0;
        }
        for (int64_t c = 0; c < nc; ++c) {
            int64_t x0 = c * step;
            sw += ((x0 + tile_size < W) ? (x0 + tile_size) : W) - x0;
        }
        *total_px = sh * sw;
    }
    return nr * nc;
}

extern "C" int64_t gm_tile_plan_fill(int32_t H, int32_t W, int32_t tile_size, int32_t overlap,
                                     int32_t row_begin, int32_t row_end,
                                     gm_tile* tiles_host, int64_t cap, int64_t* total_px) {
    int32_t nr = 0, nc = 0;
    int64_t n = gm_tile_plan_count(H, W, tile_size, overlap, &nr, &nc, nullptr);
    if (n < 0) return n;
    if (row_end < 0 || row_end > nr) row_end = nr;
    if (row_begin < 0) row_begin = 0;
    if (row_begin > row_end) return GM_EINVAL;
    const int64_t count = (int64_t)(row_end - row_begin) * nc;
    if (!tiles_host || cap < count) return GM_ENOSPC;
    const int64_t step = (tile_size - overlap) > 1 ? (tile_size - overlap) : 1;
    int64_t off = 0, k = 0;
    for (int32_t r = row_begin; r < row_end; ++r) {
        const int64_t y0 = r * step;
        const int64_t h = ((y0 + tile_size < H) ? (y0 + tile_size) : H) - y0;
        for (int32_t c = 0; c < nc; ++c) {
            const int64_t x0 = c * step;
            const int64_t w = ((x0 + tile_size < W) ? (x0 + tile_size) : W) - x0;
            gm_tile t;
            t.y0 = (int32_t)y0; t.x0 = (int32_t)x0; t.h = (int32_t)h; t.w = (int32_t)w;
            t.px_off = off;
            tiles_host[k++] = t;
            off += h * w;
        }
    }
    if (total_px) *total_px = off;
    return count;
}

// ------------------------------------------------------------------------------------------

namespace {

constexpr int GATHER_THREADS = 256;

__global__ void __launch_bounds__(GATHER_THREADS)
k_tile_gather3(const uint8_t* __restrict__ map, int W, long long map_bytes,
               const gm_tile* __restrict__ tiles, uint8_t* __restrict__ out) {
    const gm_tile t = tiles[blockIdx.x];
    const long long row_bytes = 3LL * t.w;
    const long long n_bytes = row_bytes * t.h;
    uint8_t* dst0 = out + 3LL * t.px_off;
    // 16-byte aligned vectors covering [dst0, dst0 + n_bytes)
    const unsigned long long dst_addr = reinterpret_cast<unsigned long long>(dst0);
    const long long lead = (long long)(dst_addr & 15ULL);        // bytes before dst0 in vector 0
    const long long n_vec = (lead + n_bytes + 15) >> 4;
    const long long v = (long long)blockIdx.y * GATHER_THREADS + threadIdx.x;
    if (v >= n_vec) return;
    long long b0 = v * 16 - lead;            // first tile byte of this vector (may be < 0)
    long long b1 = b0 + 16;                  // one past the last
    const uint8_t* src_tile = map + ((long long)t.y0 * W + t.x0) * 3LL;
    const long long src_pitch = 3LL * W;
    const long long r0 = (b0 >= 0 ? b0 : 0) / row_bytes;
    const long long c0 = b0 - r0 * row_bytes;
    const uint8_t* s = src_tile + r0 * src_pitch + c0;
    // the fifth aligned word may reach 3 bytes past s+15: keep it inside the map buffer
    const bool in_map = (s - map) + 20 <= map_bytes;
    if (b0 >= 0 && b1 <= n_bytes && c0 + 16 <= row_bytes && in_map) {
        const unsigned long long sa = reinterpret_cast<unsigned long long>(s);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~3ULL);
        const unsigned sh = (unsigned)(sa & 3ULL) * 8u;
        uint32_t w0 = __ldg(sw), w1 = __ldg(sw + 1), w2 = __ldg(sw + 2), w3 = __ldg(sw + 3);
        uint4 o;
        if (sh == 0) {
            o = make_uint4(w0, w1, w2, w3);
        } else {
            uint32_t w4 = __ldg(sw + 4);
            o.x = __funnelshift_r(w0, w1, sh);
            o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh);
            o.w = __funnelshift_r(w3, w4, sh);
        }
        *reinterpret_cast<uint4*>(dst0 + b0) = o;
        return;
    }
    if (b0 < 0) b0 = 0;
    if (b1 > n_bytes) b1 = n_bytes;
    for (long long b = b0; b < b1; ++b) {
        const long long r = b / row_bytes;
        const long long c = b - r * row_bytes;
        dst0[b] = __ldg(src_tile + r * src_pitch + c);
    }
}

}  // namespace

extern "C" int gm_tile_gather_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                                 const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                                 uint8_t* out_dev, void* stream) {
    if (!map_dev || !tiles_dev || !out_dev || H <= 0 || W <= 0 || max_tile <= 0) return GM_EINVAL;
    if (n_tiles == 0) return GM_OK;
    if (n_tiles < 0) return GM_EINVAL;
    const long long max_bytes = 3LL * max_tile * max_tile;
    const long long max_vec = (max_bytes + 15 + 15) / 16;
    const unsigned gy = (unsigned)((max_vec + GATHER_THREADS - 1) / GATHER_THREADS);
    if (gy > 65535u) return GM_ERANGE;
    dim3 grid((unsigned)n_tiles, gy);
    k_tile_gather3<<<grid, GATHER_THREADS, 0, gm_stream(stream)>>>(map_dev, W, 3LL * H * W, tiles_dev, out_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
