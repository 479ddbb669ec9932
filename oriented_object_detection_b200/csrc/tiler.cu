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
constexpr int GATHER_VPT = 4;          // 16-byte destination vectors per thread

// A CTA produces GATHER_THREADS * GATHER_VPT consecutive 16-byte vectors of one tile; a thread takes
// every GATHER_THREADS-th of them, so each load and store instruction of a warp covers 512 contiguous
// bytes.  A vector's 16 source bytes start at an arbitrary byte of one map row: five aligned 32-bit
// loads cover them (W*3 is not a multiple of 16 in general, so wider loads would need a word rotation
// that costs more ALU-pipe slots than the loads save), four funnel shifts realign them.
// All GATHER_VPT vectors' loads are issued before the first store.  Vectors that straddle a tile-row
// boundary or the ends of a tile take the per-byte path (1 in 78 for 416-px tiles).
// Blocks are numbered tile-major (all chunks of tile 0, then tile 1, ...): the rows two vertically
// adjacent tiles share are then read ~one tile row apart and still sit in L2 (the opposite order reads
// them a whole pass apart and fetches the overlap from DRAM twice).
__global__ void __launch_bounds__(GATHER_THREADS)
k_tile_gather3(const uint8_t* __restrict__ map, int W, long long map_bytes,
               const gm_tile* __restrict__ tiles, int chunks_per_tile, uint8_t* __restrict__ out) {
    const int ti = blockIdx.x / chunks_per_tile;
    const int chunk = blockIdx.x - ti * chunks_per_tile;
    const gm_tile t = tiles[ti];
    // byte offsets inside a tile fit 32 bits (3 * GM_MAX_TILE^2 < 2^22)
    const int rb = 3 * t.w;                                      // bytes per tile row
    const int n_bytes = rb * t.h;
    uint8_t* dst0 = out + 3LL * t.px_off;
    const int lead = (int)(reinterpret_cast<unsigned long long>(dst0) & 15ULL);   // bytes before dst0 in vector 0
    const int n_vec = (lead + n_bytes + 15) >> 4;
    const int vbase = chunk * (GATHER_THREADS * GATHER_VPT) + threadIdx.x;
    if (vbase - (int)threadIdx.x >= n_vec) return;
    const long long tile_off = ((long long)t.y0 * W + t.x0) * 3LL;
    const uint8_t* src_tile = map + tile_off;
    const long long src_pitch = 3LL * W;
    const long long room = map_bytes - tile_off - 20;            // five aligned words from offset o stay inside the map iff o <= room
    const unsigned int magic = 0xffffffffu / (unsigned int)rb + 1u;     // ceil(2^32 / rb): row = offset / rb by multiply-high
    uint32_t w[GATHER_VPT][5];
    int boff[GATHER_VPT];
    unsigned int sh[GATHER_VPT];
    bool fast[GATHER_VPT];
#pragma unroll
    for (int k = 0; k < GATHER_VPT; ++k) {
        const int v = vbase + k * GATHER_THREADS;
        const int b0 = v * 16 - lead;                            // first tile byte of the vector (may be < 0)
        const unsigned int ub = (unsigned int)max(b0, 0);
        unsigned int r = __umulhi(ub, magic);
        if (r * (unsigned int)rb > ub) --r;
        const int c = b0 - (int)r * rb;
        const long long so = (long long)r * src_pitch + c;
        boff[k] = b0;
        fast[k] = v < n_vec && b0 >= 0 && b0 + 16 <= n_bytes && c + 16 <= rb && so <= room;
        sh[k] = 0u;
        if (fast[k]) {
            const unsigned long long sa = reinterpret_cast<unsigned long long>(src_tile + so);
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(sa & ~3ULL);
            sh[k] = (unsigned int)(sa & 3ULL) * 8u;
#pragma unroll
            for (int i = 0; i < 5; ++i) w[k][i] = __ldg(sw + i);
        }
    }
#pragma unroll
    for (int k = 0; k < GATHER_VPT; ++k) {
        const int v = vbase + k * GATHER_THREADS;
        if (v >= n_vec) break;
        if (fast[k]) {
            *reinterpret_cast<uint4*>(dst0 + boff[k]) = make_uint4(__funnelshift_r(w[k][0], w[k][1], sh[k]), __funnelshift_r(w[k][1], w[k][2], sh[k]),
                                                                   __funnelshift_r(w[k][2], w[k][3], sh[k]), __funnelshift_r(w[k][3], w[k][4], sh[k]));
        } else {
            int a = max(boff[k], 0);
            const int e = min(boff[k] + 16, n_bytes);
            for (; a < e; ++a) {
                unsigned int r = __umulhi((unsigned int)a, magic);
                if (r * (unsigned int)rb > (unsigned int)a) --r;
                dst0[a] = __ldg(src_tile + (long long)r * src_pitch + (a - (int)r * rb));
            }
        }
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
    const long long chunks = (max_vec + GATHER_THREADS * GATHER_VPT - 1) / (GATHER_THREADS * GATHER_VPT);
    const long long blocks = chunks * n_tiles;
    if (blocks > 0x7fffffffLL) return GM_ERANGE;
    k_tile_gather3<<<(unsigned)blocks, GATHER_THREADS, 0, gm_stream(stream)>>>(map_dev, W, 3LL * H * W, tiles_dev, (int)chunks, out_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}
