// a3: 4-channel [R,G,B,DT-Edge] builder for a batch of tiles.
//
// Reference: build_multich 4-ch branch (Detect_OBB.py:95-133) and its Train twin
// dt_edge_channel_from_bgr / build_4ch_CHW_from_bgr_dtedge (Train_OBB.py:615-664), at the
// bit level OpenCV 4.13 (IPP off) + numpy 2.3 compute them (restated
// and pinned in oracle/pixel.py).  Every tile is independent (borders, percentiles and
// min/max are tile-local), so the batch dimension is the tile.
//
// Pipeline (one launch each, all tiles at once, intermediates in a caller workspace):
//   k_grad         BGR map -> gray -> fixed-point Gaussian stack -> Scharr -> S = max_s(gx^2+gy^2)
//   k_select_grad  per tile: exact order statistics of S at the p_hi percentile -> numpy's
//                  float64 lerp threshold -> integer threshold S_thr; min/max -> normalize consts
//                  (k_otsu_grad instead when DT_BIN_METHOD = "otsu": GM_DTEDGE_OTSU)
//   k_edge_open    S >= S_thr -> 3x3 cross open -> bit-packed zero mask
//   k_chamfer      per tile, one warp: 3x3 chamfer DT (16.16 fixed point), forward + backward
//                  raster pass as per-row min-plus scans
//   k_select_dist  per tile: p1 / p99 of the chamfer field
//   k_tail         normalise, exp(-d/3) in float64, blend with the normalised gradient, pack
//                  [R,G,B,DT] (HWC or CHW)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <mutex>
#include <type_traits>
#include <cstring>
#include "gm_common.cuh"
#include "dtedge_grad.cuh"
#include "dtedge_otsu.cuh"

// chunk x stream split of a whole-plan build (gm_dtedge_build_u8); 1 x 1 = one range on the caller's stream
#ifndef GM_GRAD_DEFAULT_CARVEOUT
#define GM_GRAD_DEFAULT_CARVEOUT -1
#endif
#ifndef GM_SELECT_DEFAULT_THREADS
#define GM_SELECT_DEFAULT_THREADS 1024   // threads per selection CTA (one CTA per tile)
#endif
#ifndef GM_DTEDGE_DEFAULT_SKEW
#define GM_DTEDGE_DEFAULT_SKEW 0
#endif
#ifndef GM_DTEDGE_DEFAULT_CHUNKS
#define GM_DTEDGE_DEFAULT_CHUNKS 1
#define GM_DTEDGE_DEFAULT_STREAMS 1
#endif

namespace {

// ------------------------------------------------------------------------------------------
// Gaussian taps (host): 8.8 fixed point, as cv2 builds them for CV_8U sources.

struct GmTaps {
    int n_scales;
    int max_radius;
    int radius[GM_MAX_SCALES];
    unsigned short taps[GM_MAX_SCALES][2 * GM_MAX_RADIUS + 1];
};

int make_taps(const gm_dtedge_params* p, GmTaps* out) {
    if (p->n_sigmas < 1 || p->n_sigmas > GM_MAX_SCALES) return GM_EINVAL;
    out->n_scales = p->n_sigmas;
    out->max_radius = 0;
    for (int s = 0; s < p->n_sigmas; ++s) {
        const double sigma = (double)p->sigmas[s];
        if (!(sigma > 0.0)) {
            out->radius[s] = 0;
            out->taps[s][0] = 256;
            continue;
        }
        const int n = ((int)nearbyint(sigma * 6.0 + 1.0)) | 1;     // round half to even like cvRound
        const int half = n / 2;
        if (half > GM_MAX_RADIUS) return GM_ERANGE;
        double vals[GM_MAX_RADIUS];
        const double scale = -0.5 * 0.25 / (sigma * sigma);
        double sum = 0.0;
        int x = 1 - n;
        for (int i = 0; i < half; ++i, x += 2) {
            vals[i] = exp((double)(x * x) * scale);
            sum += vals[i];
        }
        const double inv = 1.0 / (2.0 * sum + 1.0);
        double err = 0.0;
        int acc = 0;
        for (int i = 0; i < half; ++i) {
            const double adj = vals[i] * inv * 256.0 + err;
            const int v = (int)nearbyint(adj);
            err = adj - (double)v;
            out->taps[s][i] = (unsigned short)v;
            out->taps[s][n - 1 - i] = (unsigned short)v;
            acc += v;
        }
        out->taps[s][half] = (unsigned short)(256 - 2 * acc);
        out->radius[s] = half;
        if (half > out->max_radius) out->max_radius = half;
    }
    return GM_OK;
}

// The reference's configured stack (0, 0.6, 1.2, 2.4) -> radii (0, 2, 4, 7) with byte-sized taps
// takes the IDP-based kernel of dtedge_grad.cuh; everything else the generic k_grad below.
bool fast_grad_coef(const GmTaps& t, gradfast::Coef* c) {
    if (t.n_scales != 4 || t.radius[0] != 0) return false;
    unsigned short packed[3][15];
    for (int s = 0; s < 3; ++s) {
        const int R = gradfast::radius_of(s);
        if (t.radius[s + 1] != R) return false;
        for (int k = 0; k <= 2 * R; ++k) {
            if (t.taps[s + 1][k] > 255) return false;
            packed[s][k] = t.taps[s + 1][k];
        }
    }
    gradfast::pack_coef(packed, c);
    return true;
}

// Per-tile constants produced by the two select kernels.
struct TileParams {
    unsigned int s_thr;      // edge <=> S >= s_thr
    float nrm_scale;         // cv2.normalize(acc, 0, 1, MINMAX): fmaf(acc, scale, shift)
    float nrm_shift;
    float inv_den;           // fp32 1/dist_den for the fast path of k_tail
    double dist_lo;          // p1 of dist
    double dist_den;         // max(1e-6, p99 - p1)
};

// zero-mask words of tile `ti` start here (see gm_dtedge_workspace_views)
__host__ __device__ inline long long zbits_offset(long long px_off, int ti, int max_tile) {
    return (px_off >> 5) + (long long)ti * (max_tile + 1);
}

// ------------------------------------------------------------------------------------------
// k_grad: 32x32 output pixels per CTA.

constexpr int GB = 32;          // block edge
constexpr int GRAD_THREADS = 256;
constexpr int GRAD_PIX = GB * GB / GRAD_THREADS;

__global__ void __launch_bounds__(GRAD_THREADS)
k_grad(const uint8_t* __restrict__ map, int W, const gm_tile* __restrict__ tiles,
       const GmTaps taps, unsigned int* __restrict__ S_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const gm_tile t = tiles[blockIdx.x];
    const int nbx = (t.w + GB - 1) / GB;
    const int nby = (t.h + GB - 1) / GB;
    if ((int)blockIdx.y >= nbx * nby) return;
    const int bx = ((int)blockIdx.y % nbx) * GB;
    const int by = ((int)blockIdx.y / nbx) * GB;
    const int R = taps.max_radius;
    const int HALO = R + 1;
    const int PW = GB + 2 * HALO;
    const int PH = GB + 2 * HALO;
    constexpr int NC = GB + 2;                     // columns kept after the horizontal pass
    unsigned char* gray = smem;                                            // PH x PW
    unsigned short* hbuf = reinterpret_cast<unsigned short*>(smem + ((PH * PW + 15) & ~15));  // (GB+2+2R) x NC
    unsigned char* blur = reinterpret_cast<unsigned char*>(hbuf + (GB + 2 + 2 * R) * NC);      // NC x NC
    const int tid = threadIdx.x;

    // gray patch with REFLECT_101 taken at the TILE border; a symmetric kernel over the
    // reflected extension reproduces cv2's border handling of every later stage as well.
    for (int i = tid; i < PH * PW; i += GRAD_THREADS) {
        const int py = i / PW;
        const int px = i - py * PW;
        const int ty = gm_reflect101(by - HALO + py, t.h);
        const int tx = gm_reflect101(bx - HALO + px, t.w);
        const uint8_t* p = map + ((long long)(t.y0 + ty) * W + (t.x0 + tx)) * 3LL;
        const unsigned b = __ldg(p), g = __ldg(p + 1), r = __ldg(p + 2);
        gray[i] = (unsigned char)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
    }
    __syncthreads();

    unsigned int smax[GRAD_PIX];
#pragma unroll
    for (int k = 0; k < GRAD_PIX; ++k) smax[k] = 0u;

    for (int s = 0; s < taps.n_scales; ++s) {
        const int r = taps.radius[s];
        const unsigned char* img;
        int pitch, base;
        if (r == 0) {
            img = gray; pitch = PW; base = (HALO - 1) * PW + (HALO - 1);
        } else {
            const int nrows = GB + 2 + 2 * r;
            const int ntap = 2 * r + 1;
            for (int i = tid; i < nrows * NC; i += GRAD_THREADS) {
                const int row = i / NC;
                const int col = i - row * NC;
                const unsigned char* g = gray + (HALO - 1 - r + row) * PW + (HALO - 1 + col - r);
                unsigned int a = 0;
                for (int k = 0; k < ntap; ++k) a += (unsigned int)taps.taps[s][k] * g[k];
                hbuf[i] = (unsigned short)a;
            }
            __syncthreads();
            for (int i = tid; i < NC * NC; i += GRAD_THREADS) {
                const int y = i / NC;
                const int x = i - y * NC;
                const unsigned short* hp = hbuf + y * NC + x;
                unsigned int v = 0;
                for (int k = 0; k < ntap; ++k) v += (unsigned int)taps.taps[s][k] * hp[k * NC];
                blur[i] = (unsigned char)((v + 32768u) >> 16);
            }
            __syncthreads();
            img = blur; pitch = NC; base = 0;
        }
#pragma unroll
        for (int k = 0; k < GRAD_PIX; ++k) {
            const int oy = (tid >> 5) + k * (GRAD_THREADS / 32);
            const int ox = tid & 31;
            const unsigned char* q = img + base + oy * pitch + ox;
            const int p00 = q[0], p01 = q[1], p02 = q[2];
            const int p10 = q[pitch], p12 = q[pitch + 2];
            const int p20 = q[2 * pitch], p21 = q[2 * pitch + 1], p22 = q[2 * pitch + 2];
            const int gx = 3 * (p02 - p00) + 10 * (p12 - p10) + 3 * (p22 - p20);
            const int gy = 3 * (p20 - p00) + 10 * (p21 - p01) + 3 * (p22 - p02);
            const unsigned int S = (unsigned int)(gx * gx + gy * gy);
            smax[k] = max(smax[k], S);
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < GRAD_PIX; ++k) {
        const int oy = by + (tid >> 5) + k * (GRAD_THREADS / 32);
        const int ox = bx + (tid & 31);
        if (oy < t.h && ox < t.w) S_out[t.px_off + (long long)oy * t.w + ox] = smax[k];
    }
}

// ------------------------------------------------------------------------------------------
// Exact order statistics of a tile's uint32 keys: one CTA per tile.
//   pass 0 (global)  histogram over 2112 logarithmic bins (exponent + 6 mantissa bits), min / max;
//                    16-byte loads, two in flight per thread; zero keys counted in a register.
//   pass 1 (global)  the keys of the (<= 4) bins holding the wanted ranks are compacted into a
//                    shared-memory list (a log bin holds ~0.5 % of a tile's keys);
//   finish           MSB-first 11-bit radix select over that list, in shared memory.
// If the bins hold more keys than the list (constant tiles, few distinct values) the radix passes run
// over the global keys instead.  Keys inside one bin with exponent < 6 are a single value: no finish.

constexpr int SEL_THREADS = 1024;      // launch bound; the kernels run with any multiple of 32 up to it (SEL_NT)
#define SEL_NT ((int)blockDim.x)
constexpr int SEL_BINS = 2048;         // refinement histogram: 11 bits per pass
constexpr int SEL_LOGBINS = 33 * 64;   // pass 0
constexpr int SEL_MAXR = 4;
#ifndef GM_SEL_SLOTS
#define GM_SEL_SLOTS 24
#endif
#ifndef GM_SEL_DIST_SIGMAS
#define GM_SEL_DIST_SIGMAS 12
#endif
constexpr int SEL_LIST = GM_SEL_SLOTS * 1024;   // compacted keys kept in shared memory (the sampled path gives every thread SEL_LIST / blockDim.x private slots)

__device__ __forceinline__ int log_bin(unsigned int key) {
    if (key == 0u) return 0;
    const int e = 31 - __clz(key);
    const unsigned int top = (e >= 6) ? (key >> (e - 6)) : (key << (6 - e));
    return ((e + 1) << 6) | (int)(top & 63u);
}

// Warp-cooperative: first bin whose inclusive prefix count exceeds `target`; returns the
// bin and the count strictly below it.
__device__ void warp_find_bin(const unsigned int* hist, int nbins, unsigned int target,
                              int* bin_out, unsigned int* below_out) {
    const int lane = gm_lane();
    const int per = (nbins + 31) / 32;
    const int b0 = lane * per;
    unsigned int sum = 0;
    for (int i = 0; i < per; ++i) {
        const int b = b0 + i;
        if (b < nbins) sum += hist[b];
    }
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const unsigned int excl = incl - sum;
    const unsigned int hit = __ballot_sync(0xffffffffu, incl > target);
    const int src = hit ? (__ffs(hit) - 1) : 31;
    if (lane == src) {
        unsigned int c = excl;
        int found = min(b0 + per, nbins) - 1;
        for (int i = 0; i < per; ++i) {
            const int b = b0 + i;
            if (b >= nbins) break;
            const unsigned int hcount = hist[b];
            if (c + hcount > target) { found = b; break; }
            c += hcount;
        }
        *bin_out = found;
        *below_out = c;
    }
}

struct SelShared {
    unsigned int hist[SEL_MAXR][SEL_BINS];   // pass 0 uses the first SEL_LOGBINS words of the flat array
    unsigned int list[SEL_LIST];
    unsigned int lo[SEL_MAXR];      // low end of the key interval still holding the rank
    int rem[SEL_MAXR];              // undecided low bits
    unsigned int base[SEL_MAXR];    // number of keys below the interval
    int hid[SEL_MAXR];              // histogram used by this rank (ranks in one interval share)
    int bin[SEL_MAXR];
    unsigned int below[SEL_MAXR];
    unsigned int red_min[32], red_max[32];
    unsigned int n_list;
    unsigned int n_cand;
};

// Calls f(key) for every key of a tile: 16-byte loads over the aligned body, two per thread in
// flight, scalar head and tail.
template <typename F>
__device__ __forceinline__ void scan_keys(const unsigned int* __restrict__ keys, int n, F f) {
    const int tid = threadIdx.x;
    const int head = min(n, (int)((4u - (unsigned int)((reinterpret_cast<unsigned long long>(keys) >> 2) & 3ULL)) & 3u));
    const int nvec = (n - head) >> 2;
    const uint4* kv = reinterpret_cast<const uint4*>(keys + head);
    int i = tid;
    for (; i + SEL_NT < nvec; i += 2 * SEL_NT) {
        const uint4 a = kv[i];
        const uint4 b = kv[i + SEL_NT];
        f(a.x); f(a.y); f(a.z); f(a.w);
        f(b.x); f(b.y); f(b.z); f(b.w);
    }
    if (i < nvec) {
        const uint4 a = kv[i];
        f(a.x); f(a.y); f(a.z); f(a.w);
    }
    if (tid < head) f(keys[tid]);
    const int t0 = head + 4 * nvec;
    if (t0 + tid < n) f(keys[t0 + tid]);
}

// MSB-first 11-bit radix refinement of the intervals [lo[r], lo[r] + 2^rem[r]) (base[r] keys lie below
// lo[r]) until every rank is a single key.  visit(f) calls f(key) for this thread's share of the candidate
// keys; keys outside an interval are ignored, so the candidates may be any superset of the keys <= the
// wanted ones inside the interval.
template <typename Visit, typename SH>
__device__ __forceinline__ void refine_ranks(Visit visit, const unsigned int* ranks, int nr, SH& sh) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    for (;;) {
        int any = 0;
        for (int r = 0; r < nr; ++r) any |= sh.rem[r];
        if (!any) break;
        // share a histogram between ranks that are still in the same interval
        if (tid == 0) {
            int nh = 0;
            for (int r = 0; r < nr; ++r) {
                if (sh.rem[r] == 0) { sh.hid[r] = -1; continue; }
                int found = -1;
                for (int q = 0; q < r; ++q)
                    if (sh.hid[q] >= 0 && sh.lo[q] == sh.lo[r] && sh.rem[q] == sh.rem[r]) { found = sh.hid[q]; break; }
                sh.hid[r] = found >= 0 ? found : nh++;
            }
        }
        for (int i = tid; i < SEL_MAXR * SEL_BINS; i += SEL_NT) (&sh.hist[0][0])[i] = 0u;
        __syncthreads();
        unsigned int lo_r[SEL_MAXR];
        int rem_r[SEL_MAXR], sft_r[SEL_MAXR], hid_r[SEL_MAXR];
#pragma unroll
        for (int r = 0; r < SEL_MAXR; ++r) {
            hid_r[r] = -1; lo_r[r] = 0u; rem_r[r] = 0; sft_r[r] = 0;
            if (r < nr) {
                lo_r[r] = sh.lo[r]; rem_r[r] = sh.rem[r]; hid_r[r] = sh.hid[r];
                const int bits = min(11, rem_r[r]);
                sft_r[r] = rem_r[r] - bits;
                // only the first rank of a shared histogram counts into it.  (The ids are compared as shared, not as
                // already marked: the first form marked a rank by negating its id in place and compared against the marked
                // ids, so a THIRD rank of one histogram toggled back and every key was counted twice - reachable only when
                // three or more ranks refine inside one interval, which the small-tile kernels do from their first round.)
                bool dup = false;
#pragma unroll
                for (int q = 0; q < SEL_MAXR; ++q) dup |= (q < r) && (q < nr) && (sh.hid[q] >= 0) && (sh.hid[q] == sh.hid[r]);
                if (dup) hid_r[r] = -2;
            }
        }
        auto count = [&](unsigned int k) {
#pragma unroll
            for (int r = 0; r < SEL_MAXR; ++r) {
                if (hid_r[r] >= 0) {
                    const unsigned int d = k - lo_r[r];
                    // rem == 32 (the whole key range: small-tile kernels start there) - a 32-bit shift is not defined in C
                    if (k >= lo_r[r] && (rem_r[r] >= 32 || (d >> rem_r[r]) == 0u)) atomicAdd(&sh.hist[hid_r[r]][d >> sft_r[r]], 1u);
                }
            }
        };
        visit(count);
        __syncthreads();
        if (warp < nr && sh.rem[warp] > 0) {
            const int h = sh.hid[warp];
            warp_find_bin(sh.hist[h], SEL_BINS, ranks[warp] - sh.base[warp], &sh.bin[warp], &sh.below[warp]);
        }
        __syncthreads();
        if (tid < nr && sh.rem[tid] > 0) {
            const int bits = min(11, sh.rem[tid]);
            const int sft = sh.rem[tid] - bits;
            sh.lo[tid] += (unsigned int)sh.bin[tid] << sft;
            sh.rem[tid] = sft;
            sh.base[tid] += sh.below[tid];
        }
        __syncthreads();
    }
}

// Sampled front end of the selection (tiles of >= SEL_SAMPLE_MIN keys; ranks come in pairs (r, r+1), one
// pair per percentile).  The two full passes of block_select cost ~40 thread instructions per key; here
//   1. every SEL_STRIDE-th 16-byte vector (~6 % of the keys) goes into the log-bin histogram;
//   2. per pair, the bin edges around the sample ranks  r*m/n -+ (SIGMAS sigma + 8)  give a key interval [lo, hi]
//      that holds both ranks unless the sample is badly off;
//   3. ONE pass over all keys counts the keys below lo and inside [lo, hi] per pair (+ min / max) and
//      keeps the keys inside in shared memory, each thread in its own slots - no atomics, no branches
//      (a single-valued interval needs no list);
//   4. the counts PROVE whether [lo, hi] holds the wanted ranks.  If so the radix refinement runs over
//      the list and the result is the exact order statistic; if not (or the list overflowed) the caller
//      falls through to the two-pass method.  Either way the answer is exact, never approximate.
constexpr int SEL_SAMPLE_MIN = 32768;
constexpr int SEL_STRIDE = 16;
constexpr int SEL_SAMPLE_VECS = 3072;       // at most this many sampled vectors (larger tiles sample sparser)

__device__ __forceinline__ void log_bin_edges(int bin, unsigned int* lo, unsigned int* hi) {
    if (bin == 0) { *lo = 0u; *hi = 0u; return; }
    const int e = (bin >> 6) - 1;
    const unsigned int m = (unsigned int)(bin & 63);
    if (e >= 6) { *lo = (64u | m) << (e - 6); *hi = *lo + ((1u << (e - 6)) - 1u); }
    else { *lo = (64u | m) >> (6 - e); *hi = *lo; }
}

__device__ unsigned int g_sel_stats[8];      // sampled-path outcomes: [2*(NB-1)] proven, [2*(NB-1)+1] fell back; [4] list overflow, [5]/[6] rank outside bracket 0/1, [7] bracket too wide

template <int NB, bool MINMAX, int SIGMAS>
__device__ bool sampled_select(const unsigned int* __restrict__ keys, int n, const unsigned int* ranks,
                               unsigned int* vals, unsigned int* kmin, unsigned int* kmax, SelShared& sh) {
    static_assert(NB == 1 || NB == 2, "one or two percentiles");
    if (n < SEL_SAMPLE_MIN) return false;
    __shared__ unsigned int s_lo[NB], s_span[NB], s_below[NB], s_in[NB];
    __shared__ int s_bin[2 * NB];
    __shared__ unsigned int s_scratch[2 * NB];
    __shared__ int s_ok, s_over;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    unsigned int* hist0 = &sh.hist[0][0];
    for (int i = tid; i < SEL_LOGBINS; i += SEL_NT) hist0[i] = 0u;
    if (tid == 0) s_over = 0;
    if (tid < NB) { s_below[tid] = 0u; s_in[tid] = 0u; }
    __syncthreads();
    const int head = min(n, (int)((4u - (unsigned int)((reinterpret_cast<unsigned long long>(keys) >> 2) & 3ULL)) & 3u));
    const int nvec = (n - head) >> 2;
    const uint4* kv = reinterpret_cast<const uint4*>(keys + head);
    const int stride = max(SEL_STRIDE, (nvec + SEL_SAMPLE_VECS - 1) / SEL_SAMPLE_VECS);
    const int ns = (nvec + stride - 1) / stride;
    const int m = 4 * ns;
    {
        unsigned int zeros = 0u;
        auto f = [&](unsigned int k) { if (k == 0u) ++zeros; else atomicAdd(&hist0[log_bin(k)], 1u); };
        for (int i = tid; i < ns; i += SEL_NT) {
            const uint4 a = kv[(long long)i * stride];
            f(a.x); f(a.y); f(a.z); f(a.w);
        }
        if (zeros) atomicAdd(&hist0[0], zeros);
    }
    __syncthreads();
    if (warp < 2 * NB) {
        const int b = warp >> 1, upper = warp & 1;
        const float q = (float)ranks[2 * b] / (float)n;
        const int d = (int)((float)SIGMAS * sqrtf((float)m * q * (1.f - q))) + 8;
        long long sr = (long long)ranks[2 * b + upper] * m / n + (upper ? d + 1 : -d);
        // a sample rank below 0 opens the interval downwards (bin -1 -> lo = 0); one beyond the sample stops at
        // the bin of the largest sampled key (a wanted key above it fails the proof and takes the two-pass path)
        if (sr >= m) sr = m - 1;
        if (sr >= 0) warp_find_bin(hist0, SEL_LOGBINS, (unsigned int)sr, &s_bin[warp], &s_scratch[warp]);
        else if ((tid & 31) == 0) s_bin[warp] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned int lo[NB], hi[NB];
        for (int b = 0; b < NB; ++b) {
            unsigned int e0, e1;
            if (s_bin[2 * b] < 0) lo[b] = 0u; else { log_bin_edges(s_bin[2 * b], &e0, &e1); lo[b] = e0; }
            if (s_bin[2 * b + 1] < 0) hi[b] = 0xffffffffu; else { log_bin_edges(s_bin[2 * b + 1], &e0, &e1); hi[b] = e1; }
            if (hi[b] < lo[b]) hi[b] = lo[b];
        }
        if (NB == 2 && lo[NB - 1] <= hi[0]) {         // overlapping intervals: both pairs use the union
            const unsigned int l = min(lo[0], lo[NB - 1]), h = max(hi[0], hi[NB - 1]);
            lo[0] = lo[NB - 1] = l; hi[0] = hi[NB - 1] = h;
        }
        for (int b = 0; b < NB; ++b) { s_lo[b] = lo[b]; s_span[b] = hi[b] - lo[b]; }
    }
    __syncthreads();
    // Counting pass.  Branch free and atomic free: a key inside a (multi-valued) interval goes to the next of
    // this thread's private slots, list[tid + j * blockDim.x]; a thread that runs out of slots only counts.
    // The per-key sequence is spelled in PTX (predicated add / store) - the compiler's own rendering of the
    // same C++ took 13 instructions and a branch per key instead of 8 and none.
    const int nt = SEL_NT;
    const unsigned int list_s = (unsigned int)__cvta_generic_to_shared(sh.list);
    const unsigned int slot_end = list_s + 4u * (unsigned int)((SEL_LIST / nt) * nt);   // slot addresses at or beyond do not exist
    unsigned int slot = list_s + 4u * (unsigned int)tid;                               // next private slot (shared address)
    {
        unsigned int lo[NB], span[NB], below[NB], in[NB], step[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            lo[b] = s_lo[b]; span[b] = s_span[b]; below[b] = 0u; in[b] = 0u;
            step[b] = span[b] != 0u ? 4u * (unsigned int)nt : 0u;                  // single-valued interval: nothing to keep
        }
        if (NB == 2 && lo[0] == lo[NB - 1] && span[0] == span[NB - 1]) step[NB - 1] = 0u;   // merged intervals: keep a key once
        unsigned int end[NB];                                                      // step 0: no slot ever qualifies
#pragma unroll
        for (int b = 0; b < NB; ++b) end[b] = step[b] != 0u ? slot_end : 0u;
        unsigned int mn = 0xffffffffu, mx = 0u;
        // zflag: the first interval of a pair of percentiles is the single key 0 (p1 of a distance field with
        // more than ~2 % edge pixels) - then it only needs its zeros counted.
        auto pass = [&](auto zflag) {
            constexpr bool Z0 = decltype(zflag)::value;
            auto visit1 = [&](unsigned int k) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (Z0 && b == 0) {
                        asm volatile("{\n\t"
                                     ".reg .pred pz;\n\t"
                                     "setp.eq.u32 pz, %1, 0;\n\t"
                                     "@pz add.u32 %0, %0, 1;\n\t"
                                     "}"
                                     : "+r"(in[b]) : "r"(k));
                        continue;
                    }
                    asm volatile("{\n\t"
                                 ".reg .pred pb, pi, ps;\n\t"
                                 ".reg .u32 t;\n\t"
                                 "setp.lt.u32 pb, %3, %4;\n\t"
                                 "@pb add.u32 %0, %0, 1;\n\t"
                                 "sub.u32 t, %3, %4;\n\t"
                                 "setp.le.u32 pi, t, %5;\n\t"
                                 "@pi add.u32 %1, %1, 1;\n\t"
                                 "setp.lt.and.u32 ps, %2, %6, pi;\n\t"
                                 "@ps st.shared.u32 [%2], %3;\n\t"
                                 "@pi add.u32 %2, %2, %7;\n\t"
                                 "}"
                                 : "+r"(below[b]), "+r"(in[b]), "+r"(slot)
                                 : "r"(k), "r"(lo[b]), "r"(span[b]), "r"(end[b]), "r"(step[b])
                                 : "memory");
                }
            };
            auto visit4 = [&](const uint4 a) {
                if (MINMAX) { mn = min(mn, min(min(a.x, a.y), min(a.z, a.w))); mx = max(mx, max(max(a.x, a.y), max(a.z, a.w))); }
                visit1(a.x); visit1(a.y); visit1(a.z); visit1(a.w);
            };
            int i = tid;
            for (; i + 3 * nt < nvec; i += 4 * nt) {
                const uint4 a = kv[i], c = kv[i + nt], e = kv[i + 2 * nt], g = kv[i + 3 * nt];
                visit4(a); visit4(c); visit4(e); visit4(g);
            }
            for (; i < nvec; i += nt) visit4(kv[i]);
            if (tid < head) { const unsigned int k = keys[tid]; if (MINMAX) { mn = min(mn, k); mx = max(mx, k); } visit1(k); }
            const int t0 = head + 4 * nvec;
            if (t0 + tid < n) { const unsigned int k = keys[t0 + tid]; if (MINMAX) { mn = min(mn, k); mx = max(mx, k); } visit1(k); }
        };
        if (NB == 2 && lo[0] == 0u && span[0] == 0u) pass(std::true_type{});
        else pass(std::false_type{});
        if (slot >= slot_end + 4u * (unsigned int)nt) s_over = 1;        // more keys than private slots (benign race: all writers store 1)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                below[b] += __shfl_xor_sync(0xffffffffu, below[b], d);
                in[b] += __shfl_xor_sync(0xffffffffu, in[b], d);
            }
            if (MINMAX) {
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
            }
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) { atomicAdd(&s_below[b], below[b]); atomicAdd(&s_in[b], in[b]); }
            if (MINMAX) { sh.red_min[warp] = mn; sh.red_max[warp] = mx; }
        }
    }
    __syncthreads();
    if (tid == 0) {
        bool ok = s_over == 0;                                // some thread ran out of private slots
        if (!ok) atomicAdd(&g_sel_stats[4], 1u);
        for (int b = 0; b < NB; ++b) {
            const bool holds = s_below[b] <= ranks[2 * b] && ranks[2 * b + 1] < s_below[b] + s_in[b];
            if (!holds) atomicAdd(&g_sel_stats[5 + b], 1u);
            const int rem = s_span[b] == 0u ? 0 : 32 - __clz(s_span[b]);
            if (rem >= 32) atomicAdd(&g_sel_stats[7], 1u);
            ok = ok && holds && rem < 32;
            for (int e = 0; e < 2; ++e) {
                sh.lo[2 * b + e] = s_lo[b]; sh.rem[2 * b + e] = rem; sh.base[2 * b + e] = s_below[b];
            }
        }
        s_ok = ok ? 1 : 0;
        atomicAdd(&g_sel_stats[2 * (NB - 1) + (ok ? 0 : 1)], 1u);
        if (MINMAX && ok) {
            unsigned int a = 0xffffffffu, c = 0u;
            for (int w = 0; w < SEL_NT / 32; ++w) { a = min(a, sh.red_min[w]); c = max(c, sh.red_max[w]); }
            *kmin = a; *kmax = c;
        }
    }
    __syncthreads();
    if (!s_ok) return false;
    {
        const int n_mine = (int)((min(slot, slot_end + 4u * (unsigned int)tid) - list_s) >> 2);       // one past this thread's last kept slot
        refine_ranks([&](auto f) { for (int i = tid; i < n_mine; i += nt) f(sh.list[i]); }, ranks, 2 * NB, sh);
    }
    if (tid < 2 * NB) vals[tid] = sh.lo[tid];
    __syncthreads();
    return true;
}

// ranks[] ascending, nr <= SEL_MAXR.  On return vals[r] = the rank-th smallest key (0-based).
// pre_hist != nullptr: pass 0 was done by the producer of the keys (global histogram + min/max).
__device__ void block_select(const unsigned int* __restrict__ keys, int n, const unsigned int* ranks,
                             int nr, unsigned int* vals, unsigned int* kmin, unsigned int* kmax,
                             SelShared& sh) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    unsigned int* hist0 = &sh.hist[0][0];
    for (int i = tid; i < SEL_LOGBINS; i += SEL_NT) hist0[i] = 0u;
    if (tid == 0) { sh.n_list = 0u; sh.n_cand = 0u; }
    __syncthreads();
    {
        unsigned int mn = 0xffffffffu, mx = 0u, zeros = 0u;
        scan_keys(keys, n, [&](unsigned int k) {
            mn = min(mn, k); mx = max(mx, k);
            if (k == 0u) ++zeros; else atomicAdd(&hist0[log_bin(k)], 1u);
        });
        if (zeros) atomicAdd(&hist0[0], zeros);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        if ((tid & 31) == 0) { sh.red_min[warp] = mn; sh.red_max[warp] = mx; }
    }
    __syncthreads();
    if (warp < nr) warp_find_bin(hist0, SEL_LOGBINS, ranks[warp], &sh.bin[warp], &sh.below[warp]);
    if (tid == 0) {
        unsigned int a = 0xffffffffu, b = 0u;
        for (int w = 0; w < SEL_NT / 32; ++w) { a = min(a, sh.red_min[w]); b = max(b, sh.red_max[w]); }
        *kmin = a; *kmax = b;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned int cand = 0u;
        for (int r = 0; r < nr; ++r) {
            const int bin = sh.bin[r];
            const int e = (bin >> 6) - 1;
            const unsigned int m = (unsigned int)(bin & 63);
            if (bin == 0) { sh.lo[r] = 0u; sh.rem[r] = 0; }
            else if (e >= 6) { sh.lo[r] = (64u | m) << (e - 6); sh.rem[r] = e - 6; }
            else { sh.lo[r] = (64u | m) >> (6 - e); sh.rem[r] = 0; }
            sh.base[r] = sh.below[r];
            bool dup = false;
            for (int q = 0; q < r; ++q) dup |= (sh.bin[q] == bin);
            if (!dup && sh.rem[r] > 0) cand += hist0[bin];
        }
        sh.n_cand = cand;
    }
    __syncthreads();
    {
        int any = 0;
        for (int r = 0; r < nr; ++r) any |= sh.rem[r];
        if (!any) {
            if (tid < nr) vals[tid] = sh.lo[tid];
            __syncthreads();
            return;
        }
    }
    // the refinement passes read either the compacted list or, if it would not fit, the global keys
    const bool use_list = sh.n_cand <= (unsigned int)SEL_LIST;
    if (use_list) {
        unsigned int lo_r[SEL_MAXR];
        int rem_r[SEL_MAXR];
#pragma unroll
        for (int r = 0; r < SEL_MAXR; ++r) {
            lo_r[r] = 0u; rem_r[r] = -1;
            if (r < nr && sh.rem[r] > 0) {
                bool dup = false;
#pragma unroll
                for (int q = 0; q < SEL_MAXR; ++q) dup |= (q < r) && (sh.bin[q] == sh.bin[r]);
                if (!dup) { lo_r[r] = sh.lo[r]; rem_r[r] = sh.rem[r]; }
            }
        }
        scan_keys(keys, n, [&](unsigned int k) {
            bool in = false;
#pragma unroll
            for (int r = 0; r < SEL_MAXR; ++r)
                if (rem_r[r] >= 0) in |= (k >= lo_r[r]) && (((k - lo_r[r]) >> rem_r[r]) == 0u);
            if (in) sh.list[atomicAdd(&sh.n_list, 1u)] = k;
        });
        __syncthreads();
    }
    {
        const unsigned int* src = use_list ? sh.list : keys;
        const int nsrc = use_list ? (int)sh.n_list : n;
        refine_ranks([&](auto f) { for (int i = tid; i < nsrc; i += SEL_NT) f(src[i]); }, ranks, nr, sh);
    }
    if (tid < nr) vals[tid] = sh.lo[tid];
    __syncthreads();
}

// numpy percentile(method="linear") on a float32 array, one q, float64 result:
// a + (b-a)*g, or b - (b-a)*(1-g) when g >= 0.5; (b-a) is a float32 subtraction.
__device__ double np_lerp(float a, float b, double g) {
    const double d = (double)__fsub_rn(b, a);
    if (g >= 0.5) return __dsub_rn((double)b, __dmul_rn(d, __dsub_rn(1.0, g)));
    return __dadd_rn((double)a, __dmul_rn(d, g));
}

__device__ __forceinline__ float acc_of(unsigned int S) { return sqrtf((float)S); }       // cv2.magnitude
__device__ __forceinline__ float dist_of(unsigned int t) { return __fmul_rn((float)t, 1.0f / 65536.0f); }

// Integer threshold on S and the cv2.normalize constants of a tile from the two order statistics around the percentile
// rank (numpy lerp weight g), the smallest and the largest S (shared by k_select_grad and the small-tile kernel).
__device__ void grad_params_from_ranks(unsigned int v0, unsigned int v1, double g, unsigned int kmin, unsigned int kmax,
                                       TileParams* out) {
    const double hi = np_lerp(acc_of(v0), acc_of(v1), g);
    // smallest S with float64(acc(S)) >= hi  (acc is monotone in S)
    unsigned int up = 0xffffffffu;              // 0xffffffff = no pixel reaches the threshold
    if ((double)acc_of(0u) >= hi) up = 0u;
    else {
        unsigned int a = 0u, b = v1;            // acc(b) >= hi always holds for the upper order statistic
        if (!((double)acc_of(b) >= hi)) { a = b; b = 0xffffffffu; }
        if (b != 0xffffffffu) {
            while (b - a > 1u) {                // acc(a) < hi <= acc(b)
                const unsigned int m = a + ((b - a) >> 1);
                if ((double)acc_of(m) >= hi) b = m; else a = m;
            }
        }
        up = b;
    }
    out->s_thr = up;
    const double dmin = (double)acc_of(kmin), dmax = (double)acc_of(kmax);
    const double rng = __dsub_rn(dmax, dmin);
    const double scale = (rng > DBL_EPSILON) ? __ddiv_rn(1.0, rng) : 0.0;
    const double shift = __dsub_rn(0.0, __dmul_rn(dmin, scale));
    out->nrm_scale = (float)scale;
    out->nrm_shift = (float)shift;
}

__global__ void __launch_bounds__(SEL_THREADS)
k_select_grad(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles, double q_hi,
              TileParams* __restrict__ params, int sample) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    SelShared& sh = *reinterpret_cast<SelShared*>(sel_smem);
    __shared__ unsigned int ranks[2], vals[2], kmin, kmax;
    const gm_tile t = tiles[blockIdx.x];
    const int n = t.h * t.w;
    __shared__ double g_sh;
    if (threadIdx.x == 0) {
        const double vi = __dmul_rn((double)(n - 1), q_hi);
        const double fl = floor(vi);
        g_sh = __dsub_rn(vi, fl);
        unsigned int lo = (unsigned int)fl;
        if (lo > (unsigned int)(n - 1)) lo = (unsigned int)(n - 1);
        ranks[0] = lo;
        ranks[1] = min(lo + 1u, (unsigned int)(n - 1));
    }
    __syncthreads();
    if (!sample || !sampled_select<1, true, 6>(S + t.px_off, n, ranks, vals, &kmin, &kmax, sh))
        block_select(S + t.px_off, n, ranks, 2, vals, &kmin, &kmax, sh);
    if (threadIdx.x == 0) grad_params_from_ranks(vals[0], vals[1], g_sh, kmin, kmax, &params[blockIdx.x]);
}

// DT_BIN_METHOD = "otsu" (Detect_OBB.py:109-111): replaces k_select_grad.  One CTA per tile, two passes over the
// tile's S: (1) min / max -> the constants of both cv2.normalize calls (0..255 for the 8-bit image Otsu sees, 0..1 for
// the blend of the tail); (2) the 256-bin histogram of acc8 (warp-aggregated shared atomics: gradient images pile up in
// the lowest bins), then thread 0 runs OpenCV's Otsu scan in float64 and turns "acc8 > thr" into the integer
// threshold on S that k_edge_open consumes (dtedge_otsu.cuh; the same functions run on the host in the tests).
__global__ void __launch_bounds__(SEL_THREADS)
k_otsu_grad(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles, TileParams* __restrict__ params) {
    __shared__ unsigned int hist[256 + 1];              // [256]: lanes past the end of the tile
    __shared__ unsigned int wmin[32], wmax[32];
    __shared__ float fs_sh, fh_sh;
    const gm_tile t = tiles[blockIdx.x];
    const int n = t.h * t.w;
    const unsigned int* __restrict__ keys = S + t.px_off;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    unsigned int lo = 0xffffffffu, hi = 0u;
    for (int i = tid; i < n; i += nt) {
        const unsigned int v = keys[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { wmin[warp] = lo; wmax[warp] = hi; }
    for (int i = tid; i <= 256; i += nt) hist[i] = 0u;
    __syncthreads();
    if (warp == 0) {
        lo = lane < nw ? wmin[lane] : 0xffffffffu;
        hi = lane < nw ? wmax[lane] : 0u;
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) {
            wmin[0] = lo; wmax[0] = hi;
            float fs, fh;
            otsu::normalize_constants(acc_of(lo), acc_of(hi), 0.0, 255.0, &fs, &fh);
            fs_sh = fs; fh_sh = fh;
        }
    }
    __syncthreads();
    const float fs = fs_sh, fh = fh_sh;
    const int trips = (n + nt - 1) / nt;                // uniform trip count: every lane takes part in match_any
    for (int k = 0; k < trips; ++k) {
        const int i = k * nt + tid;
        const unsigned int b = i < n ? otsu::acc8_of(keys[i], fs, fh) : 256u;
        const unsigned int peers = __match_any_sync(0xffffffffu, b);
        if (lane == __ffs(peers) - 1) atomicAdd(&hist[b], (unsigned int)__popc(peers));
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned int kmin = wmin[0], kmax = wmax[0];
        const int thr8 = otsu::threshold_from_hist(hist, (long long)n);
        params[blockIdx.x].s_thr = otsu::s_threshold(kmin, kmax, thr8, fs, fh);
        float ns, nh;
        otsu::normalize_constants(acc_of(kmin), acc_of(kmax), 0.0, 1.0, &ns, &nh);
        params[blockIdx.x].nrm_scale = ns;
        params[blockIdx.x].nrm_shift = nh;
    }
}

__device__ void dist_params_from_ranks(const unsigned int* vals, const double* g, TileParams* out) {
    const double p1 = np_lerp(dist_of(vals[0]), dist_of(vals[1]), g[0]);
    const double p99 = np_lerp(dist_of(vals[2]), dist_of(vals[3]), g[1]);
    out->dist_lo = p1;
    const double span = __dsub_rn(p99, p1);
    const double den = span > 1e-6 ? span : 1e-6;
    out->dist_den = den;
    out->inv_den = (float)(1.0 / den);
}

__global__ void __launch_bounds__(SEL_THREADS)
k_select_dist(const unsigned int* __restrict__ T, const gm_tile* __restrict__ tiles,
              TileParams* __restrict__ params, int sample) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    SelShared& sh = *reinterpret_cast<SelShared*>(sel_smem);
    __shared__ unsigned int ranks[4], vals[4], kmin, kmax;
    __shared__ double g_sh[2];
    const gm_tile t = tiles[blockIdx.x];
    const int n = t.h * t.w;
    if (threadIdx.x == 0) {
        const double qs[2] = {1.0 / 100.0, 99.0 / 100.0};
        for (int i = 0; i < 2; ++i) {
            const double vi = __dmul_rn((double)(n - 1), qs[i]);
            const double fl = floor(vi);
            g_sh[i] = __dsub_rn(vi, fl);
            unsigned int lo = (unsigned int)fl;
            if (lo > (unsigned int)(n - 1)) lo = (unsigned int)(n - 1);
            ranks[2 * i] = lo;
            ranks[2 * i + 1] = min(lo + 1u, (unsigned int)(n - 1));
        }
    }
    __syncthreads();
    if (!sample || !sampled_select<2, false, GM_SEL_DIST_SIGMAS>(T + t.px_off, n, ranks, vals, &kmin, &kmax, sh))
        block_select(T + t.px_off, n, ranks, 4, vals, &kmin, &kmax, sh);
    if (threadIdx.x == 0) dist_params_from_ranks(vals, g_sh, &params[blockIdx.x]);
}

// ------------------------------------------------------------------------------------------
// k_edge_open: threshold + 3x3 cross open on bit rows.  A CTA owns EO_ROWS rows
// of one tile.  Phase A streams S once: a warp turns 32 consecutive pixels into one word of the raw
// edge mask with a coalesced 128-byte load, a compare and a ballot (4 words in flight per warp),
// for the CTA's rows plus two halo rows on each side.  Phases B and C are word-parallel in shared
// memory: erosion (out-of-image neighbours count as set) and dilation (they count as clear) are
// shifts with carries from the neighbouring words, ANDs and ORs.  ~0.3 warp instructions per pixel.

constexpr int EO_ROWS = 32;
constexpr int EO_THREADS = 256;
constexpr int EO_MAXW = GM_MAX_TILE / 32;       // words per bit row

__global__ void __launch_bounds__(EO_THREADS)
k_edge_open(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles,
            const TileParams* __restrict__ params, int morph_open, int max_tile, int tile_base,
            unsigned int* __restrict__ zbits) {
    // column 0 and column wpr+1 of every row are the virtual words left / right of the tile
    __shared__ unsigned int E[EO_ROWS + 4][EO_MAXW + 2];
    __shared__ unsigned int Er[EO_ROWS + 2][EO_MAXW + 2];
    const gm_tile t = tiles[blockIdx.x];
    const int y_first = blockIdx.y * EO_ROWS;
    if (y_first >= t.h) return;
    const int rows = min(EO_ROWS, t.h - y_first);
    const int wpr = (t.w + 31) >> 5;
    const int lane = gm_lane();
    const int warp = threadIdx.x >> 5;
    const unsigned int thr = params[blockIdx.x].s_thr;
    const unsigned int* St = S + t.px_off;
    const unsigned int tail_mask = (t.w & 31) ? ((1u << (t.w & 31)) - 1u) : 0xffffffffu;   // valid bits of the last word

    // ---- phase A: E[r][1 + cw] for tile rows y_first - 2 + r, r in [0, rows + 4); outside the tile = all ones.
    // A warp takes whole rows (no per-word index arithmetic) and walks their words four at a time.
    if (((t.w & 3) == 0) && ((t.px_off & 3) == 0)) {
        // 16-byte rows: a lane loads 4 pixels, a warp 128 pixels per instruction and all of a row's loads are in
        // flight together (4x the bytes in flight of the word-per-lane walk below - the phase is latency bound).
        // The 4 compare bits of a lane are a nibble of word lane / 8; 8 lanes OR their nibbles together.
        const int nvr = (t.w + 127) >> 7;                         // 128-pixel groups per row (<= GM_MAX_TILE / 128)
        for (int r = warp; r < rows + 4; r += EO_THREADS / 32) {
            const int y = y_first - 2 + r;
            const bool row_ok = y >= 0 && y < t.h;
            const uint4* srow = reinterpret_cast<const uint4*>(St + (long long)(row_ok ? y : 0) * t.w) + lane;
            for (int k0 = 0; k0 < nvr; k0 += 4) {
                uint4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int x = ((k0 + k) << 7) + 4 * lane;
                    v[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);   // "edge" outside the tile
                    if (row_ok && x < t.w) v[k] = srow[(k0 + k) << 5];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    unsigned int nib = (v[k].x >= thr ? 1u : 0u) | (v[k].y >= thr ? 2u : 0u) | (v[k].z >= thr ? 4u : 0u) | (v[k].w >= thr ? 8u : 0u);
                    nib <<= 4 * (lane & 7);
                    nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
                    nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
                    nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
                    const int c = ((k0 + k) << 2) + (lane >> 3);
                    if ((lane & 7) == 0 && c < wpr) E[r][1 + c] = nib;
                }
            }
        }
    } else
    for (int r = warp; r < rows + 4; r += EO_THREADS / 32) {
        const int y = y_first - 2 + r;
        const bool row_ok = y >= 0 && y < t.h;
        const unsigned int* srow = St + (long long)(row_ok ? y : 0) * t.w + lane;
        for (int c0 = 0; c0 < wpr; c0 += 4) {
            unsigned int v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = ((c0 + k) << 5) + lane;
                v[k] = 0xffffffffu;                               // "edge" for everything outside the tile
                if (row_ok && x < t.w) v[k] = srow[(c0 + k) << 5];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned int bits = __ballot_sync(0xffffffffu, v[k] >= thr);
                if (lane == 0 && c0 + k < wpr) E[r][1 + c0 + k] = bits;
            }
        }
    }
    for (int r = threadIdx.x; r < rows + 4; r += EO_THREADS) { E[r][0] = 0xffffffffu; E[r][wpr + 1] = 0xffffffffu; }
    __syncthreads();

    unsigned int* zt = zbits + zbits_offset(t.px_off, tile_base + (int)blockIdx.x, max_tile);
    if (morph_open <= 0) {
        for (int i = threadIdx.x; i < rows * wpr; i += EO_THREADS) {
            const int r = i / wpr, c = i - r * wpr;
            unsigned int e = E[r + 2][1 + c];
            if (c == wpr - 1) e &= tail_mask;
            zt[(long long)(y_first + r) * wpr + c] = e;
        }
        return;
    }
    // ---- phase B: erosion of rows y_first - 1 .. y_first + rows (Er row r <-> tile row y_first - 1 + r); outside = clear
    for (int i = threadIdx.x; i < (rows + 2) * wpr; i += EO_THREADS) {
        const int r = i / wpr, c = i - r * wpr;
        const int y = y_first - 1 + r;
        unsigned int er = 0u;
        if (y >= 0 && y < t.h) {
            const unsigned int m = E[r + 1][1 + c], l = E[r + 1][c], rt = E[r + 1][2 + c];
            er = m & ((m << 1) | (l >> 31)) & ((m >> 1) | (rt << 31)) & E[r][1 + c] & E[r + 2][1 + c];
            if (c == wpr - 1) er &= tail_mask;
        }
        Er[r][1 + c] = er;
    }
    for (int r = threadIdx.x; r < rows + 2; r += EO_THREADS) { Er[r][0] = 0u; Er[r][wpr + 1] = 0u; }
    __syncthreads();
    // ---- phase C: dilation -> the opened edge mask = zero set of the distance transform
    for (int i = threadIdx.x; i < rows * wpr; i += EO_THREADS) {
        const int r = i / wpr, c = i - r * wpr;
        const unsigned int m = Er[r + 1][1 + c], l = Er[r + 1][c], rt = Er[r + 1][2 + c];
        unsigned int op = m | (m << 1) | (l >> 31) | (m >> 1) | (rt << 31) | Er[r][1 + c] | Er[r + 2][1 + c];
        if (c == wpr - 1) op &= tail_mask;
        zt[(long long)(y_first + r) * wpr + c] = op;
    }
}

// DT_MORPH_OPEN = n >= 2 (Detect_OBB.py:116-118): cv2 runs n erosions and THEN n dilations with the 3x3 cross (an opening
// applied n times would be the opening itself).  Same band decomposition as k_edge_open with a halo of 2 n rows; the two
// buffers ping-pong through the 2 n steps.  Out-of-image neighbours count as set during the erosions and as clear during
// the dilations, in rows (outside the tile), in the bits of the last word beyond the tile width and in the two virtual
// words beside every row.  Not a tuned path: the reference's configured value is 1.
constexpr int EO_MAX_ITER = 8;

__global__ void __launch_bounds__(EO_THREADS)
k_edge_open_iter(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles,
                 const TileParams* __restrict__ params, int n_iter, int max_tile, int tile_base,
                 unsigned int* __restrict__ zbits) {
    __shared__ unsigned int buf[2][EO_ROWS + 4 * EO_MAX_ITER][EO_MAXW + 2];
    const gm_tile t = tiles[blockIdx.x];
    const int y_first = blockIdx.y * EO_ROWS;
    if (y_first >= t.h) return;
    const int rows = min(EO_ROWS, t.h - y_first);
    const int wpr = (t.w + 31) >> 5;
    const int lane = gm_lane();
    const int warp = threadIdx.x >> 5;
    const unsigned int thr = params[blockIdx.x].s_thr;
    const unsigned int* St = S + t.px_off;
    const unsigned int tail_mask = (t.w & 31) ? ((1u << (t.w & 31)) - 1u) : 0xffffffffu;
    const int halo = 2 * n_iter;
    const int R = rows + 2 * halo;                               // buffer row r <-> tile row y_first - halo + r
    // raw edge bits; everything outside the tile is "set"
    for (int r = warp; r < R; r += EO_THREADS / 32) {
        const int y = y_first - halo + r;
        const bool row_ok = y >= 0 && y < t.h;
        const unsigned int* srow = St + (long long)(row_ok ? y : 0) * t.w + lane;
        for (int c = 0; c < wpr; ++c) {
            const int x = (c << 5) + lane;
            unsigned int v = 0xffffffffu;
            if (row_ok && x < t.w) v = srow[c << 5];
            const unsigned int bits = __ballot_sync(0xffffffffu, v >= thr);
            if (lane == 0) buf[0][r][1 + c] = bits;
        }
    }
    for (int r = threadIdx.x; r < R; r += EO_THREADS) { buf[0][r][0] = 0xffffffffu; buf[0][r][wpr + 1] = 0xffffffffu; }
    __syncthreads();
    int cur = 0;
    for (int step = 1; step <= 2 * n_iter; ++step) {
        const bool erode = step <= n_iter;
        const bool last_erode = step == n_iter;
        const unsigned int outside = (erode && !last_erode) ? 0xffffffffu : 0u;   // what the NEXT step sees outside the tile
        unsigned int (*src)[EO_MAXW + 2] = buf[cur];
        unsigned int (*dst)[EO_MAXW + 2] = buf[cur ^ 1];
        // rows [step, R - step) are defined after this step
        for (int i = threadIdx.x; i < (R - 2 * step) * wpr; i += EO_THREADS) {
            const int r = step + i / wpr, c = i % wpr;
            const int y = y_first - halo + r;
            unsigned int o = outside;
            if (y >= 0 && y < t.h) {
                const unsigned int m = src[r][1 + c], l = src[r][c], rt = src[r][2 + c];
                if (erode) o = m & ((m << 1) | (l >> 31)) & ((m >> 1) | (rt << 31)) & src[r - 1][1 + c] & src[r + 1][1 + c];
                else o = m | (m << 1) | (l >> 31) | (m >> 1) | (rt << 31) | src[r - 1][1 + c] | src[r + 1][1 + c];
                if (c == wpr - 1) o = (erode && !last_erode) ? (o | ~tail_mask) : (o & tail_mask);
            }
            dst[r][1 + c] = o;
        }
        for (int r = threadIdx.x; r < R; r += EO_THREADS) { dst[r][0] = outside; dst[r][wpr + 1] = outside; }
        __syncthreads();
        cur ^= 1;
    }
    unsigned int* zt = zbits + zbits_offset(t.px_off, tile_base + (int)blockIdx.x, max_tile);
    for (int i = threadIdx.x; i < rows * wpr; i += EO_THREADS) {
        const int r = i / wpr, c = i - r * wpr;
        zt[(long long)(y_first + r) * wpr + c] = buf[cur][halo + r][1 + c];
    }
}

// ------------------------------------------------------------------------------------------
// k_chamfer: cv2.distanceTransform(DIST_L2, 3) = two raster passes of a 3x3 chamfer mask in 16.16
// fixed point.  One CTA of NW warps owns a tile; a lane holds 4 consecutive columns, a warp 128.
// Forward row:  c[j] = min(prev[j]+HV, prev[j-1]+DG, prev[j+1]+DG)  (0 at zero pixels), then the
// min-plus scan f[j] = min_{k<=j} c[k] + (j-k)*HV.  In the shifted coordinate g[j] = f[j] - j*HV the
// scan is a plain prefix-min and the three neighbour terms get lane-independent constants:
//     cs[j] = min(g'[j] + HV, g'[j-1] + DG - HV, g'[j+1] + DG + HV),   g[j] = min_{k<=j} cs[k]
// (3 adds + one 3-input min per pixel, a 3-step local scan, a 5-shuffle warp scan, one CTA barrier
// to pass the warp totals).  The left neighbour g[j-1] is the exclusive prefix itself; the right one
// needs only the unscanned cs of the next lane.  The backward pass mirrors it with g[j] = b[j] + j*HV
// and a suffix-min.  Rows of the zero mask / forward field are prefetched 4 rows ahead in registers.
// Out-of-image is BIG; any final value >= 2^29 marks a tile without a zero pixel, for which cv2
// saturates to DIST_MAX = UINT_MAX - DG.

constexpr int CH_HV = 62587;
constexpr int CH_DG = 89738;
constexpr int CH_BIG = 1 << 30;
constexpr int CH_INF = 1 << 29;
constexpr unsigned int CH_DIST_MAX = 0xffffffffu - (unsigned int)CH_DG;

__device__ __forceinline__ int min3i(int a, int b, int c) { return __vimin3_s32(a, b, c); }

// Columns at or beyond the tile width take part in the scans as ordinary non-zero pixels of a wider
// image: the 3x3 chamfer value is a shortest-path length on the 8-connected grid, a shortest path
// between two pixels of the tile never needs to leave their bounding box, and whatever the padding
// columns hold is an upper bound of their own distance - so they never lower a value inside the
// tile, and the compute path needs no per-column masks (only loads and stores are masked).
#ifndef GM_CHAMFER_MINB
#define GM_CHAMFER_MINB 12                // resident CTAs per SM asked of the register allocator: 98 -> 80 registers for <2,8,4,16>, no spills
#endif                                    // (measured on B200: c3 0.373 -> 0.365 ms, c5 step 7.82 -> 7.63 ms; 16 CTAs = 64 registers spill and lose: 0.403 / 8.97)
constexpr int chamfer_minb(int nw) { return GM_CHAMFER_MINB * nw * 32 <= 2048 ? GM_CHAMFER_MINB : 2048 / (nw * 32); }
template <int NW, int PX, int AHEAD = 0, int L2AHEAD = 0>
__global__ void __launch_bounds__(NW * 32, chamfer_minb(NW))
k_chamfer(const gm_tile* __restrict__ tiles, int max_tile, int tile_base, const unsigned int* __restrict__ zbits,
          unsigned int* __restrict__ T) {
    static_assert(PX == 4 || PX == 8 || PX == 16, "a lane owns 4, 8 or 16 consecutive columns");
    constexpr int CH_AHEAD = AHEAD > 0 ? AHEAD : ((PX >= 16) ? 2 : 4);      // rows prefetched ahead of the scan
    constexpr int NV = PX / 4;                        // 16-byte vectors per lane and row
    // [row parity][0: warp totals, 1: unscanned value of each warp's edge column][warp]
    __shared__ __align__(16) int xch[2][2][(NW + 3) & ~3];
    const int ti = blockIdx.x;
    const gm_tile t = tiles[ti];
    const int lane = gm_lane();
    const int wq = threadIdx.x >> 5;
    const int w = t.w, h = t.h;
    const int wpr = (w + 31) >> 5;
    const int j0 = wq * (32 * PX) + lane * PX;           // first column of this lane
    const bool vec = ((w & 3) == 0) && ((t.px_off & 3) == 0);
    const int n_ok = min(PX, max(0, w - j0));            // valid columns of this lane
    const int zword = j0 >> 5;                           // word of the row's bit mask holding this lane's bits
    const int zshift = j0 & 31;
    const bool zok = zword < wpr;
    const unsigned int* zrow = zbits + zbits_offset(t.px_off, tile_base + ti, max_tile) + (zok ? zword : 0);
    unsigned int* Trow = T + t.px_off + min(j0, max(w - 1, 0));      // clamped so the pointer stays inside the tile
    int neg_j[PX];                                       // value of a zero pixel in the shifted coordinate
#pragma unroll
    for (int e = 0; e < PX; ++e) neg_j[e] = -(j0 + e) * CH_HV;

    // ================= forward: top -> bottom, prefix-min over g[j] = f[j] - j*HV
    int g[PX];
#pragma unroll
    for (int e = 0; e < PX; ++e) g[e] = CH_BIG;
    int g_left = CH_BIG, g_right = CH_BIG;               // previous row's neighbours of columns j0-1 / j0+PX
    unsigned int zq[CH_AHEAD];
#pragma unroll
    for (int k = 0; k < CH_AHEAD; ++k) zq[k] = (zok && k < h) ? zrow[(long long)k * wpr] : 0u;
    const unsigned int* zpre = zrow + (long long)CH_AHEAD * wpr;
    unsigned int* Tst = Trow;
    // The ring of prefetched rows is rotated by unrolling the row loop CH_AHEAD times (static slots).
    for (int yb = 0; yb < h; yb += CH_AHEAD) {
#pragma unroll
      for (int slot = 0; slot < CH_AHEAD; ++slot) {
        const int y = yb + slot;
        if (y >= h) break;
        const unsigned int zc = zq[slot] >> zshift;
        zq[slot] = (zok && y + CH_AHEAD < h) ? *zpre : 0u;
        zpre += wpr;
        int cs[PX];
#pragma unroll
        for (int e = 0; e < PX; ++e) {
            const int l = (e == 0) ? g_left : g[e - 1];
            const int r = (e == PX - 1) ? g_right : g[e + 1];
            const int c = min3i(g[e] + CH_HV, l + (CH_DG - CH_HV), r + (CH_DG + CH_HV));
            cs[e] = ((zc >> e) & 1u) ? neg_j[e] : c;
        }
        const int cs_next = __shfl_down_sync(0xffffffffu, cs[0], 1);       // unscanned first column of lane+1
        int pm[PX];
        pm[0] = cs[0];
#pragma unroll
        for (int e = 1; e < PX; ++e) pm[e] = min(pm[e - 1], cs[e]);
        int incl = pm[PX - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) incl = min(incl, __shfl_up_sync(0xffffffffu, incl, d));   // lanes < d get their own value back
        int excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = CH_BIG;
        int carry = CH_BIG, next_first = CH_BIG;
        if (NW > 1) {
            const int par = y & 1;
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) { xch[par][0][wq] = total; xch[par][1][wq] = cs[0]; }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < NW - 1; ++q) carry = (q < wq) ? min(carry, xch[par][0][q]) : carry;
            next_first = (wq + 1 < NW) ? xch[par][1][min(wq + 1, NW - 1)] : CH_BIG;
        }
        const int before = min(excl, carry);               // == g[j0-1] of this row
#pragma unroll
        for (int e = 0; e < PX; ++e) g[e] = min(pm[e], before);
        g_left = before;
        // g[j0+PX] of this row = min(everything up to j0+PX-1, unscanned cs of column j0+PX)
        g_right = min(g[PX - 1], (lane == 31) ? next_first : cs_next);
        if (vec) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (4 * v < n_ok)
                    *reinterpret_cast<uint4*>(Tst + 4 * v) = make_uint4((unsigned int)(g[4 * v + 0] - neg_j[4 * v + 0]), (unsigned int)(g[4 * v + 1] - neg_j[4 * v + 1]),
                                                                       (unsigned int)(g[4 * v + 2] - neg_j[4 * v + 2]), (unsigned int)(g[4 * v + 3] - neg_j[4 * v + 3]));
        } else {
#pragma unroll
            for (int e = 0; e < PX; ++e) if (e < n_ok) Tst[e] = (unsigned int)(g[e] - neg_j[e]);
        }
        Tst += w;
      }
    }
    __syncthreads();

    // ================= backward: bottom -> top, suffix-min over g[j] = b[j] + j*HV
#pragma unroll
    for (int e = 0; e < PX; ++e) g[e] = CH_BIG;
    g_left = CH_BIG; g_right = CH_BIG;
    uint4 fq[CH_AHEAD][NV];
    auto load_row = [&](const unsigned int* row, bool in_range, uint4 (&dst)[NV]) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            uint4 val = make_uint4(CH_BIG, CH_BIG, CH_BIG, CH_BIG);
            if (in_range) {
                if (vec) { if (4 * v < n_ok) val = *reinterpret_cast<const uint4*>(row + 4 * v); }
                else {
                    if (4 * v + 0 < n_ok) val.x = row[4 * v + 0];
                    if (4 * v + 1 < n_ok) val.y = row[4 * v + 1];
                    if (4 * v + 2 < n_ok) val.z = row[4 * v + 2];
                    if (4 * v + 3 < n_ok) val.w = row[4 * v + 3];
                }
            }
            dst[v] = val;
        }
    };
    unsigned int* Tcur = Trow + (long long)(h - 1) * w;
#pragma unroll
    for (int k = 0; k < CH_AHEAD; ++k) load_row(Tcur - (long long)k * w, h - 1 - k >= 0, fq[k]);
    const unsigned int* Tpre = Tcur - (long long)CH_AHEAD * w;
    for (int yb = h - 1; yb >= 0; yb -= CH_AHEAD) {
#pragma unroll
      for (int slot = 0; slot < CH_AHEAD; ++slot) {
        const int y = yb - slot;
        if (y < 0) break;
        int f[PX];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            f[4 * v + 0] = (int)fq[slot][v].x; f[4 * v + 1] = (int)fq[slot][v].y;
            f[4 * v + 2] = (int)fq[slot][v].z; f[4 * v + 3] = (int)fq[slot][v].w;
        }
        load_row(Tpre, y - CH_AHEAD >= 0, fq[slot]);
        if (L2AHEAD > 0 && y - L2AHEAD >= 0 && n_ok > 0) {
            // The register ring cannot hide DRAM latency on its own: a warp has six scoreboards, loads of several
            // rows share one, and waiting for the oldest row then waits for the youngest as well.  So rows far
            // ahead are pulled into L2 (no scoreboard) and the ring only has to cover an L2 hit: 0.419 -> 0.383 ms.
            // (A cp.async ring in shared memory, 5-12 rows deep, measured the same 0.383-0.394 ms and was dropped.)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(Tcur - (long long)L2AHEAD * w));
        }
        Tpre -= w;
        int cs[PX];
#pragma unroll
        for (int e = 0; e < PX; ++e) {
            const int l = (e == 0) ? g_left : g[e - 1];
            const int r = (e == PX - 1) ? g_right : g[e + 1];
            const int c = min3i(g[e] + CH_HV, l + (CH_DG + CH_HV), r + (CH_DG - CH_HV));
            cs[e] = min(c, f[e] - neg_j[e]);
        }
        const int cs_prev = __shfl_up_sync(0xffffffffu, cs[PX - 1], 1);    // unscanned last column of lane-1
        int pm[PX];
        pm[PX - 1] = cs[PX - 1];
#pragma unroll
        for (int e = PX - 2; e >= 0; --e) pm[e] = min(pm[e + 1], cs[e]);
        int incl = pm[0];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) incl = min(incl, __shfl_down_sync(0xffffffffu, incl, d)); // lanes >= 32-d get their own value back
        int excl = __shfl_down_sync(0xffffffffu, incl, 1);
        if (lane == 31) excl = CH_BIG;
        int carry = CH_BIG, prev_last = CH_BIG;
        if (NW > 1) {
            const int par = y & 1;
            const int total = __shfl_sync(0xffffffffu, incl, 0);
            const int last_cs = __shfl_sync(0xffffffffu, cs[PX - 1], 31);
            if (lane == 0) { xch[par][0][wq] = total; xch[par][1][wq] = last_cs; }
            __syncthreads();
#pragma unroll
            for (int q = 1; q < NW; ++q) carry = (q > wq) ? min(carry, xch[par][0][q]) : carry;
            prev_last = (wq > 0) ? xch[par][1][max(wq - 1, 0)] : CH_BIG;
        }
        const int after = min(excl, carry);                // == g[j0+PX] of this row
#pragma unroll
        for (int e = 0; e < PX; ++e) g[e] = min(pm[e], after);
        g_right = after;
        g_left = min(g[0], (lane == 0) ? prev_last : cs_prev);
        unsigned int o[PX];
#pragma unroll
        for (int e = 0; e < PX; ++e) {
            const int v = g[e] + neg_j[e];
            o[e] = (v >= CH_INF) ? CH_DIST_MAX : (unsigned int)v;
        }
        if (vec) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
                if (4 * v < n_ok) *reinterpret_cast<uint4*>(Tcur + 4 * v) = make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        } else {
#pragma unroll
            for (int e = 0; e < PX; ++e) if (e < n_ok) Tcur[e] = o[e];
        }
        Tcur -= w;
      }
    }
}

// ------------------------------------------------------------------------------------------
// k_tail: normalise, exp(-d/3), blend, truncate, pack.  One thread owns 4 consecutive pixels of a
// tile row: 16-byte loads of S and of the chamfer field, the 12 map bytes from aligned 32-bit words,
// one 16-byte store of the four [R,G,B,DT] pixels (HWC) or four 4-byte plane stores (CHW).
//
// The byte is trunc(255 * clip(0.7*exp(-d/3) + 0.3*nrm)) with d, exp and the blend in float64
// (numpy 2 semantics).  An fp32 estimate (approximate sqrt / exp2, float-float p1)
// is within ~1e-4 of the float64 value (sqrt.approx 2^-23 and ex2.approx 2^-22 relative, the fp32 0.7 and
// the final * 255 rounding), so it already decides the truncation unless it lands within 6e-4 of an
// integer; only those pixels (~0.1 %) are redone on the exact float64 path.

constexpr int TAIL_THREADS = 256;
#ifndef GM_TAIL_ROWS
#define GM_TAIL_ROWS 64
#endif
constexpr int TAIL_ROWS = GM_TAIL_ROWS;   // rows of one tile per CTA (grid.y covers max_tile / TAIL_ROWS); measured on 416-px tiles: 8 rows 0.409, 16 0.379, 32 0.365, 64 0.354 ms

__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exact byte (float64 where numpy computes in float64)
__device__ __noinline__ unsigned int tail_byte_exact(unsigned int S, unsigned int T, const TileParams& p) {
    const float nrm = __fmaf_rn(acc_of(S), p.nrm_scale, p.nrm_shift);
    const float g03 = __fmul_rn(0.3f, nrm);
    double d = __ddiv_rn(__dsub_rn((double)dist_of(T), p.dist_lo), p.dist_den);
    d = fmin(fmax(d, 0.0), 1.0);
    const double soft = exp(__ddiv_rn(-d, 3.0));
    double v = __dadd_rn(__dmul_rn(0.7, soft), (double)g03);
    v = fmin(fmax(v, 0.0), 1.0);
    return (unsigned int)__dmul_rn(v, 255.0);      // truncation like astype(uint8)
}

// fp32 estimate; returns the byte, sets `unsure` when the estimate is too close to an integer
__device__ __forceinline__ unsigned int tail_byte_fast(unsigned int S, unsigned int T, const TileParams& p,
                                                       float lo_hi, float lo_lo, bool& unsure) {
    const float nrm = fmaf(sqrt_approx((float)S), p.nrm_scale, p.nrm_shift);
    const float diff = (dist_of(T) - lo_hi) - lo_lo;
    const float d = __saturatef(diff * p.inv_den);
    const float e = ex2_approx(d * (-1.4426950408889634f / 3.f));
    const float q = __saturatef(fmaf(0.7f, e, 0.3f * nrm)) * 255.f;
    const float r = (q + 8388608.f) - 8388608.f;          // nearest integer
    const float fr = q - r;                               // in [-0.5, 0.5]
    unsure = fabsf(fr) < 6e-4f;
    return (unsigned int)(int)r - (fr < 0.f ? 1u : 0u);
}

struct TailIn {                  // everything one 4-pixel group needs from memory
    uint4 s, t;                  // S and chamfer field of the 4 pixels
    unsigned int v0, v1, v2;     // their 12 map bytes B G R B G R ...
    int i0;                      // pixel index of the group inside the tile
    int cnt;                     // pixels of the group inside the row (1..4)
    bool vec;                    // 16-byte aligned full group
};

// Rows [y_first, y_first + n_rows) of tile t by the calling CTA (thread `tid` of `nt`); shared by k_tail and the
// small-tile kernel that runs the tail right behind its percentile selection.
__device__ __forceinline__ void tail_rows(const uint8_t* __restrict__ map, int W, long long map_bytes, const gm_tile& t,
                                          const TileParams& p, const unsigned int* __restrict__ S,
                                          const unsigned int* __restrict__ T, int layout, uint8_t* __restrict__ out,
                                          int y_first, int n_rows, int tid, int nt) {
    const int gpr = (t.w + 3) >> 2;                        // 4-pixel groups per row
    const int n_groups = n_rows * gpr;
    const float lo_hi = (float)p.dist_lo;
    const float lo_lo = (float)(p.dist_lo - (double)lo_hi);
    const long long n = (long long)t.h * t.w;
    // g / gpr by multiply-high: exact for g < 2^16 and gpr <= GM_MAX_TILE / 4 (a single-group row divides by 1)
    const unsigned int magic = gpr > 1 ? 0xffffffffu / (unsigned int)gpr + 1u : 0u;
    const unsigned int* Sb = S + t.px_off;
    const unsigned int* Tb = T + t.px_off;

    auto load = [&](int g, TailIn& in) {
        const int yy = gpr > 1 ? (int)__umulhi((unsigned int)g, magic) : g;
        const int x = (g - yy * gpr) << 2;
        const int y = y_first + yy;
        in.cnt = min(4, t.w - x);
        in.i0 = y * t.w + x;
        const unsigned int* Sp = Sb + in.i0;
        const unsigned int* Tp = Tb + in.i0;
        in.vec = in.cnt == 4 && ((reinterpret_cast<unsigned long long>(Sp) & 15ULL) == 0ULL);
        if (in.vec) {
            in.s = *reinterpret_cast<const uint4*>(Sp);
            in.t = *reinterpret_cast<const uint4*>(Tp);
        } else {
            in.s = make_uint4(Sp[0], in.cnt > 1 ? Sp[1] : 0u, in.cnt > 2 ? Sp[2] : 0u, in.cnt > 3 ? Sp[3] : 0u);
            in.t = make_uint4(Tp[0], in.cnt > 1 ? Tp[1] : 0u, in.cnt > 2 ? Tp[2] : 0u, in.cnt > 3 ? Tp[3] : 0u);
        }
        const long long a_off = ((long long)(t.y0 + y) * W + (t.x0 + x)) * 3LL;
        if (in.cnt == 4 && a_off + 16 <= map_bytes) {
            const unsigned long long sa = reinterpret_cast<unsigned long long>(map + a_off);
            const unsigned int* sw = reinterpret_cast<const unsigned int*>(sa & ~3ULL);
            const unsigned int sh = (unsigned int)(sa & 3ULL) * 8u;
            const unsigned int w0 = __ldg(sw), w1 = __ldg(sw + 1), w2 = __ldg(sw + 2), w3 = __ldg(sw + 3);
            in.v0 = __funnelshift_r(w0, w1, sh); in.v1 = __funnelshift_r(w1, w2, sh); in.v2 = __funnelshift_r(w2, w3, sh);
        } else {
            unsigned int by[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) by[k] = (k < 3 * in.cnt) ? (unsigned int)__ldg(map + a_off + k) : 0u;
            in.v0 = by[0] | (by[1] << 8) | (by[2] << 16) | (by[3] << 24);
            in.v1 = by[4] | (by[5] << 8) | (by[6] << 16) | (by[7] << 24);
            in.v2 = by[8] | (by[9] << 8) | (by[10] << 16) | (by[11] << 24);
        }
    };

    auto finish = [&](const TailIn& in) {
        const unsigned int sv[4] = {in.s.x, in.s.y, in.s.z, in.s.w}, tv[4] = {in.t.x, in.t.y, in.t.z, in.t.w};
        const unsigned int v0 = in.v0, v1 = in.v1, v2 = in.v2;
        const int cnt = in.cnt;
        unsigned int dt[4];
        unsigned int redo = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            bool unsure;
            dt[e] = tail_byte_fast(sv[e], tv[e], p, lo_hi, lo_lo, unsure);
            if (unsure) redo |= 1u << e;
        }
        while (redo) {
            const int e = __ffs(redo) - 1;
            redo &= redo - 1u;
            const unsigned int v = tail_byte_exact(e == 0 ? sv[0] : e == 1 ? sv[1] : e == 2 ? sv[2] : sv[3],
                                                   e == 0 ? tv[0] : e == 1 ? tv[1] : e == 2 ? tv[2] : tv[3], p);
            if (e == 0) dt[0] = v; else if (e == 1) dt[1] = v; else if (e == 2) dt[2] = v; else dt[3] = v;
        }
        if (layout == 0) {
            // pixel e -> R | G<<8 | B<<16 | DT<<24
            uint4 o;
            o.x = __byte_perm(v0, dt[0], 0x4012);
            o.y = __byte_perm(__byte_perm(v0, v1, 0x0345), dt[1], 0x4210);
            o.z = __byte_perm(__byte_perm(v1, v2, 0x0234), dt[2], 0x4210);
            o.w = __byte_perm(v2, dt[3], 0x4123);
            unsigned int* dst = reinterpret_cast<unsigned int*>(out) + t.px_off + in.i0;
            if (in.vec) *reinterpret_cast<uint4*>(dst) = o;
            else {
                dst[0] = o.x;
                if (cnt > 1) dst[1] = o.y;
                if (cnt > 2) dst[2] = o.z;
                if (cnt > 3) dst[3] = o.w;
            }
        } else {
            // planes R, G, B, DT of n pixels each
            const unsigned int rr = __byte_perm(__byte_perm(v0, v1, 0x0052), v2, 0x7410);   // R0 R1 R2 R3
            const unsigned int gg = __byte_perm(__byte_perm(v0, v1, 0x0741), v2, 0x6210);   // G0 G1 G2 G3
            const unsigned int bb = __byte_perm(__byte_perm(v0, v1, 0x0630), v2, 0x5210);   // B0 B1 B2 B3
            const unsigned int dd = dt[0] | (dt[1] << 8) | (dt[2] << 16) | (dt[3] << 24);
            uint8_t* o = out + 4LL * t.px_off + in.i0;
            const unsigned int planes[4] = {rr, gg, bb, dd};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint8_t* q = o + (long long)c * n;
                if (cnt == 4 && ((reinterpret_cast<unsigned long long>(q) & 3ULL) == 0ULL)) {
                    *reinterpret_cast<unsigned int*>(q) = planes[c];
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (e < cnt) q[e] = (uint8_t)(planes[c] >> (8 * e));
                }
            }
        }
    };

    // one group ahead: the loads of group g + nt are in flight while group g is finished
    int g = tid;
    if (g >= n_groups) return;
    TailIn cur;
    load(g, cur);
    for (;;) {
        const int gn = g + nt;
        TailIn nxt;
        const bool more = gn < n_groups;
        if (more) load(gn, nxt);
        finish(cur);
        if (!more) break;
        cur = nxt;
        g = gn;
    }
}

#ifndef GM_TAIL_MINB
#define GM_TAIL_MINB 4                    // resident CTAs per SM asked of the register allocator: 80 -> 64 registers, no spills; the kernel waits on
#endif                                    // memory (long scoreboard), so 32 instead of 24 warps per SM: c3 0.354 -> 0.311 ms, c5 step 11.26 -> 9.88 ms (48 registers: the same; 40: 12.1)
__global__ void __launch_bounds__(TAIL_THREADS, GM_TAIL_MINB)
k_tail(const uint8_t* __restrict__ map, int W, long long map_bytes, const gm_tile* __restrict__ tiles,
       const TileParams* __restrict__ params, const unsigned int* __restrict__ S,
       const unsigned int* __restrict__ T, int layout, uint8_t* __restrict__ out) {
    const gm_tile t = tiles[blockIdx.x];
    const int y_first = blockIdx.y * TAIL_ROWS;
    if (y_first >= t.h) return;
    const TileParams p = params[blockIdx.x];
    tail_rows(map, W, map_bytes, t, p, S, T, layout, out, y_first, min(TAIL_ROWS, t.h - y_first), (int)threadIdx.x, TAIL_THREADS);
}

// ------------------------------------------------------------------------------------------
// Small tiles (side <= 128: the 128/30 scale of the dual-scale pipeline, 28,224 tiles on a 16384^2 map).  A tile's keys are
// at most 64 KB, so one CTA keeps them in shared memory and the two percentile selections become radix refinements over
// shared memory with their consumer fused behind them:
//   k_select_edge_small  = k_select_grad + k_edge_open : S read once; order statistics, threshold, bit mask, 3x3 cross open
//   k_select_tail_small  = k_select_dist + k_tail      : the chamfer field read once for the percentiles, the tail of the
//                                                         tile right behind them (its reads of the field hit L2)
// The large-tile kernels spend their time on this plan in per-tile fixed costs (a 1024-thread CTA with 131 KB of shared
// memory per 16 k keys, two global passes, one CTA per SM); these run 2 CTAs of 512 threads per SM, one global pass.
constexpr int SMALL_SIDE = 128;
constexpr int SMALL_KEYS = SMALL_SIDE * SMALL_SIDE;
constexpr int SMALL_THREADS = 512;
constexpr int SMALL_WPR = SMALL_SIDE / 32;
static_assert(SMALL_WPR == 4, "the word index of the open phases is tid & 3");

struct SelSmall {                    // the fields refine_ranks uses
    unsigned int hist[SEL_MAXR][SEL_BINS];
    unsigned int lo[SEL_MAXR];
    int rem[SEL_MAXR];
    unsigned int base[SEL_MAXR];
    int hid[SEL_MAXR];
    int bin[SEL_MAXR];
    unsigned int below[SEL_MAXR];
    unsigned int red_min[32], red_max[32];
};

struct SmallShared {
    unsigned int keys[SMALL_KEYS];
    SelSmall sel;
    unsigned int E[SMALL_SIDE + 4][SMALL_WPR + 2];      // raw edge bits, rows -2 .. h+1, one virtual word on each side
    unsigned int Er[SMALL_SIDE + 2][SMALL_WPR + 2];     // eroded bits, rows -1 .. h
};

// keys of the tile -> shared memory, min / max of them; returns after a CTA barrier
__device__ __forceinline__ void small_load_keys(const unsigned int* __restrict__ src, int n, SmallShared& sh,
                                                unsigned int* kmin, unsigned int* kmax) {
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned int mn = 0xffffffffu, mx = 0u;
    // all of a thread's loads of a batch are issued before the first store (a load -> store loop keeps one load in flight
    // per warp: 32 DRAM round trips per tile)
    if ((reinterpret_cast<unsigned long long>(src) & 15ULL) == 0ULL) {
        const uint4* src4 = reinterpret_cast<const uint4*>(src);
        uint4* dst4 = reinterpret_cast<uint4*>(sh.keys);
        const int nv = n >> 2;
        for (int base = 0; base < nv; base += 4 * SMALL_THREADS) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = base + k * SMALL_THREADS + tid;
                v[k] = i < nv ? src4[i] : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = base + k * SMALL_THREADS + tid;
                if (i < nv) {
                    dst4[i] = v[k];
                    mn = min(mn, min(min(v[k].x, v[k].y), min(v[k].z, v[k].w)));
                    mx = max(mx, max(max(v[k].x, v[k].y), max(v[k].z, v[k].w)));
                }
            }
        }
        const int i = 4 * nv + tid;
        if (i < n) { const unsigned int k = src[i]; sh.keys[i] = k; mn = min(mn, k); mx = max(mx, k); }
    } else {
        for (int base = 0; base < n; base += 8 * SMALL_THREADS) {
            unsigned int v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = base + k * SMALL_THREADS + tid;
                v[k] = i < n ? src[i] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = base + k * SMALL_THREADS + tid;
                if (i < n) { sh.keys[i] = v[k]; mn = min(mn, v[k]); mx = max(mx, v[k]); }
            }
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31) == 0) { sh.sel.red_min[warp] = mn; sh.sel.red_max[warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        unsigned int a = 0xffffffffu, b = 0u;
        for (int w = 0; w < SMALL_THREADS / 32; ++w) { a = min(a, sh.sel.red_min[w]); b = max(b, sh.sel.red_max[w]); }
        *kmin = a; *kmax = b;
    }
    __syncthreads();
}

// Exact order statistics `ranks[0..nr)` of the shared-memory keys.  Round 0 is the logarithmic histogram of block_select
// (exponent + 6 mantissa bits: gradient energies and chamfer distances pile up at a few small values, which a plain
// top-bits histogram would send to one counter), with the runs of equal bins a thread meets added in one atomic;
// then the radix refinement inside the bins that hold the ranks - over the same shared-memory keys, no compaction.
__device__ __forceinline__ void small_select(int n, const unsigned int* ranks, int nr, unsigned int* vals, SmallShared& sh) {
    __shared__ unsigned int c_min[SEL_MAXR], c_max[SEL_MAXR], n_list;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned int* hist0 = &sh.sel.hist[0][0];
    for (int i = tid; i < SEL_LOGBINS; i += SMALL_THREADS) hist0[i] = 0u;
    __syncthreads();
    {
        // run-length aggregation per thread: its keys are 512 pixels apart (a few rows), and in flat regions they share a
        // bin - 16 k atomics on ONE shared-memory counter serialise (measured: 28 us per tile with a plain histogram)
        unsigned int prev = 0xffffffffu, cnt = 0u;
        for (int i = tid; i < n; i += SMALL_THREADS) {
            const unsigned int b = (unsigned int)log_bin(sh.keys[i]);
            if (b == prev) ++cnt;
            else {
                if (cnt) atomicAdd(&hist0[prev], cnt);
                prev = b; cnt = 1u;
            }
        }
        if (cnt) atomicAdd(&hist0[prev], cnt);
    }
    __syncthreads();
    if (warp < nr) warp_find_bin(hist0, SEL_LOGBINS, ranks[warp], &sh.sel.bin[warp], &sh.sel.below[warp]);
    __syncthreads();
    if (tid < nr) {
        const int bin = sh.sel.bin[tid];
        const int e = (bin >> 6) - 1;
        const unsigned int m = (unsigned int)(bin & 63);
        if (bin == 0) { sh.sel.lo[tid] = 0u; sh.sel.rem[tid] = 0; }
        else if (e >= 6) { sh.sel.lo[tid] = (64u | m) << (e - 6); sh.sel.rem[tid] = e - 6; }
        else { sh.sel.lo[tid] = (64u | m) >> (6 - e); sh.sel.rem[tid] = 0; }
        sh.sel.base[tid] = sh.sel.below[tid];
        c_min[tid] = 0xffffffffu; c_max[tid] = 0u;
    }
    if (tid == 0) n_list = 0u;
    __syncthreads();
    {
        int any = 0;
        for (int r = 0; r < nr; ++r) any |= sh.sel.rem[r];
        if (!any) {                                               // every rank sits in a single-valued bin
            if (tid < nr) vals[tid] = sh.sel.lo[tid];
            __syncthreads();
            return;
        }
    }
    // The refinement only concerns the keys inside the (<= 4) bins that hold the ranks - about 1 % of the tile.  One scan
    // compacts them into a list (it lives where the bit rows of the open will be: dead until then) and takes the smallest
    // and largest key of every bin: a bin with ONE distinct key is decided without refining (chamfer distances and flat
    // gradients pile thousands of equal keys into a bin).  A list that overflows falls back to refining over all keys.
    unsigned int* list = &sh.E[0][0];
    constexpr unsigned int LIST_CAP = (unsigned int)((sizeof(sh.E) + sizeof(sh.Er)) / sizeof(unsigned int));
    {
        unsigned int lo_r[SEL_MAXR], mn[SEL_MAXR], mx[SEL_MAXR];
        int rem_r[SEL_MAXR];
#pragma unroll
        for (int r = 0; r < SEL_MAXR; ++r) {
            lo_r[r] = r < nr ? sh.sel.lo[r] : 0u;
            rem_r[r] = r < nr ? sh.sel.rem[r] : 0;
            mn[r] = 0xffffffffu; mx[r] = 0u;
        }
        for (int i = tid; i < n; i += SMALL_THREADS) {
            const unsigned int k = sh.keys[i];
            bool in = false;
#pragma unroll
            for (int r = 0; r < SEL_MAXR; ++r) {
                if (rem_r[r] > 0 && k >= lo_r[r] && ((k - lo_r[r]) >> rem_r[r]) == 0u) {
                    in = true;
                    mn[r] = min(mn[r], k); mx[r] = max(mx[r], k);
                }
            }
            if (in) {
                const unsigned int slot = atomicAdd(&n_list, 1u);
                if (slot < LIST_CAP) list[slot] = k;
            }
        }
#pragma unroll
        for (int r = 0; r < SEL_MAXR; ++r) {
            if (r < nr && rem_r[r] > 0) {
                const unsigned int a = __reduce_min_sync(0xffffffffu, mn[r]);
                const unsigned int b = __reduce_max_sync(0xffffffffu, mx[r]);
                if ((tid & 31) == 0) { atomicMin(&c_min[r], a); atomicMax(&c_max[r], b); }
            }
        }
    }
    __syncthreads();
    if (tid < nr && sh.sel.rem[tid] > 0 && c_min[tid] == c_max[tid]) { sh.sel.lo[tid] = c_min[tid]; sh.sel.rem[tid] = 0; }
    __syncthreads();
    const int n_src = (int)n_list;
    if (n_src <= (int)LIST_CAP)
        refine_ranks([&](auto f) { for (int i = tid; i < n_src; i += SMALL_THREADS) f(list[i]); }, ranks, nr, sh.sel);
    else
        refine_ranks([&](auto f) { for (int i = tid; i < n; i += SMALL_THREADS) f(sh.keys[i]); }, ranks, nr, sh.sel);
    if (tid < nr) vals[tid] = sh.sel.lo[tid];
    __syncthreads();
}

__global__ void __launch_bounds__(SMALL_THREADS, 2)
k_select_edge_small(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles, double q_hi,
                    TileParams* __restrict__ params, int morph_open, int max_tile, int tile_base,
                    unsigned int* __restrict__ zbits) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    SmallShared& sh = *reinterpret_cast<SmallShared*>(small_smem);
    __shared__ unsigned int ranks[2], vals[2], kmin, kmax, s_thr;
    __shared__ double g_sh;
    const gm_tile t = tiles[blockIdx.x];
    const int n = t.h * t.w;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const double vi = __dmul_rn((double)(n - 1), q_hi);
        const double fl = floor(vi);
        g_sh = __dsub_rn(vi, fl);
        unsigned int lo = (unsigned int)fl;
        if (lo > (unsigned int)(n - 1)) lo = (unsigned int)(n - 1);
        ranks[0] = lo;
        ranks[1] = min(lo + 1u, (unsigned int)(n - 1));
    }
    small_load_keys(S + t.px_off, n, sh, &kmin, &kmax);
    small_select(n, ranks, 2, vals, sh);
    if (tid == 0) {
        TileParams p = params[blockIdx.x];
        grad_params_from_ranks(vals[0], vals[1], g_sh, kmin, kmax, &p);
        params[blockIdx.x].s_thr = p.s_thr;
        params[blockIdx.x].nrm_scale = p.nrm_scale;
        params[blockIdx.x].nrm_shift = p.nrm_shift;
        s_thr = p.s_thr;
    }
    __syncthreads();
    // ---- threshold + 3x3 cross open on bit rows (the arithmetic of k_edge_open, the whole tile at once)
    const unsigned int thr = s_thr;
    const int h = t.h, w = t.w;
    const int wpr = (w + 31) >> 5;
    const unsigned int tail_mask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    // phase A: E[r][1 + c] for tile rows r - 2; outside the tile = all ones; a warp takes whole rows and makes one word per
    // step by ballot.  (Rows and words are walked as two loops: an index / wpr division per step was a third of the
    // kernel's instructions, ncu.)
    for (int r = warp; r < h + 4; r += SMALL_THREADS / 32) {
        const int y = r - 2;
        const bool row_ok = y >= 0 && y < h;
        for (int c = 0; c < wpr; ++c) {
            const int x = (c << 5) + lane;
            unsigned int v = 0xffffffffu;
            if (row_ok && x < w) v = sh.keys[y * w + x];
            const unsigned int bits = __ballot_sync(0xffffffffu, v >= thr);
            if (lane == 0) sh.E[r][1 + c] = bits;
        }
    }
    for (int r = tid; r < h + 4; r += SMALL_THREADS) { sh.E[r][0] = 0xffffffffu; sh.E[r][wpr + 1] = 0xffffffffu; }
    __syncthreads();
    unsigned int* zt = zbits + zbits_offset(t.px_off, tile_base + (int)blockIdx.x, max_tile);
    // phases B / C: thread = (row tid / 4 + k * 128, word tid % 4); SMALL_WPR == 4 words at most
    const int c = tid & (SMALL_WPR - 1);
    const int r0 = tid >> 2;
    if (morph_open <= 0) {
        if (c < wpr)
            for (int r = r0; r < h; r += SMALL_THREADS / SMALL_WPR) {
                unsigned int e = sh.E[r + 2][1 + c];
                if (c == wpr - 1) e &= tail_mask;
                zt[(long long)r * wpr + c] = e;
            }
        return;
    }
    // phase B: erosion of rows -1 .. h (Er row r <-> tile row r - 1); outside = clear
    if (c < wpr)
        for (int r = r0; r < h + 2; r += SMALL_THREADS / SMALL_WPR) {
            const int y = r - 1;
            unsigned int er = 0u;
            if (y >= 0 && y < h) {
                const unsigned int m = sh.E[r + 1][1 + c], l = sh.E[r + 1][c], rt = sh.E[r + 1][2 + c];
                er = m & ((m << 1) | (l >> 31)) & ((m >> 1) | (rt << 31)) & sh.E[r][1 + c] & sh.E[r + 2][1 + c];
                if (c == wpr - 1) er &= tail_mask;
            }
            sh.Er[r][1 + c] = er;
        }
    for (int r = tid; r < h + 2; r += SMALL_THREADS) { sh.Er[r][0] = 0u; sh.Er[r][wpr + 1] = 0u; }
    __syncthreads();
    // phase C: dilation -> the opened edge mask = zero set of the distance transform
    if (c < wpr)
        for (int r = r0; r < h; r += SMALL_THREADS / SMALL_WPR) {
            const unsigned int m = sh.Er[r + 1][1 + c], l = sh.Er[r + 1][c], rt = sh.Er[r + 1][2 + c];
            unsigned int op = m | (m << 1) | (l >> 31) | (m >> 1) | (rt << 31) | sh.Er[r][1 + c] | sh.Er[r + 2][1 + c];
            if (c == wpr - 1) op &= tail_mask;
            zt[(long long)r * wpr + c] = op;
        }
}

__global__ void __launch_bounds__(SMALL_THREADS, 2)
k_select_tail_small(const uint8_t* __restrict__ map, int W, long long map_bytes, const gm_tile* __restrict__ tiles,
                    TileParams* __restrict__ params, const unsigned int* __restrict__ S,
                    const unsigned int* __restrict__ T, int layout, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    SmallShared& sh = *reinterpret_cast<SmallShared*>(small_smem);
    __shared__ unsigned int ranks[4], vals[4], kmin, kmax;
    __shared__ double g_sh[2];
    __shared__ TileParams p_sh;
    const gm_tile t = tiles[blockIdx.x];
    const int n = t.h * t.w;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const double qs[2] = {1.0 / 100.0, 99.0 / 100.0};
        for (int i = 0; i < 2; ++i) {
            const double vi = __dmul_rn((double)(n - 1), qs[i]);
            const double fl = floor(vi);
            g_sh[i] = __dsub_rn(vi, fl);
            unsigned int lo = (unsigned int)fl;
            if (lo > (unsigned int)(n - 1)) lo = (unsigned int)(n - 1);
            ranks[2 * i] = lo;
            ranks[2 * i + 1] = min(lo + 1u, (unsigned int)(n - 1));
        }
    }
    small_load_keys(T + t.px_off, n, sh, &kmin, &kmax);
    small_select(n, ranks, 4, vals, sh);
    if (tid == 0) {
        TileParams p = params[blockIdx.x];
        dist_params_from_ranks(vals, g_sh, &p);
        params[blockIdx.x].dist_lo = p.dist_lo;
        params[blockIdx.x].dist_den = p.dist_den;
        params[blockIdx.x].inv_den = p.inv_den;
        p_sh = p;
    }
    __syncthreads();
    const TileParams p = p_sh;
    tail_rows(map, W, map_bytes, t, p, S, T, layout, out, 0, t.h, tid, SMALL_THREADS);
}

struct DtWorkspace {
    unsigned int* S;
    unsigned int* T;
    unsigned int* zbits;
    TileParams* params;
    size_t bytes;
};

DtWorkspace carve(void* ws, int64_t total_px, int32_t n_tiles, int32_t max_tile_hint) {
    GmArena a(ws, ~(size_t)0);
    DtWorkspace w;
    w.S = a.take<unsigned int>((size_t)total_px);
    w.T = a.take<unsigned int>((size_t)total_px);
    w.zbits = a.take<unsigned int>((size_t)(total_px / 32 + (int64_t)n_tiles * (max_tile_hint + 1) + 2));
    w.params = a.take<TileParams>((size_t)n_tiles);
    w.bytes = gm_align_up(a.off, 256);
    return w;
}

}  // namespace

extern "C" int gm_dtedge_select_stats(uint32_t* counts8_host, int32_t reset) {
    if (counts8_host) GM_CUDA_TRY(cudaMemcpyFromSymbol(counts8_host, g_sel_stats, 8 * sizeof(uint32_t)));
    if (reset) {
        const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        GM_CUDA_TRY(cudaMemcpyToSymbol(g_sel_stats, z, sizeof(z)));
    }
    return GM_OK;
}

extern "C" size_t gm_dtedge_workspace_bytes(int64_t total_px, int32_t n_tiles) {
    if (total_px < 0 || n_tiles < 0) return 0;
    return carve(nullptr, total_px, n_tiles, GM_MAX_TILE).bytes;
}

extern "C" int gm_dtedge_workspace_views(void* workspace_dev, int64_t total_px, int32_t n_tiles,
                                         uint32_t** S_dev, uint32_t** zero_bits_dev, uint32_t** chamfer_dev) {
    if (!workspace_dev) return GM_EINVAL;
    DtWorkspace w = carve(workspace_dev, total_px, n_tiles, GM_MAX_TILE);
    if (S_dev) *S_dev = w.S;
    if (zero_bits_dev) *zero_bits_dev = w.zbits;
    if (chamfer_dev) *chamfer_dev = w.T;
    return GM_OK;
}

// Tensor map of the BGR map for k_grad_fast<true>: uint8 [H][3 W], box = one block's patch rounded out to 16-byte boundaries (160 bytes x 80 rows), no
// swizzle, zero fill outside.  cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime
// (cudaGetDriverEntryPoint) so the library does not link libcuda.
typedef CUresult (*gm_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int grad_tensor_map(const uint8_t* map_dev, int32_t H, int32_t W, CUtensorMap* out) {
    static gm_tmap_encode_fn encode = []() -> gm_tmap_encode_fn {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<gm_tmap_encode_fn>(fn);
    }();
    if (!encode) return GM_ENODEV;
    const cuuint64_t dims[2] = {(cuuint64_t)W * 3ULL, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 3ULL};                      // bytes between rows (dim 1); a multiple of 16
    const cuuint32_t box[2] = {(cuuint32_t)gradfast::BGR_ROW_BYTES, (cuuint32_t)gradfast::PH};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(map_dev), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GM_OK : GM_EINVAL;
}

// Runs tiles [tile_begin, tile_begin + tile_count) of an n_tiles plan (the workspace is carved for the
// whole plan, every tile owns disjoint slices of it, so ranges may run on different streams).
static int dtedge_run(const uint8_t* map_dev, int32_t H, int32_t W,
                      const gm_tile* tiles_all_dev, int32_t n_tiles_all, int32_t max_tile,
                      int64_t total_px, int32_t tile_begin, int32_t tile_count, const gm_dtedge_params* params,
                      uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                      void* stream, cudaEvent_t* ev, cudaEvent_t grad_done = nullptr) {
    if (!map_dev || !tiles_all_dev || !out_dev || !params || !workspace_dev) return GM_EINVAL;
    if (H <= 0 || W <= 0 || n_tiles_all < 0 || total_px < 0) return GM_EINVAL;
    if (tile_begin < 0 || tile_count < 0 || tile_begin + tile_count > n_tiles_all) return GM_EINVAL;
    const gm_tile* tiles_dev = tiles_all_dev + tile_begin;
    const int32_t n_tiles = tile_count;
    if (max_tile <= 0 || max_tile > GM_MAX_TILE) return GM_ERANGE;
    if (params->layout != 0 && params->layout != 1) return GM_EINVAL;
    if (params->morph_open < 0 || params->morph_open > EO_MAX_ITER) return GM_ERANGE;
    if (n_tiles == 0) return GM_OK;
    if (workspace_bytes < gm_dtedge_workspace_bytes(total_px, n_tiles_all)) return GM_ENOSPC;
    GmTaps taps;
    int st = make_taps(params, &taps);
    if (st != GM_OK) return st;
    cudaStream_t s = gm_stream(stream);
    {
        static const cudaError_t attr_status = []() {
            cudaError_t e = cudaFuncSetAttribute(k_select_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelShared));
            if (e != cudaSuccess) return e;
            return cudaFuncSetAttribute(k_select_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelShared));
        }();
        if (attr_status != cudaSuccess) return (int)attr_status;
    }
    DtWorkspace w = carve(workspace_dev, total_px, n_tiles_all, GM_MAX_TILE);
    w.params += tile_begin;
    // GM_SELECT_SAMPLED=0: always take the two-pass order-statistic selection (same results; parity tests run both)
    const int sel_sample = gm_env_int("GM_SELECT_SAMPLED", 1);
    int sel_threads = gm_env_int("GM_SELECT_THREADS", GM_SELECT_DEFAULT_THREADS) & ~31;
    sel_threads = sel_threads < 128 ? 128 : (sel_threads > SEL_THREADS ? SEL_THREADS : sel_threads);
    int stage = 0;
#define GM_STAGE_MARK() do { if (ev) cudaEventRecord(ev[stage++], s); } while (0)
    GM_STAGE_MARK();

    const int nb = (max_tile + GB - 1) / GB;
    gradfast::Coef coef;
    if (!(params->flags & GM_DTEDGE_GENERIC_GRAD) && fast_grad_coef(taps, &coef)) {
        const int nbx = (max_tile + gradfast::BW - 1) / gradfast::BW, nby = (max_tile + gradfast::BH - 1) / gradfast::BH;
        {
            // shared-memory carveout in percent (-1: leave the driver's choice)
            static const int carve = gm_env_int("GM_GRAD_CARVEOUT", GM_GRAD_DEFAULT_CARVEOUT);
            static const cudaError_t carve_status = carve < 0 ? cudaSuccess :
                cudaFuncSetAttribute(gradfast::k_grad_fast<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            if (carve_status != cudaSuccess) return (int)carve_status;
        }
        // TMA path: one cp.async.bulk.tensor.2d per block fetches the BGR patch; needs a 16-byte aligned base and row pitch
        // (3 W % 16 == 0: every map whose width is a multiple of 16).  GM_GRAD_TMA=0 keeps the per-thread loads.
        static const int want_tma = gm_env_int("GM_GRAD_TMA", 1);
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        bool use_tma = false;
        if (want_tma && ((3LL * W) % 16 == 0) && ((reinterpret_cast<unsigned long long>(map_dev) & 15ULL) == 0ULL))
            use_tma = grad_tensor_map(map_dev, H, W, &tmap) == GM_OK;
        if (use_tma) {
            static const cudaError_t carve_tma = gm_env_int("GM_GRAD_CARVEOUT", GM_GRAD_DEFAULT_CARVEOUT) < 0 ? cudaSuccess :
                cudaFuncSetAttribute(gradfast::k_grad_fast<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     gm_env_int("GM_GRAD_CARVEOUT", GM_GRAD_DEFAULT_CARVEOUT));
            if (carve_tma != cudaSuccess) return (int)carve_tma;
        }
        // tile index on grid.z (<= 65,535 per launch): the blocks of a tile are scheduled together (L2 reuse of the halos)
        for (int t0 = 0; t0 < n_tiles; t0 += 65535) {
            const dim3 grid((unsigned)nbx, (unsigned)nby, (unsigned)std::min(65535, n_tiles - t0));
            if (use_tma) gradfast::k_grad_fast<true><<<grid, gradfast::THREADS, 0, s>>>(map_dev, W, 3LL * H * W, tiles_dev + t0, coef, w.S, tmap);
            else gradfast::k_grad_fast<false><<<grid, gradfast::THREADS, 0, s>>>(map_dev, W, 3LL * H * W, tiles_dev + t0, coef, w.S, tmap);
        }
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
    } else {
        const int R = taps.max_radius, HALO = R + 1, P = GB + 2 * HALO;
        const size_t smem = ((size_t)(P * P + 15) & ~(size_t)15) + (size_t)(GB + 2 + 2 * R) * (GB + 2) * 2 +
                            (size_t)(GB + 2) * (GB + 2);
        dim3 grid((unsigned)n_tiles, (unsigned)(nb * nb));
        k_grad<<<grid, GRAD_THREADS, smem, s>>>(map_dev, W, tiles_dev, taps, w.S); gm_note_launches(1);
        GM_LAUNCH_CHECK();
    }
    GM_STAGE_MARK();
    if (grad_done) GM_CUDA_TRY(cudaEventRecord(grad_done, s));
    // small-tile plans (side <= 128): selection fused with its consumer, the tile's keys in shared memory (GM_SMALL_FUSED=0
    // keeps the large-tile kernels; the parity tests run both)
    const int want_small = gm_env_int("GM_SMALL_FUSED", 1);
    const bool small = want_small && max_tile <= SMALL_SIDE && params->morph_open <= 1;
    if (small) {
        static const cudaError_t small_attr = []() {
            cudaError_t e = cudaFuncSetAttribute(k_select_edge_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
            if (e != cudaSuccess) return e;
            return cudaFuncSetAttribute(k_select_tail_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
        }();
        if (small_attr != cudaSuccess) return (int)small_attr;
    }
    if (small && !(params->flags & GM_DTEDGE_OTSU)) {
        k_select_edge_small<<<n_tiles, SMALL_THREADS, sizeof(SmallShared), s>>>(w.S, tiles_dev, params->p_hi / 100.0, w.params,
                                                                                params->morph_open, GM_MAX_TILE, tile_begin, w.zbits);
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
        GM_STAGE_MARK();
        GM_STAGE_MARK();
    } else {
        if (params->flags & GM_DTEDGE_OTSU)
            k_otsu_grad<<<n_tiles, sel_threads, 0, s>>>(w.S, tiles_dev, w.params);
        else
            k_select_grad<<<n_tiles, sel_threads, sizeof(SelShared), s>>>(w.S, tiles_dev, params->p_hi / 100.0, w.params, sel_sample);
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
        GM_STAGE_MARK();
        dim3 grid((unsigned)n_tiles, (unsigned)((max_tile + EO_ROWS - 1) / EO_ROWS));
        if (params->morph_open >= 2)
            k_edge_open_iter<<<grid, EO_THREADS, 0, s>>>(w.S, tiles_dev, w.params, params->morph_open, GM_MAX_TILE, tile_begin, w.zbits);
        else
            k_edge_open<<<grid, EO_THREADS, 0, s>>>(w.S, tiles_dev, w.params, params->morph_open, GM_MAX_TILE, tile_begin, w.zbits);
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
        GM_STAGE_MARK();
    }
    {
        // warps per tile x columns per lane; measured on B200 (8192^2, 676 tiles of 416^2): <4,4> 0.444 ms,
        // <2,8> 0.426 ms (0.383 with rows 16 ahead prefetched into L2), <1,16> 0.549 ms - the row recurrence is a dependent chain, so fewer, fatter lanes
        // only pay while the ALU pipe is the limit.  GM_CHAMFER_VARIANT selects the others for tuning.
        const int variant = gm_env_int("GM_CHAMFER_VARIANT", 0);
        if (max_tile <= 128) k_chamfer<1, 4><<<(unsigned)n_tiles, 32, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
        else if (max_tile <= 256) {
            if (variant == 1) k_chamfer<2, 4><<<(unsigned)n_tiles, 64, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
            else k_chamfer<1, 8><<<(unsigned)n_tiles, 32, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
        } else if (max_tile <= 512) {
            if (variant == 1) k_chamfer<4, 4><<<(unsigned)n_tiles, 128, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
            else if (variant == 2) k_chamfer<1, 16><<<(unsigned)n_tiles, 32, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
            else if (variant == 15) k_chamfer<2, 8><<<(unsigned)n_tiles, 64, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
            else k_chamfer<2, 8, 4, 16><<<(unsigned)n_tiles, 64, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
        } else {
            if (variant == 1) k_chamfer<8, 4><<<(unsigned)n_tiles, 256, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
            else k_chamfer<4, 8><<<(unsigned)n_tiles, 128, 0, s>>>(tiles_dev, GM_MAX_TILE, tile_begin, w.zbits, w.T);
        }
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
    }
    GM_STAGE_MARK();
    if (small) {
        k_select_tail_small<<<n_tiles, SMALL_THREADS, sizeof(SmallShared), s>>>(map_dev, W, 3LL * H * W, tiles_dev, w.params, w.S, w.T,
                                                                                params->layout, out_dev);
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
        GM_STAGE_MARK();
        GM_STAGE_MARK();
    } else {
        k_select_dist<<<n_tiles, sel_threads, sizeof(SelShared), s>>>(w.T, tiles_dev, w.params, sel_sample); gm_note_launches(1);
        GM_LAUNCH_CHECK();
        GM_STAGE_MARK();
        dim3 grid((unsigned)n_tiles, (unsigned)((max_tile + TAIL_ROWS - 1) / TAIL_ROWS));
        k_tail<<<grid, TAIL_THREADS, 0, s>>>(map_dev, W, 3LL * H * W, tiles_dev, w.params, w.S, w.T, params->layout, out_dev); gm_note_launches(1);
        GM_LAUNCH_CHECK();
        GM_STAGE_MARK();
    }
#undef GM_STAGE_MARK
    return GM_OK;
}

// Fork/join helper: tile ranges of one plan own disjoint slices of the workspace and of the output, and
// three of the six stages (chamfer, the two selects) are latency bound with a handful of warps per SM
// while the other three are issue bound.  Running a few tile ranges on side streams lets the hardware
// scheduler fill the idle issue slots of one range's latency-bound stage with another range's
// issue-bound stage.  The side streams and events are created once per device; the caller's stream
// forks into them and joins again, so the call stays stream-ordered for the caller.
namespace {

constexpr int GM_FORK_MAX = 8;

constexpr int GM_FORK_MAX_CHUNKS = 32;

struct ForkPool {
    cudaStream_t side[GM_FORK_MAX];
    cudaEvent_t fork, join[GM_FORK_MAX];
    cudaEvent_t grad_done[GM_FORK_MAX_CHUNKS];
    bool ready = false;
};

ForkPool g_fork[64];
std::mutex g_fork_mutex;

int fork_pool(ForkPool** out) {
    int dev = 0;
    GM_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GM_ERANGE;
    ForkPool& p = g_fork[dev];
    if (!p.ready) {
        for (int i = 0; i < GM_FORK_MAX; ++i) {
            GM_CUDA_TRY(cudaStreamCreateWithFlags(&p.side[i], cudaStreamNonBlocking));
            GM_CUDA_TRY(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming));
        }
        GM_CUDA_TRY(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
        for (int i = 0; i < GM_FORK_MAX_CHUNKS; ++i) GM_CUDA_TRY(cudaEventCreateWithFlags(&p.grad_done[i], cudaEventDisableTiming));
        p.ready = true;
    }
    *out = &p;
    return GM_OK;
}

}  // namespace

extern "C" int gm_dtedge_build_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                                  const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                                  int64_t total_px, const gm_dtedge_params* params,
                                  uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                                  void* stream) {
    // measured on B200 (8192^2 map, 676 tiles of 416^2): a build alone gains 5 % from 2-4 ranges on as many streams
    // (2.08 -> 1.96 ms), but beside the detection path of a step it loses 8 % (2.30 -> 2.50 ms per step), so the
    // default is one range on the caller's stream; GM_DTEDGE_CHUNKS / GM_DTEDGE_STREAMS opt in
    int chunks = gm_env_int("GM_DTEDGE_CHUNKS", GM_DTEDGE_DEFAULT_CHUNKS);
    int lanes = gm_env_int("GM_DTEDGE_STREAMS", GM_DTEDGE_DEFAULT_STREAMS);
    lanes = lanes < 1 ? 1 : (lanes > GM_FORK_MAX ? GM_FORK_MAX : lanes);
    if (chunks > n_tiles / 32) chunks = n_tiles / 32;          // a range below ~32 tiles cannot fill the GPU on its own
    if (chunks <= 1 || lanes <= 1)
        return dtedge_run(map_dev, H, W, tiles_dev, n_tiles, max_tile, total_px, 0, n_tiles, params, out_dev, workspace_dev,
                          workspace_bytes, stream, nullptr);
    if (lanes > chunks) lanes = chunks;
    std::lock_guard<std::mutex> lock(g_fork_mutex);            // the pool's events are reused call after call
    ForkPool* pool = nullptr;
    int st = fork_pool(&pool);
    if (st != GM_OK) return st;
    cudaStream_t s = gm_stream(stream);
    GM_CUDA_TRY(cudaEventRecord(pool->fork, s));
    for (int i = 0; i < lanes; ++i) GM_CUDA_TRY(cudaStreamWaitEvent(pool->side[i], pool->fork, 0));
    // skew: range k starts once range k-1 has finished its gradient kernel, so the ranges stay one stage apart and
    // the issue-bound gradient of one range runs beside the latency-bound chamfer / selection kernels of the
    // previous one (all ranges in phase would only compete for the same resource at the same time)
    const int skew = gm_env_int("GM_DTEDGE_SKEW", GM_DTEDGE_DEFAULT_SKEW);
    if (chunks > GM_FORK_MAX_CHUNKS) chunks = GM_FORK_MAX_CHUNKS;
    for (int k = 0; k < chunks && st == GM_OK; ++k) {
        const int32_t t0 = (int32_t)((int64_t)n_tiles * k / chunks), t1 = (int32_t)((int64_t)n_tiles * (k + 1) / chunks);
        if (skew && k > 0) GM_CUDA_TRY(cudaStreamWaitEvent(pool->side[k % lanes], pool->grad_done[k - 1], 0));
        st = dtedge_run(map_dev, H, W, tiles_dev, n_tiles, max_tile, total_px, t0, t1 - t0, params, out_dev, workspace_dev,
                        workspace_bytes, pool->side[k % lanes], nullptr, skew ? pool->grad_done[k] : nullptr);
    }
    for (int i = 0; i < lanes; ++i) {                            // always join, also after an error
        cudaEventRecord(pool->join[i], pool->side[i]);
        cudaStreamWaitEvent(s, pool->join[i], 0);
    }
    return st;
}

extern "C" int gm_dtedge_build_range_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                                        const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                                        int64_t total_px, int32_t tile_begin, int32_t tile_count,
                                        const gm_dtedge_params* params,
                                        uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                                        void* stream) {
    return dtedge_run(map_dev, H, W, tiles_dev, n_tiles, max_tile, total_px, tile_begin, tile_count, params, out_dev,
                      workspace_dev, workspace_bytes, stream, nullptr);
}

extern "C" int gm_dtedge_build_timed(const uint8_t* map_dev, int32_t H, int32_t W,
                                     const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                                     int64_t total_px, const gm_dtedge_params* params,
                                     uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                                     void* stream, float* stage_ms_host) {
    if (!stage_ms_host) return GM_EINVAL;
    cudaEvent_t ev[GM_DTEDGE_STAGES + 1];
    for (int i = 0; i <= GM_DTEDGE_STAGES; ++i) GM_CUDA_TRY(cudaEventCreate(&ev[i]));
    int st = dtedge_run(map_dev, H, W, tiles_dev, n_tiles, max_tile, total_px, 0, n_tiles, params, out_dev, workspace_dev,
                        workspace_bytes, stream, ev);
    if (st == GM_OK && n_tiles > 0) {
        cudaError_t e = cudaEventSynchronize(ev[GM_DTEDGE_STAGES]);
        if (e != cudaSuccess) st = (int)e;
        for (int i = 0; i < GM_DTEDGE_STAGES && st == GM_OK; ++i) cudaEventElapsedTime(&stage_ms_host[i], ev[i], ev[i + 1]);
    } else {
        for (int i = 0; i < GM_DTEDGE_STAGES; ++i) stage_ms_host[i] = 0.f;
    }
    for (int i = 0; i <= GM_DTEDGE_STAGES; ++i) cudaEventDestroy(ev[i]);
    return st;
}
