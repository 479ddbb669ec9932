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
//   k_edge_open    S >= S_thr -> 3x3 cross open -> bit-packed zero mask
//   k_chamfer      per tile, one warp: 3x3 chamfer DT (16.16 fixed point), forward + backward
//                  raster pass as per-row min-plus scans
//   k_select_dist  per tile: p1 / p99 of the chamfer field
//   k_tail         normalise, exp(-d/3) in float64, blend with the normalised gradient, pack
//                  [R,G,B,DT] (HWC or CHW)
#include <cfloat>
#include <cmath>
#include "gm_common.cuh"

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

// Per-tile constants produced by the two select kernels.
struct TileParams {
    unsigned int s_thr;      // edge <=> S >= s_thr
    float nrm_scale;         // cv2.normalize(acc, 0, 1, MINMAX): fmaf(acc, scale, shift)
    float nrm_shift;
    unsigned int pad;
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
// Exact order statistics of a tile's uint32 keys: one CTA per tile, MSB-first radix select.
// Pass 0 bins by (position of the top set bit, next three bits) so skewed distributions
// (gradient energies, distances) spread over ~100 bins; later passes take 11 bits each.

constexpr int SEL_THREADS = 1024;
constexpr int SEL_BINS = 2048;
constexpr int SEL_MAXR = 4;

__device__ __forceinline__ int log_bin(unsigned int key) {
    if (key == 0u) return 0;
    const int e = 31 - __clz(key);
    const unsigned int top = (e >= 3) ? (key >> (e - 3)) : (key << (3 - e));
    return ((e + 1) << 3) | (int)(top & 7u);
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
    unsigned int hist[SEL_MAXR][SEL_BINS];
    unsigned int lo[SEL_MAXR];      // low end of the key interval still holding the rank
    int rem[SEL_MAXR];              // undecided low bits
    unsigned int base[SEL_MAXR];    // number of keys below the interval
    int hid[SEL_MAXR];              // histogram used by this rank (ranks in one interval share)
    int bin[SEL_MAXR];
    unsigned int below[SEL_MAXR];
    unsigned int red_min[32], red_max[32];
};

// ranks[] ascending, nr <= SEL_MAXR.  On return vals[r] = the rank-th smallest key (0-based).
__device__ void block_select(const unsigned int* __restrict__ keys, int n, const unsigned int* ranks,
                             int nr, unsigned int* vals, unsigned int* kmin, unsigned int* kmax,
                             SelShared& sh) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    for (int i = tid; i < SEL_BINS; i += SEL_THREADS) sh.hist[0][i] = 0u;
    __syncthreads();
    unsigned int mn = 0xffffffffu, mx = 0u;
    for (int i = tid; i < n; i += SEL_THREADS) {
        const unsigned int k = keys[i];
        mn = min(mn, k); mx = max(mx, k);
        atomicAdd(&sh.hist[0][log_bin(k)], 1u);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    }
    if ((tid & 31) == 0) { sh.red_min[warp] = mn; sh.red_max[warp] = mx; }
    __syncthreads();
    if (warp < nr) warp_find_bin(sh.hist[0], 33 * 8, ranks[warp] , &sh.bin[warp], &sh.below[warp]);
    if (tid == 0) {
        unsigned int a = 0xffffffffu, b = 0u;
        for (int w = 0; w < SEL_THREADS / 32; ++w) { a = min(a, sh.red_min[w]); b = max(b, sh.red_max[w]); }
        *kmin = a; *kmax = b;
    }
    __syncthreads();
    if (tid < nr) {
        const int bin = sh.bin[tid];
        const int e = (bin >> 3) - 1;
        const unsigned int m = (unsigned int)(bin & 7);
        if (bin == 0) { sh.lo[tid] = 0u; sh.rem[tid] = 0; }
        else if (e >= 3) { sh.lo[tid] = (8u | m) << (e - 3); sh.rem[tid] = e - 3; }
        else { sh.lo[tid] = (8u | m) >> (3 - e); sh.rem[tid] = 0; }
        sh.base[tid] = sh.below[tid];
    }
    __syncthreads();
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
        for (int i = tid; i < SEL_MAXR * SEL_BINS; i += SEL_THREADS) (&sh.hist[0][0])[i] = 0u;
        __syncthreads();
        unsigned int lo_r[SEL_MAXR];
        int rem_r[SEL_MAXR], sft_r[SEL_MAXR], hid_r[SEL_MAXR];
#pragma unroll
        for (int r = 0; r < SEL_MAXR; ++r) {
            hid_r[r] = -1;
            if (r < nr) {
                lo_r[r] = sh.lo[r]; rem_r[r] = sh.rem[r]; hid_r[r] = sh.hid[r];
                const int bits = min(11, rem_r[r]);
                sft_r[r] = rem_r[r] - bits;
                // only the first rank of a shared histogram counts into it
                for (int q = 0; q < r; ++q) if (hid_r[q] == hid_r[r]) hid_r[r] = -2 - hid_r[r];
            }
        }
        for (int i = tid; i < n; i += SEL_THREADS) {
            const unsigned int k = keys[i];
#pragma unroll
            for (int r = 0; r < SEL_MAXR; ++r) {
                if (hid_r[r] >= 0) {
                    const unsigned int d = k - lo_r[r];
                    if (k >= lo_r[r] && (d >> rem_r[r]) == 0u) atomicAdd(&sh.hist[hid_r[r]][d >> sft_r[r]], 1u);
                }
            }
        }
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

__global__ void __launch_bounds__(SEL_THREADS)
k_select_grad(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles, double q_hi,
              TileParams* __restrict__ params) {
    __shared__ SelShared sh;
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
    block_select(S + t.px_off, n, ranks, 2, vals, &kmin, &kmax, sh);
    if (threadIdx.x == 0) {
        const double hi = np_lerp(acc_of(vals[0]), acc_of(vals[1]), g_sh);
        // smallest S with float64(acc(S)) >= hi  (acc is monotone in S)
        unsigned int up = 0xffffffffu;              // 0xffffffff = no pixel reaches the threshold
        if ((double)acc_of(0u) >= hi) up = 0u;
        else {
            unsigned int a = 0u, b = vals[1];       // acc(b) >= hi always holds for the upper order statistic
            if (!((double)acc_of(b) >= hi)) { a = b; b = 0xffffffffu; }
            if (b != 0xffffffffu) {
                while (b - a > 1u) {                // acc(a) < hi <= acc(b)
                    const unsigned int m = a + ((b - a) >> 1);
                    if ((double)acc_of(m) >= hi) b = m; else a = m;
                }
            }
            up = b;
        }
        params[blockIdx.x].s_thr = up;
        const double dmin = (double)acc_of(kmin), dmax = (double)acc_of(kmax);
        const double rng = __dsub_rn(dmax, dmin);
        const double scale = (rng > DBL_EPSILON) ? __ddiv_rn(1.0, rng) : 0.0;
        const double shift = __dsub_rn(0.0, __dmul_rn(dmin, scale));
        params[blockIdx.x].nrm_scale = (float)scale;
        params[blockIdx.x].nrm_shift = (float)shift;
    }
}

__global__ void __launch_bounds__(SEL_THREADS)
k_select_dist(const unsigned int* __restrict__ T, const gm_tile* __restrict__ tiles,
              TileParams* __restrict__ params) {
    __shared__ SelShared sh;
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
    block_select(T + t.px_off, n, ranks, 4, vals, &kmin, &kmax, sh);
    if (threadIdx.x == 0) {
        const double p1 = np_lerp(dist_of(vals[0]), dist_of(vals[1]), g_sh[0]);
        const double p99 = np_lerp(dist_of(vals[2]), dist_of(vals[3]), g_sh[1]);
        params[blockIdx.x].dist_lo = p1;
        const double span = __dsub_rn(p99, p1);
        params[blockIdx.x].dist_den = span > 1e-6 ? span : 1e-6;
    }
}

// ------------------------------------------------------------------------------------------
// k_edge_open: threshold + 3x3 cross open, 32x32 pixels per CTA, bit-packed output.

constexpr int EO_THREADS = 256;

__global__ void __launch_bounds__(EO_THREADS)
k_edge_open(const unsigned int* __restrict__ S, const gm_tile* __restrict__ tiles,
            const TileParams* __restrict__ params, int morph_open, int max_tile,
            unsigned int* __restrict__ zbits) {
    __shared__ unsigned char edge[36][40];
    __shared__ unsigned char ero[34][36];
    const gm_tile t = tiles[blockIdx.x];
    const int nbx = (t.w + 31) / 32;
    const int nby = (t.h + 31) / 32;
    if ((int)blockIdx.y >= nbx * nby) return;
    const int bx = ((int)blockIdx.y % nbx) * 32;
    const int by = ((int)blockIdx.y / nbx) * 32;
    const unsigned int thr = params[blockIdx.x].s_thr;
    const unsigned int* St = S + t.px_off;
    const int tid = threadIdx.x;
    // edge bits on the block + 2 ring; code 2 = outside the tile
    for (int i = tid; i < 36 * 36; i += EO_THREADS) {
        const int py = i / 36, px = i - py * 36;
        const int y = by - 2 + py, x = bx - 2 + px;
        unsigned char v = 2;
        if (y >= 0 && y < t.h && x >= 0 && x < t.w) v = (St[(long long)y * t.w + x] >= thr) ? 1 : 0;
        edge[py][px] = v;
    }
    __syncthreads();
    if (morph_open > 0) {
        // erosion: out-of-tile neighbours do not constrain; result only meaningful inside the tile
        for (int i = tid; i < 34 * 34; i += EO_THREADS) {
            const int py = i / 34, px = i - py * 34;          // block + 1 ring
            const int ey = py + 1, ex = px + 1;
            unsigned char c = edge[ey][ex];
            unsigned char v = 0;
            if (c != 2) {
                v = (c == 1) && (edge[ey - 1][ex] != 0) && (edge[ey + 1][ex] != 0) &&
                    (edge[ey][ex - 1] != 0) && (edge[ey][ex + 1] != 0);
            }
            ero[py][px] = v;        // outside the tile -> 0 (dilation ignores it)
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int oy = (tid >> 5) + 8 * k;
        const int ox = tid & 31;
        bool z;
        if (morph_open > 0) {
            const int py = oy + 1, px = ox + 1;
            z = ero[py][px] | ero[py - 1][px] | ero[py + 1][px] | ero[py][px - 1] | ero[py][px + 1];
        } else {
            z = edge[oy + 2][ox + 2] == 1;
        }
        const int y = by + oy, x = bx + ox;
        const bool inside = (y < t.h) && (x < t.w);
        const unsigned int word = __ballot_sync(0xffffffffu, z && inside);
        if (ox == 0 && y < t.h) {
            const int wpr = (t.w + 31) >> 5;
            zbits[zbits_offset(t.px_off, blockIdx.x, max_tile) + (long long)y * wpr + (bx >> 5)] = word;
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_chamfer: cv2.distanceTransform(DIST_L2, 3) = two raster passes of a 3x3 chamfer mask
// in 16.16 fixed point.  One warp owns a tile; lane l holds E consecutive columns of the
// current row in registers.  Each row is  c[j] = min(prev[j-1]+DG, prev[j]+HV, prev[j+1]+DG)
// (0 at zero pixels) followed by the min-plus scan f[j] = min_{k<=j} c[k] + (j-k)*HV, done as
// a prefix-min of c[k]-k*HV (local sequential scan + 5-step warp scan).  The backward pass
// mirrors it.  Out-of-image is INF; any final value >= INF is a tile without a zero pixel,
// for which cv2 saturates to DIST_MAX = UINT_MAX - DG.

constexpr int CH_HV = 62587;
constexpr int CH_DG = 89738;
constexpr int CH_INF = 1 << 29;
constexpr unsigned int CH_DIST_MAX = 0xffffffffu - (unsigned int)CH_DG;
constexpr int CH_WARPS = 4;

template <int E>
__device__ __forceinline__ unsigned int load_zero_bits(const unsigned int* __restrict__ zrow, int wpr, int lane) {
    // bits [lane*E, lane*E+E) of the row; E <= 32
    const int b0 = lane * E;
    const int w0 = b0 >> 5;
    const int sh = b0 & 31;
    unsigned int lo = (w0 < wpr) ? zrow[w0] : 0u;
    unsigned int hi = (w0 + 1 < wpr) ? zrow[w0 + 1] : 0u;
    const unsigned int v = __funnelshift_r(lo, hi, sh);
    return (E == 32) ? v : (v & ((1u << E) - 1u));
}

template <int E>
__global__ void __launch_bounds__(CH_WARPS * 32)
k_chamfer(const gm_tile* __restrict__ tiles, int n_tiles, int max_tile,
          const unsigned int* __restrict__ zbits, unsigned int* __restrict__ T) {
    const int ti = blockIdx.x * CH_WARPS + (threadIdx.x >> 5);
    if (ti >= n_tiles) return;
    const gm_tile t = tiles[ti];
    const int lane = gm_lane();
    const int w = t.w, h = t.h;
    const int wpr = (w + 31) >> 5;
    const unsigned int* zb = zbits + zbits_offset(t.px_off, ti, max_tile);
    unsigned int* Tt = T + t.px_off;
    const int col0 = lane * E;

    int prev[E];
#pragma unroll
    for (int e = 0; e < E; ++e) prev[e] = CH_INF;

    // ---- forward: top -> bottom, left -> right
    unsigned int zcur = load_zero_bits<E>(zb, wpr, lane);
    for (int y = 0; y < h; ++y) {
        unsigned int znext = 0u;
        if (y + 1 < h) znext = load_zero_bits<E>(zb + (long long)(y + 1) * wpr, wpr, lane);
        int left = __shfl_up_sync(0xffffffffu, prev[E - 1], 1);
        int right = __shfl_down_sync(0xffffffffu, prev[0], 1);
        if (lane == 0) left = CH_INF;
        if (lane == 31) right = CH_INF;
        int q[E];
        int run = CH_INF * 2;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = (e == 0) ? left : prev[e - 1];
            const int r = (e == E - 1) ? right : prev[e + 1];
            int c = min(prev[e] + CH_HV, min(l, r) + CH_DG);
            if ((zcur >> e) & 1u) c = 0;
            if (col0 + e >= w) c = CH_INF;
            run = min(run, c - e * CH_HV);
            q[e] = run;
        }
        // warp exclusive prefix-min of (lane total - col0*HV)
        const int tot = run - col0 * CH_HV;
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl = min(incl, v);
        }
        int excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = CH_INF * 2;
        const int carry = excl + col0 * CH_HV;       // best c[k]-k*HV of lower lanes, rebased to this lane
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int f = min(q[e], carry) + e * CH_HV;
            if (col0 + e >= w) f = CH_INF;
            prev[e] = f;
            if (col0 + e < w) Tt[(long long)y * w + col0 + e] = (unsigned int)f;
        }
        zcur = znext;
    }

    // ---- backward: bottom -> top, right -> left
#pragma unroll
    for (int e = 0; e < E; ++e) prev[e] = CH_INF;
    int cur[E];
#pragma unroll
    for (int e = 0; e < E; ++e) cur[e] = (h > 0 && col0 + e < w) ? (int)Tt[(long long)(h - 1) * w + col0 + e] : CH_INF;
    for (int y = h - 1; y >= 0; --y) {
        int nxt[E];
#pragma unroll
        for (int e = 0; e < E; ++e) nxt[e] = (y > 0 && col0 + e < w) ? (int)Tt[(long long)(y - 1) * w + col0 + e] : CH_INF;
        int left = __shfl_up_sync(0xffffffffu, prev[E - 1], 1);
        int right = __shfl_down_sync(0xffffffffu, prev[0], 1);
        if (lane == 0) left = CH_INF;
        if (lane == 31) right = CH_INF;
        int q[E];
        int run = CH_INF * 2;
#pragma unroll
        for (int e = E - 1; e >= 0; --e) {
            const int l = (e == 0) ? left : prev[e - 1];
            const int r = (e == E - 1) ? right : prev[e + 1];
            int c = min(min(cur[e], prev[e] + CH_HV), min(l, r) + CH_DG);
            if (col0 + e >= w) c = CH_INF;
            run = min(run, c + e * CH_HV);
            q[e] = run;
        }
        const int tot = run + col0 * CH_HV;
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, incl, d);
            if (lane + d < 32) incl = min(incl, v);
        }
        int excl = __shfl_down_sync(0xffffffffu, incl, 1);
        if (lane == 31) excl = CH_INF * 2;
        const int carry = excl - col0 * CH_HV;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int f = min(q[e], carry) - e * CH_HV;
            if (col0 + e >= w) f = CH_INF;
            prev[e] = f;
            if (col0 + e < w)
                Tt[(long long)y * w + col0 + e] = (f >= CH_INF) ? CH_DIST_MAX : (unsigned int)f;
            cur[e] = nxt[e];
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_tail: per pixel, float64 exactly where numpy computes in float64.

constexpr int TAIL_THREADS = 256;

__global__ void __launch_bounds__(TAIL_THREADS)
k_tail(const uint8_t* __restrict__ map, int W, const gm_tile* __restrict__ tiles,
       const TileParams* __restrict__ params, const unsigned int* __restrict__ S,
       const unsigned int* __restrict__ T, int layout, uint8_t* __restrict__ out) {
    const gm_tile t = tiles[blockIdx.x];
    const int n = t.h * t.w;
    const int i = blockIdx.y * TAIL_THREADS + threadIdx.x;
    if (i >= n) return;
    const TileParams p = params[blockIdx.x];
    const int y = i / t.w;
    const int x = i - y * t.w;
    const float acc = acc_of(S[t.px_off + i]);
    const float dist = dist_of(T[t.px_off + i]);
    double d = __ddiv_rn(__dsub_rn((double)dist, p.dist_lo), p.dist_den);
    d = fmin(fmax(d, 0.0), 1.0);
    const double soft = exp(__ddiv_rn(-d, 3.0));
    const float nrm = __fmaf_rn(acc, p.nrm_scale, p.nrm_shift);
    const float g03 = __fmul_rn(0.3f, nrm);
    double v = __dadd_rn(__dmul_rn(0.7, soft), (double)g03);
    v = fmin(fmax(v, 0.0), 1.0);
    const unsigned int dt = (unsigned int)__dmul_rn(v, 255.0);      // truncation like astype(uint8)
    const uint8_t* px = map + ((long long)(t.y0 + y) * W + (t.x0 + x)) * 3LL;
    const unsigned int b = __ldg(px), g = __ldg(px + 1), r = __ldg(px + 2);
    if (layout == 0) {
        reinterpret_cast<unsigned int*>(out)[t.px_off + i] = r | (g << 8) | (b << 16) | (dt << 24);
    } else {
        uint8_t* o = out + 4LL * t.px_off;
        o[i] = (uint8_t)r;
        o[(long long)n + i] = (uint8_t)g;
        o[2LL * n + i] = (uint8_t)b;
        o[3LL * n + i] = (uint8_t)dt;
    }
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

static int dtedge_run(const uint8_t* map_dev, int32_t H, int32_t W,
                      const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                      int64_t total_px, const gm_dtedge_params* params,
                      uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                      void* stream, cudaEvent_t* ev) {
    if (!map_dev || !tiles_dev || !out_dev || !params || !workspace_dev) return GM_EINVAL;
    if (H <= 0 || W <= 0 || n_tiles < 0 || total_px < 0) return GM_EINVAL;
    if (max_tile <= 0 || max_tile > GM_MAX_TILE) return GM_ERANGE;
    if (params->layout != 0 && params->layout != 1) return GM_EINVAL;
    if (params->morph_open < 0 || params->morph_open > 1) return GM_ERANGE;
    if (n_tiles == 0) return GM_OK;
    if (workspace_bytes < gm_dtedge_workspace_bytes(total_px, n_tiles)) return GM_ENOSPC;
    GmTaps taps;
    int st = make_taps(params, &taps);
    if (st != GM_OK) return st;
    cudaStream_t s = gm_stream(stream);
    DtWorkspace w = carve(workspace_dev, total_px, n_tiles, GM_MAX_TILE);
    int stage = 0;
#define GM_STAGE_MARK() do { if (ev) cudaEventRecord(ev[stage++], s); } while (0)
    GM_STAGE_MARK();

    const int nb = (max_tile + GB - 1) / GB;
    {
        const int R = taps.max_radius, HALO = R + 1, P = GB + 2 * HALO;
        const size_t smem = ((size_t)(P * P + 15) & ~(size_t)15) + (size_t)(GB + 2 + 2 * R) * (GB + 2) * 2 +
                            (size_t)(GB + 2) * (GB + 2);
        dim3 grid((unsigned)n_tiles, (unsigned)(nb * nb));
        k_grad<<<grid, GRAD_THREADS, smem, s>>>(map_dev, W, tiles_dev, taps, w.S); gm_note_launches(1);
        GM_LAUNCH_CHECK();
    }
    GM_STAGE_MARK();
    k_select_grad<<<n_tiles, SEL_THREADS, 0, s>>>(w.S, tiles_dev, params->p_hi / 100.0, w.params); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    GM_STAGE_MARK();
    {
        dim3 grid((unsigned)n_tiles, (unsigned)(nb * nb));
        k_edge_open<<<grid, EO_THREADS, 0, s>>>(w.S, tiles_dev, w.params, params->morph_open, GM_MAX_TILE, w.zbits); gm_note_launches(1);
        GM_LAUNCH_CHECK();
    }
    GM_STAGE_MARK();
    {
        const unsigned blocks = (unsigned)((n_tiles + CH_WARPS - 1) / CH_WARPS);
        if (max_tile <= 128) k_chamfer<4><<<blocks, CH_WARPS * 32, 0, s>>>(tiles_dev, n_tiles, GM_MAX_TILE, w.zbits, w.T);
        else if (max_tile <= 256) k_chamfer<8><<<blocks, CH_WARPS * 32, 0, s>>>(tiles_dev, n_tiles, GM_MAX_TILE, w.zbits, w.T);
        else if (max_tile <= 416) k_chamfer<13><<<blocks, CH_WARPS * 32, 0, s>>>(tiles_dev, n_tiles, GM_MAX_TILE, w.zbits, w.T);
        else if (max_tile <= 512) k_chamfer<16><<<blocks, CH_WARPS * 32, 0, s>>>(tiles_dev, n_tiles, GM_MAX_TILE, w.zbits, w.T);
        else k_chamfer<32><<<blocks, CH_WARPS * 32, 0, s>>>(tiles_dev, n_tiles, GM_MAX_TILE, w.zbits, w.T);
        gm_note_launches(1);
        GM_LAUNCH_CHECK();
    }
    GM_STAGE_MARK();
    k_select_dist<<<n_tiles, SEL_THREADS, 0, s>>>(w.T, tiles_dev, w.params); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    GM_STAGE_MARK();
    {
        const long long max_px = (long long)max_tile * max_tile;
        dim3 grid((unsigned)n_tiles, (unsigned)((max_px + TAIL_THREADS - 1) / TAIL_THREADS));
        k_tail<<<grid, TAIL_THREADS, 0, s>>>(map_dev, W, tiles_dev, w.params, w.S, w.T, params->layout, out_dev); gm_note_launches(1);
        GM_LAUNCH_CHECK();
    }
    GM_STAGE_MARK();
#undef GM_STAGE_MARK
    return GM_OK;
}

extern "C" int gm_dtedge_build_u8(const uint8_t* map_dev, int32_t H, int32_t W,
                                  const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                                  int64_t total_px, const gm_dtedge_params* params,
                                  uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                                  void* stream) {
    return dtedge_run(map_dev, H, W, tiles_dev, n_tiles, max_tile, total_px, params, out_dev, workspace_dev,
                      workspace_bytes, stream, nullptr);
}

extern "C" int gm_dtedge_build_timed(const uint8_t* map_dev, int32_t H, int32_t W,
                                     const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_tile,
                                     int64_t total_px, const gm_dtedge_params* params,
                                     uint8_t* out_dev, void* workspace_dev, size_t workspace_bytes,
                                     void* stream, float* stage_ms_host) {
    if (!stage_ms_host) return GM_EINVAL;
    cudaEvent_t ev[GM_DTEDGE_STAGES + 1];
    for (int i = 0; i <= GM_DTEDGE_STAGES; ++i) GM_CUDA_TRY(cudaEventCreate(&ev[i]));
    int st = dtedge_run(map_dev, H, W, tiles_dev, n_tiles, max_tile, total_px, params, out_dev, workspace_dev,
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
