// a6-a9, a11, a12: tile->map remap, border filter, strike angle, exact class-wise greedy rotated
// NMS (per tile and global) and the dual-scale late fusion.
//
// Reference: detect_symbols body (Detect_OBB.py:228-264), merge_detections (:176-200),
// cross_scale_consensus_filter (:347-423).  Both reference loops are sequential and O(n^2) in
// shapely calls; here
//   1. boxes are prepared once (pair-local fp32 geometry around an fp32 reference point) and binned on a
//      uniform grid whose cell is the largest box extent, keyed by (group, cell row, cell col)
//      and radix-sorted, so each box meets only the boxes of its 3x3 cell neighbourhood;
//   2. pairs that pass the AABB test get the rotated IoU (fp32, float64 re-check within 1e-4
//      of the threshold so the decision equals the reference's float64 comparison);
//   3. the sequential greedy semantics are recovered exactly by a priority fixpoint inside
//      one cooperative kernel: a box is decided once all its higher-priority neighbours are
//      (NMS), respectively once every earlier box within two hops is (fusion).
// The per-tile stage of the pipeline (at most max_det boxes per tile) has its own form: one CTA per tile, ranks by counting,
// one IoU bit per same-class pair, a sweep of the bit masks in score order (k_tile_nms, gm_tile_postprocess_bounded).
#include <cooperative_groups.h>
#include "gm_common.cuh"
#ifndef GM_COOP_DEFAULT_MAX_BLOCKS
#define GM_COOP_DEFAULT_MAX_BLOCKS 32       // 0 = as many CTAs as fit (one wave); step 2.25 ms with 0, 2.21 with 74 or 32, 2.33 with 8
#endif
#include "geom.cuh"

namespace cg = cooperative_groups;

namespace {

// ========================================================================================
// Exclusive scan of uint32 (3 launches): 4096 items per block.

constexpr int SCAN_T = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total, unsigned int* sh) {
    // sh: at least 33 words
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        unsigned int w = lane < nw ? sh[lane] : 0u;
        unsigned int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < nw) sh[lane] = wi - w;
        if (lane == nw - 1) sh[32] = wi;
    }
    __syncthreads();
    const unsigned int r = sh[warp] + incl - v;
    *total = sh[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T)
k_scan_local(const unsigned int* __restrict__ in, unsigned int* __restrict__ out, long long n,
             unsigned int* __restrict__ block_sums) {
    __shared__ unsigned int sh[33];
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    unsigned int v[SCAN_ITEMS];
    unsigned int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        sum += v[k];
    }
    unsigned int total;
    unsigned int run = block_exclusive_scan(sum, &total, sh);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T)
k_scan_tops(unsigned int* __restrict__ block_sums, int nb, unsigned int* __restrict__ total_out) {
    __shared__ unsigned int sh[33];
    unsigned int carry = 0;
    for (int base = 0; base < nb; base += SCAN_T) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < nb ? block_sums[i] : 0u;
        unsigned int total;
        const unsigned int e = block_exclusive_scan(v, &total, sh);
        if (i < nb) block_sums[i] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_T)
k_scan_add(unsigned int* __restrict__ out, long long n, const unsigned int* __restrict__ block_sums) {
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    const unsigned int add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

inline long long scan_blocks(long long n) { return (n + SCAN_BLOCK - 1) / SCAN_BLOCK; }

// tmp: scan_blocks(n) words.  total_out may be null.
int exclusive_scan_u32(const unsigned int* in, unsigned int* out, long long n, unsigned int* tmp,
                       unsigned int* total_out, cudaStream_t s) {
    if (n <= 0) {
        if (total_out) GM_CUDA_TRY(cudaMemsetAsync(total_out, 0, sizeof(unsigned int), s));
        return GM_OK;
    }
    const long long nb = scan_blocks(n);
    k_scan_local<<<(unsigned)nb, SCAN_T, 0, s>>>(in, out, n, tmp); gm_note_launches(1);
    k_scan_tops<<<1, SCAN_T, 0, s>>>(tmp, (int)nb, total_out); gm_note_launches(1);
    k_scan_add<<<(unsigned)nb, SCAN_T, 0, s>>>(out, n, tmp); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

// ========================================================================================
// Stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass.

constexpr int RS_T = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_BLOCK = RS_T * RS_ITEMS;
constexpr int RS_MAX_PASSES = 8;            // 64-bit keys

inline long long rs_blocks(long long n) { return (n + RS_BLOCK - 1) / RS_BLOCK; }

__global__ void __launch_bounds__(RS_T)
k_rs_hist(const unsigned long long* __restrict__ keys, long long n, int shift, unsigned int* __restrict__ hist,
          int nb, unsigned int* __restrict__ digit_total) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_BLOCK;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const long long i = base + k * RS_T + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    const unsigned int c = h[threadIdx.x];
    hist[(long long)threadIdx.x * nb + blockIdx.x] = c;
    if (c) atomicAdd(&digit_total[threadIdx.x], c);          // 256 counters of this pass (zeroed once per sort)
}

// Offsets of one pass in ONE launch (it replaces a three-kernel exclusive scan over 256 * nb entries): a warp per digit.
// The digit's base is the sum of the totals of the smaller digits (accumulated by k_rs_hist); the warp then turns its row
// of per-block counts into running offsets, 32 blocks per step.
__global__ void __launch_bounds__(256)
k_rs_offsets(unsigned int* __restrict__ hist, int nb, const unsigned int* __restrict__ digit_total) {
    const int lane = threadIdx.x & 31;
    const int d = (int)blockIdx.x * 8 + (threadIdx.x >> 5);              // 32 CTAs x 8 warps = 256 digits
    unsigned int base = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int q = k * 32 + lane;
        base += q < d ? digit_total[q] : 0u;
    }
    base = __reduce_add_sync(0xffffffffu, base);
    unsigned int* row = hist + (long long)d * nb;
    unsigned int carry = base;
    for (int b0 = 0; b0 < nb; b0 += 128) {
        unsigned int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                     // four independent loads in flight per lane
            const int b = b0 + k * 32 + lane;
            v[k] = b < nb ? row[b] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int b = b0 + k * 32 + lane;
            unsigned int incl = v[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (b < nb) row[b] = carry + incl - v[k];
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

__global__ void __launch_bounds__(RS_T)
k_rs_scatter(const unsigned long long* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
             unsigned long long* __restrict__ keys_out, unsigned int* __restrict__ vals_out,
             long long n, int shift, const unsigned int* __restrict__ hist_scanned, int nb) {
    __shared__ unsigned int cnt[RS_T / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (RS_T / 32) * 256; i += RS_T) (&cnt[0][0])[i] = 0u;
    __syncthreads();
    // warp w owns the contiguous segment [base + w*256, +256): rounds of 32 keep the order
    const long long seg = (long long)blockIdx.x * RS_BLOCK + (long long)warp * (32 * RS_ITEMS);
    unsigned long long key[RS_ITEMS];
    unsigned int val[RS_ITEMS], rank[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const long long i = seg + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[i] : 0ull;
        val[r] = ok ? vals_in[i] : 0u;
        dig[r] = ok ? ((unsigned)(key[r] >> shift) & 255u) : 256u;
        const unsigned int peers = __match_any_sync(0xffffffffu, dig[r]);
        const unsigned int before = __popc(peers & ((1u << lane) - 1u));
        unsigned int pre = 0u;
        if (ok) pre = cnt[warp][dig[r]];
        __syncwarp();
        if (ok && before == 0u) cnt[warp][dig[r]] = pre + __popc(peers);
        __syncwarp();
        rank[r] = pre + before;
    }
    __syncthreads();
    {
        const int d = threadIdx.x;      // RS_T == 256 digits
        unsigned int run = hist_scanned[(long long)d * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_T / 32; ++w) {
            const unsigned int c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        if (dig[r] < 256u) {
            const unsigned int pos = cnt[warp][dig[r]] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

struct SortBufs {
    unsigned long long *ka, *kb;
    unsigned int *va, *vb;
    unsigned int* hist;       // 256 * rs_blocks(n)
    unsigned int* scan_tmp;   // scan_blocks(256 * rs_blocks(n))
    unsigned int* dtot;       // RS_MAX_PASSES * 256 digit totals, one row per pass
};

// Sorts (ka, va) by the low `bits` bits of the key; returns 0 if the result is in (ka, va),
// 1 if in (kb, vb), negative on error.
int radix_sort_pairs(const SortBufs& b, long long n, int bits, cudaStream_t s) {
    if (n <= 0) return 0;
    const int passes = (bits + 7) / 8;
    if (passes > RS_MAX_PASSES) return -1001;
    const long long nb = rs_blocks(n);
    unsigned long long* kin = b.ka; unsigned long long* kout = b.kb;
    unsigned int* vin = b.va; unsigned int* vout = b.vb;
    // Offsets of a pass: one 256-warp kernel while a warp's row of per-block counts is short (the launch-bound regime:
    // a rank's share of a sharded step), the three-kernel scan over 256 * nb entries when rows are long (measured on the
    // c5 step at N = 1, nb = 1294: 6.75 ms against 6.15 ms for the detection path with the one-kernel form everywhere).
    static const long long fused_max_nb = gm_env_int("GM_RS_FUSED_MAX_BLOCKS", 640);
    const bool fused = nb <= fused_max_nb;
    if (cudaMemsetAsync(b.dtot, 0, (size_t)passes * 256 * sizeof(unsigned int), s) != cudaSuccess) return -1000;
    for (int p = 0; p < passes; ++p) {
        k_rs_hist<<<(unsigned)nb, RS_T, 0, s>>>(kin, n, p * 8, b.hist, (int)nb, b.dtot + p * 256); gm_note_launches(1);
        if (fused) {
            k_rs_offsets<<<32, 256, 0, s>>>(b.hist, (int)nb, b.dtot + p * 256); gm_note_launches(1);
        } else {
            int st = exclusive_scan_u32(b.hist, b.hist, 256 * nb, b.scan_tmp, nullptr, s);
            if (st != GM_OK) return st > 0 ? -st : st;
        }
        k_rs_scatter<<<(unsigned)nb, RS_T, 0, s>>>(kin, vin, kout, vout, n, p * 8, b.hist, (int)nb); gm_note_launches(1);
        unsigned long long* tk = kin; kin = kout; kout = tk;
        unsigned int* tv = vin; vin = vout; vout = tv;
    }
    if (cudaGetLastError() != cudaSuccess) return -1000;
    return passes & 1;
}

// ========================================================================================
// Box preparation, grid binning.

__device__ __forceinline__ unsigned int enc_f32(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned int e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

struct Extent {            // encoded with enc_f32
    unsigned int minx, miny, maxx, maxy, maxext;
    unsigned int edge_count;         // pairs found (may exceed capacity)
    unsigned int undecided[2];
    unsigned int n_out;
    unsigned int maxrad;             // largest Chebyshev distance of a corner from its box's corner-mean centre (enc_f32, rounded up)
    unsigned int pad[2];
};

__global__ void k_extent_init(Extent* e) {
    e->minx = e->miny = 0xffffffffu;
    e->maxx = e->maxy = 0u;
    e->maxext = enc_f32(0.f);
    e->maxrad = enc_f32(0.f);
    e->edge_count = 0u;
    e->undecided[0] = e->undecided[1] = 0u;
    e->n_out = 0u;
}

// Stable conf-descending sort key: ascending radix order == descending confidence.
__device__ __forceinline__ unsigned int conf_key_desc(float c) { return ~enc_f32(c); }

#ifndef GM_PREPARE_MINB
#define GM_PREPARE_MINB 1                 // minimum resident CTAs per SM asked of the register allocator (tuning)
#endif
__global__ void __launch_bounds__(256, GM_PREPARE_MINB)
k_prepare(const double* __restrict__ boxes, const float* __restrict__ conf, const int* __restrict__ major,
          long long n, QPoly* __restrict__ qp, QWin* __restrict__ qw, float4* __restrict__ aabb, Extent* __restrict__ ext,
          unsigned long long* __restrict__ sort_key, unsigned int* __restrict__ sort_val) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float mnx = 3e38f, mny = 3e38f, mxx = -3e38f, mxy = -3e38f, mext = 0.f, mrad = 0.f;
    QPoly p = QPoly{};
    QWin wn = QWin{};
    if (i < n) {
        const double* b = boxes + i * 8;
        qbox_from_corners(b, p, wn);
        qpoly_mark_concave(b, p);
        double x0 = fmin(fmin(b[0], b[2]), fmin(b[4], b[6])), x1 = fmax(fmax(b[0], b[2]), fmax(b[4], b[6]));
        double y0 = fmin(fmin(b[1], b[3]), fmin(b[5], b[7])), y1 = fmax(fmax(b[1], b[3]), fmax(b[5], b[7]));
        // outward-rounded fp32 AABB: never rejects a pair the float64 boxes would overlap
        const float fx0 = __double2float_rd(x0), fy0 = __double2float_rd(y0);
        const float fx1 = __double2float_ru(x1), fy1 = __double2float_ru(y1);
        aabb[i] = make_float4(fx0, fy0, fx1, fy1);
        const bool finite = isfinite(fx0) && isfinite(fy0) && isfinite(fx1) && isfinite(fy1);
        if (finite) {
            mnx = fx0; mny = fy0; mxx = fx1; mxy = fy1; mext = fmaxf(fx1 - fx0, fy1 - fy0);
            // reach of the box around the centre the border filter tests (mean of the corners, Detect_OBB.py:159-165)
            const double cx = 0.25 * ((b[0] + b[2]) + (b[4] + b[6])), cy = 0.25 * ((b[1] + b[3]) + (b[5] + b[7]));
            mrad = __double2float_ru(fmax(fmax(x1 - cx, cx - x0), fmax(y1 - cy, cy - y0)));
        }
        const unsigned long long hi = major ? (unsigned long long)(unsigned int)major[i] : 0ull;
        sort_key[i] = (hi << 32) | (unsigned long long)conf_key_desc(conf[i]);
        sort_val[i] = (unsigned int)i;
    }
    {
        // The 64- and 96-byte records of a warp are contiguous in memory (2 KB and 3 KB): staged through shared memory
        // (odd word pitch: conflict free) and written as coalesced 16-byte vectors.  Every thread storing its own record
        // made each store instruction touch 32 separate sectors, and the kernel sat on the LSU queue (ncu: lg_throttle).
        __shared__ unsigned int stage[256 / 32][32 * 25];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const long long w0 = i - lane;                           // first box of this warp
        if (w0 < n) {
            unsigned int* st = stage[warp];
            const int cnt = (int)((n - w0) < 32 ? (n - w0) : 32);
            const unsigned int* pw = reinterpret_cast<const unsigned int*>(&p);
#pragma unroll
            for (int k = 0; k < 16; ++k) st[lane * 17 + k] = pw[k];
            __syncwarp();
            uint4* dst = reinterpret_cast<uint4*>(qp + w0);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int g = r * 32 + lane, rec = g >> 2, part = (g & 3) * 4;
                if (rec < cnt) dst[g] = make_uint4(st[rec * 17 + part], st[rec * 17 + part + 1], st[rec * 17 + part + 2], st[rec * 17 + part + 3]);
            }
            __syncwarp();
            const unsigned int* ww = reinterpret_cast<const unsigned int*>(&wn);
#pragma unroll
            for (int k = 0; k < 24; ++k) st[lane * 25 + k] = ww[k];
            __syncwarp();
            uint4* dw = reinterpret_cast<uint4*>(qw + w0);
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const int g = r * 32 + lane, rec = g / 6, part = (g - rec * 6) * 4;
                if (rec < cnt) dw[g] = make_uint4(st[rec * 25 + part], st[rec * 25 + part + 1], st[rec * 25 + part + 2], st[rec * 25 + part + 3]);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, d));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, d));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, d));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, d));
        mext = fmaxf(mext, __shfl_xor_sync(0xffffffffu, mext, d));
        mrad = fmaxf(mrad, __shfl_xor_sync(0xffffffffu, mrad, d));
    }
    if ((threadIdx.x & 31) == 0 && mxx >= mnx) {
        atomicMin(&ext->minx, enc_f32(mnx));
        atomicMin(&ext->miny, enc_f32(mny));
        atomicMax(&ext->maxx, enc_f32(mxx));
        atomicMax(&ext->maxy, enc_f32(mxy));
        atomicMax(&ext->maxext, enc_f32(mext));
        atomicMax(&ext->maxrad, enc_f32(mrad));
    }
}

// rank[input index] and order[rank] from the sorted value array.
__global__ void __launch_bounds__(256)
k_ranks(const unsigned int* __restrict__ sorted_val, long long n, unsigned int* __restrict__ rank,
        int* __restrict__ order) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const unsigned int i = sorted_val[r];
    rank[i] = (unsigned int)r;
    if (order) order[r] = (int)i;
}

constexpr int CELL_BITS = 12;
constexpr int CELL_MAX = (1 << CELL_BITS) - 1;

__global__ void __launch_bounds__(256)
k_cell_keys(const float4* __restrict__ aabb, const int* __restrict__ group, const unsigned char* __restrict__ active,
            unsigned int inactive_group, long long n, const Extent* __restrict__ ext,
            unsigned long long* __restrict__ key, unsigned int* __restrict__ val) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float minx = dec_f32(ext->minx), miny = dec_f32(ext->miny);
    const float maxx = dec_f32(ext->maxx), maxy = dec_f32(ext->maxy);
    const float span = fmaxf(maxx - minx, maxy - miny);
    float cell = fmaxf(dec_f32(ext->maxext) * 1.001f + 1e-3f, span / 4000.f);
    const float4 a = aabb[i];
    const float cx = 0.5f * (a.x + a.z), cy = 0.5f * (a.y + a.w);
    int ix = (int)((cx - minx) / cell), iy = (int)((cy - miny) / cell);
    const bool ok = (cx == cx) && (cy == cy) && (!active || active[i]);
    ix = min(max(ix, 0), CELL_MAX);
    iy = min(max(iy, 0), CELL_MAX);
    // inactive / non-finite boxes get a group of their own (max_group + 1) that nobody queries
    const unsigned int gi = (unsigned int)group[i];
    const unsigned long long g = (ok && gi < inactive_group) ? (unsigned long long)gi : (unsigned long long)inactive_group;
    key[i] = (g << (2 * CELL_BITS)) | ((unsigned long long)iy << CELL_BITS) | (unsigned long long)ix;
    val[i] = (unsigned int)i;
}

__device__ __forceinline__ long long lower_bound_u64(const unsigned long long* __restrict__ a, long long n,
                                                     unsigned long long v) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool aabb_overlap(const float4& a, const float4& b) {
    return a.x <= b.z && b.x <= a.z && a.y <= b.w && b.y <= a.w;
}

struct Edge { int hi, lo; };       // NMS: hi suppresses lo.  Fusion: hi < lo in flat order.

// Threshold-adjacent pairs since the last reset (north star: "pairs whose IoU lies within 1e-5 of a threshold are
// reported separately"): [0] pairs whose fp32 IoU fell within 1e-4 of the threshold and were therefore decided from the
// float64 IoU, [1] those of them whose float64 IoU lies within 1e-5 of the threshold.  Cold path: two atomics per such pair.
__device__ unsigned long long g_adjacent_stats[2];

// kFusion == false: pair (j,i) is an edge iff same group (class [x tile]), rank[j] < rank[i], IoU >= thr.
// kFusion == true : same class, different scale, both active, IoU >= thr; stored once (hi < lo).
// The float64 paths of the pair loop are rare (a concave quad; a value within 1e-4 of the threshold: ~1e-4 of the pairs)
// but their registers would set the kernel's allocation (206): behind calls they cost the hot loop nothing.
__device__ __noinline__ double discover_f64_general(const double* a, const double* b) { return iou_f64_general(a, b); }
__device__ __noinline__ double discover_f64_convex(const double* a, const double* b) { return iou_f64_from_corners(a, b); }
#ifdef GM_DISCOVER_INLINE_F64
#define DISC_F64_GENERAL iou_f64_general
#define DISC_F64_CONVEX iou_f64_from_corners
#else
#define DISC_F64_GENERAL discover_f64_general
#define DISC_F64_CONVEX discover_f64_convex
#endif
#ifndef GM_DISCOVER_MINB
#define GM_DISCOVER_MINB 12               // resident CTAs per SM asked of the register allocator: 40 registers (174 uncapped = 2 CTAs of 128 threads); measured in DESIGN.md section 4.4
#endif
template <bool kFusion>
__global__ void __launch_bounds__(128, GM_DISCOVER_MINB)
k_discover(const unsigned long long* __restrict__ skey, const unsigned int* __restrict__ sidx, long long n,
           const QPoly* __restrict__ qp, const QWin* __restrict__ qw, const float4* __restrict__ aabb,
           const double* __restrict__ boxes, const unsigned int* __restrict__ rank,
           const int* __restrict__ scale, unsigned int inactive_group, double thr, Extent* __restrict__ ext,
           Edge* __restrict__ edges, double* __restrict__ edge_iou, long long cap,
           unsigned int* __restrict__ degree) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned long long key = skey[p];
    const unsigned long long g = key >> (2 * CELL_BITS);
    if (g >= (unsigned long long)inactive_group) return;
    const int iy = (int)((key >> CELL_BITS) & CELL_MAX), ix = (int)(key & CELL_MAX);
    const unsigned int i = sidx[p];
    const QPoly A = qp[i];          // this thread's box is the window, every candidate the polygon cut by it
    const QWin Aw = qw[i];
    const float4 ai = aabb[i];
    const unsigned int ri = kFusion ? i : rank[i];
    const int si = kFusion ? scale[i] : 0;
    for (int dy = -1; dy <= 1; ++dy) {
        const int y = iy + dy;
        if (y < 0 || y > CELL_MAX) continue;
        const unsigned long long row = (g << (2 * CELL_BITS)) | ((unsigned long long)y << CELL_BITS);
        const unsigned long long klo = row | (unsigned long long)max(ix - 1, 0);
        const unsigned long long khi = row | (unsigned long long)min(ix + 1, CELL_MAX);
        for (long long q = lower_bound_u64(skey, n, klo); q < n && skey[q] <= khi; ++q) {
            const unsigned int j = sidx[q];
            if (j == i) continue;
            const unsigned int rj = kFusion ? j : rank[j];
            if (rj >= ri) continue;                        // each unordered pair once
            if (kFusion && scale[j] == si) continue;
            if (!aabb_overlap(ai, aabb[j])) continue;
            // the reference's float64 decision: fp32 first, float64 from the raw corners within 1e-4 of the threshold
            const QPoly Pj = qp[j];
            double v = (double)qbox_iou(Pj, A, Aw);
            if ((Pj.valid | A.valid) & 2) {
                // a concave simple quad (hand-made input): valid for shapely, outside the fp32 forms
                v = DISC_F64_GENERAL(boxes + (long long)i * 8, boxes + (long long)j * 8);
            } else if (fabs(v - thr) < 1e-4) {
                v = DISC_F64_CONVEX(boxes + (long long)i * 8, boxes + (long long)j * 8);
                atomicAdd(&g_adjacent_stats[0], 1ull);
                if (fabs(v - thr) < 1e-5) atomicAdd(&g_adjacent_stats[1], 1ull);
            }
            if (!(v >= thr)) continue;
            const unsigned int pos = atomicAdd(&ext->edge_count, 1u);
            if ((long long)pos < cap) {
                Edge e; e.hi = (int)j; e.lo = (int)i;
                edges[pos] = e;
                if (kFusion) {
                    // float64 value for the reference's tie-break on IoU
                    edge_iou[pos] = DISC_F64_CONVEX(boxes + (long long)i * 8, boxes + (long long)j * 8);
                    atomicAdd(&degree[i], 1u);
                    atomicAdd(&degree[j], 1u);
                }
            }
        }
    }
}

// ========================================================================================
// NMS fixpoint.  state: 0 undecided, 1 kept, 2 suppressed, 3 deferred.
//
// `dfr` (optional, cross-band merge): dfr[i] != 0 marks a box whose fate cannot be settled on this rank - it may overlap a
// box of another row band (a seam candidate, k_seam_candidates), or a higher-priority neighbour of it is itself deferred.
// Such a box is still suppressed for good by a KEPT higher-priority neighbour (kept boxes are final: they are decided
// only when every higher-priority neighbour is), otherwise it ends in state 3 once all its higher-priority neighbours
// are decided, and deferral propagates to its lower-priority neighbours.  Boxes that end in 1 or 2 have no deferred
// ancestor, so their state equals the single-rank result; the deferred ones are resolved after the seam exchange.
__global__ void __launch_bounds__(512)
k_nms_fixpoint(const Edge* __restrict__ edges, long long cap, long long n, Extent* __restrict__ ext,
               unsigned char* __restrict__ state, unsigned char* __restrict__ sup,
               unsigned int* __restrict__ blk, unsigned char* __restrict__ dfr) {
    cg::grid_group grid = cg::this_grid();
    const long long n_edges = (long long)ext->edge_count;
    if (n_edges > cap) return;                              // overflow: caller reruns with more room
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (unsigned int round = 1;; ++round) {
        for (long long e = tid; e < n_edges; e += nthreads) {
            const Edge ed = edges[e];
            if (state[ed.lo] == 0) {
                const unsigned char sh = state[ed.hi];
                if (sh == 1) sup[ed.lo] = 1;
                else if (sh == 0) blk[ed.lo] = round;
                else if (sh == 3) dfr[ed.lo] = 1;           // state 3 exists only when dfr is given
            }
        }
        grid.sync();
        unsigned int und = 0;
        for (long long i = tid; i < n; i += nthreads) {
            if (state[i] == 0) {
                if (sup[i]) state[i] = 2;
                else if (blk[i] == round) ++und;
                else state[i] = (dfr && dfr[i]) ? 3 : 1;
            }
        }
        und = __reduce_add_sync(0xffffffffu, und);
        if ((threadIdx.x & 31) == 0 && und) atomicAdd(&ext->undecided[round & 1], und);
        grid.sync();
        const unsigned int left = *((volatile unsigned int*)&ext->undecided[round & 1]);
        if (left == 0u) break;
        if (tid == 0) ext->undecided[(round + 1) & 1] = 0u;
        grid.sync();
    }
}

__global__ void __launch_bounds__(256)
k_keep_flags(const int* __restrict__ order, const unsigned char* __restrict__ state, long long n,
             unsigned int* __restrict__ flag_by_rank, unsigned char* __restrict__ keep) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = order[r];
    const unsigned char k = state[i] == 1;
    flag_by_rank[r] = k;
    if (keep) keep[i] = k;
}

__global__ void __launch_bounds__(256)
k_compact_kept(const int* __restrict__ order, const unsigned int* __restrict__ flag, const unsigned int* __restrict__ pos,
               long long n, int* __restrict__ kept_idx) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (flag[r]) kept_idx[pos[r]] = order[r];
}

__global__ void k_finish_count(const Extent* __restrict__ ext, long long cap, const unsigned int* __restrict__ total,
                               long long* __restrict__ n_out) {
    const long long e = (long long)ext->edge_count;
    *n_out = (e > cap) ? -e : (long long)*total;
}

// ========================================================================================
// Fusion fixpoint.

struct Adj { int nb; int pad; double iou; };

__global__ void __launch_bounds__(256)
k_adj_fill(const Edge* __restrict__ edges, const double* __restrict__ edge_iou, long long cap,
           const Extent* __restrict__ ext, const unsigned int* __restrict__ off,
           unsigned int* __restrict__ cursor, Adj* __restrict__ adj) {
    const long long n_edges = (long long)ext->edge_count;
    if (n_edges > cap) return;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const Edge ed = edges[e];
    Adj a; a.pad = 0; a.iou = edge_iou[e];
    a.nb = ed.lo; adj[off[ed.hi] + atomicAdd(&cursor[ed.hi], 1u)] = a;
    a.nb = ed.hi; adj[off[ed.lo] + atomicAdd(&cursor[ed.lo], 1u)] = a;
}

__global__ void __launch_bounds__(256)
k_fuse_init(const float* __restrict__ conf, long long n, double conf_low, unsigned char* __restrict__ active,
            unsigned char* __restrict__ done, int* __restrict__ emit, unsigned int* __restrict__ degree,
            unsigned int* __restrict__ cursor) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char a = (double)conf[i] >= conf_low;
    active[i] = a;
    done[i] = !a;
    emit[i] = -1;
    degree[i] = 0u;
    cursor[i] = 0u;
}

// done: 0 = not yet visited, 1 = visited (processed, claimed as a partner, or filtered out).
__global__ void __launch_bounds__(512)
k_fuse_fixpoint(const unsigned int* __restrict__ off, const unsigned int* __restrict__ degree,
                const Adj* __restrict__ adj, const float* __restrict__ conf, long long n, long long cap,
                double conf_high, Extent* __restrict__ ext, unsigned char* __restrict__ done,
                unsigned char* __restrict__ ready, int* __restrict__ emit) {
    cg::grid_group grid = cg::this_grid();
    if ((long long)ext->edge_count > cap) return;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (unsigned int round = 1;; ++round) {
        // phase A: i may act once every earlier box within two hops has been visited
        for (long long i = tid; i < n; i += nthreads) {
            if (done[i]) continue;
            bool ok = true;
            const unsigned int o = off[i], d = degree[i];
            for (unsigned int a = 0; a < d && ok; ++a) {
                const int p = adj[o + a].nb;
                if (done[p]) continue;                    // already visited: nobody can claim it any more
                if (p < i) { ok = false; break; }
                const unsigned int op = off[p], dp = degree[p];
                for (unsigned int b = 0; b < dp; ++b) {
                    const int q = adj[op + b].nb;
                    if (q < i && !done[q]) { ok = false; break; }
                }
            }
            ready[i] = ok;
        }
        grid.sync();
        // phase B: ready boxes pick their partner; no two of them share a candidate
        unsigned int und = 0;
        for (long long i = tid; i < n; i += nthreads) {
            if (done[i]) continue;
            if (!ready[i]) { ++und; continue; }
            const unsigned int o = off[i], d = degree[i];
            int best = -1;
            float best_conf = -1.f;
            double best_iou = 0.0;
            for (unsigned int a = 0; a < d; ++a) {
                const Adj e = adj[o + a];
                if (done[e.nb]) continue;
                const float cp = conf[e.nb];
                const bool better = (best < 0) || (cp > best_conf) ||
                                    (cp == best_conf && (e.iou > best_iou || (e.iou == best_iou && e.nb < best)));
                if (better) { best = e.nb; best_conf = cp; best_iou = e.iou; }
            }
            const float ci = conf[i];
            if (best < 0) {
                emit[i] = ((double)ci >= conf_high) ? (int)i : -1;
            } else {
                emit[i] = (ci >= best_conf) ? (int)i : best;
                done[best] = 1;
            }
            done[i] = 1;
        }
        und = __reduce_add_sync(0xffffffffu, und);
        if ((threadIdx.x & 31) == 0 && und) atomicAdd(&ext->undecided[round & 1], und);
        grid.sync();
        const unsigned int left = *((volatile unsigned int*)&ext->undecided[round & 1]);
        if (left == 0u) break;
        if (tid == 0) ext->undecided[(round + 1) & 1] = 0u;
        grid.sync();
    }
}

__global__ void __launch_bounds__(256)
k_emit_flags(const int* __restrict__ emit, long long n, unsigned int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = emit[i] >= 0;
}

__global__ void __launch_bounds__(256)
k_emit_compact(const int* __restrict__ emit, const unsigned int* __restrict__ pos, long long n,
               int* __restrict__ kept_idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && emit[i] >= 0) kept_idx[pos[i]] = emit[i];
}

__global__ void __launch_bounds__(256)
k_iota(int* __restrict__ out, long long n, long long* __restrict__ count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)i;
    if (i == 0) *count = n;
}

// ========================================================================================
// Tile post-processing front end: remap, border filter, angle.

__global__ void __launch_bounds__(256)
k_tile_remap(const float* __restrict__ local, const int* __restrict__ cls, const int* __restrict__ tile_id,
             long long n, const gm_tile* __restrict__ tiles, int n_tiles, int max_class, int margin,
             int angle_class, double* __restrict__ gbox, double* __restrict__ angle,
             int* __restrict__ group, unsigned char* __restrict__ pass) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = tile_id[i];
    const gm_tile tl = tiles[min(max(t, 0), n_tiles - 1)];
    const float* b = local + i * 8;
    double g[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        g[2 * k] = (double)b[2 * k] + (double)tl.x0;
        g[2 * k + 1] = (double)b[2 * k + 1] + (double)tl.y0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) gbox[i * 8 + k] = g[k];
    bool ok = (t >= 0 && t < n_tiles);
    if (margin > 0) {
        // mean of the four map-space corners, then back to tile space (Detect_OBB.py:159-174)
        const double cx = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(g[0], g[2]), g[4]), g[6]), 4.0);
        const double cy = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(g[1], g[3]), g[5]), g[7]), 4.0);
        const double rx = __dsub_rn(cx, (double)tl.x0), ry = __dsub_rn(cy, (double)tl.y0);
        const double m = (double)margin;
        ok = ok && (m <= rx) && (rx <= (double)(tl.w - margin)) && (m <= ry) && (ry <= (double)(tl.h - margin));
    }
    pass[i] = ok;
    const int c = cls[i];
    double ang = 0.0;
    if (c == angle_class) {
        // Detect_OBB.py:135-142 on the TILE-LOCAL points (fp32 values promoted to float64)
        const double a = __dmul_rn(atan2(__dsub_rn((double)b[6], (double)b[0]), __dsub_rn((double)b[7], (double)b[1])),
                                   180.0 / 3.141592653589793);
        ang = a > 0.0 ? __dsub_rn(180.0, a) : fabs(a);
    }
    angle[i] = ang;
    group[i] = t * (max_class + 1) + min(max(c, 0), max_class);
}

__global__ void __launch_bounds__(256)
k_gather_records(const int* __restrict__ kept_idx, const long long* __restrict__ n_kept,
                 const double* __restrict__ gbox, const int* __restrict__ cls, const float* __restrict__ conf,
                 const double* __restrict__ angle, long long n, double* __restrict__ out_boxes,
                 int* __restrict__ out_cls, float* __restrict__ out_conf, double* __restrict__ out_angle,
                 int* __restrict__ out_src) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long m = *n_kept;
    if (k >= m || k >= n) return;
    const int i = kept_idx[k];
#pragma unroll
    for (int c = 0; c < 8; ++c) out_boxes[k * 8 + c] = gbox[(long long)i * 8 + c];
    out_cls[k] = cls[i];
    out_conf[k] = conf[i];
    if (out_angle) out_angle[k] = angle ? angle[i] : 0.0;
    out_src[k] = i;
}

// ========================================================================================
// Workspace layout + drivers.

int g_num_sms = 0;

int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return GM_NUM_SMS_B200;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = GM_NUM_SMS_B200;
        g_num_sms = v;
    }
    return g_num_sms;
}

struct MergeWs {
    Extent* ext;
    QPoly* qp;
    QWin* qw;
    float4* aabb;
    unsigned int* rank;
    SortBufs sort;
    Edge* edges;
    double* edge_iou;
    unsigned char *state, *sup, *flag8, *active, *dfr;
    unsigned int *blk, *flag, *pos, *total, *degree, *cursor, *off;
    Adj* adj;
    int *emit, *order_tmp, *group;
    double *gbox, *angle;
    int* kept_tmp;
    long long* count_tmp;
    size_t bytes;
};

long long default_edge_cap(long long n, long long cap) { return cap > 0 ? cap : 16 * n + 1024; }

MergeWs carve_merge(void* ws, long long n, long long cap, bool fusion, bool tiles) {
    GmArena a(ws, ~(size_t)0);
    MergeWs w{};
    const size_t N = (size_t)(n > 0 ? n : 1);
    w.ext = a.take<Extent>(1);
    w.qp = a.take<QPoly>(N);
    w.qw = a.take<QWin>(N);
    w.aabb = a.take<float4>(N);
    w.rank = a.take<unsigned int>(N);
    w.sort.ka = a.take<unsigned long long>(N);
    w.sort.kb = a.take<unsigned long long>(N);
    w.sort.va = a.take<unsigned int>(N);
    w.sort.vb = a.take<unsigned int>(N);
    const size_t nh = (size_t)(256 * rs_blocks((long long)N));
    w.sort.hist = a.take<unsigned int>(nh);
    w.sort.scan_tmp = a.take<unsigned int>((size_t)(scan_blocks((long long)nh) + scan_blocks((long long)N)) + 2);
    w.sort.dtot = a.take<unsigned int>((size_t)RS_MAX_PASSES * 256);
    w.edges = a.take<Edge>((size_t)cap);
    w.state = a.take<unsigned char>(N);
    w.sup = a.take<unsigned char>(N);
    w.dfr = a.take<unsigned char>(N);
    w.blk = a.take<unsigned int>(N);
    w.flag = a.take<unsigned int>(N);
    w.pos = a.take<unsigned int>(N);
    w.total = a.take<unsigned int>(4);
    w.order_tmp = a.take<int>(N);
    if (fusion) {
        w.edge_iou = a.take<double>((size_t)cap);
        w.active = a.take<unsigned char>(N);
        w.flag8 = a.take<unsigned char>(N);
        w.degree = a.take<unsigned int>(N);
        w.cursor = a.take<unsigned int>(N);
        w.off = a.take<unsigned int>(N);
        w.adj = a.take<Adj>((size_t)(2 * cap));
        w.emit = a.take<int>(N);
    }
    if (tiles) {
        w.group = a.take<int>(N);
        w.gbox = a.take<double>(8 * N);
        w.angle = a.take<double>(N);
        w.active = a.take<unsigned char>(N);
        w.kept_tmp = a.take<int>(N);
        w.count_tmp = a.take<long long>(1);
    }
    w.bytes = gm_align_up(a.off, 256);
    return w;
}

int bits_for(unsigned long long max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0ull) ++b;
    return b;
}

template <typename K, typename... Args>
int launch_cooperative(K kernel, int threads, long long work_items, cudaStream_t s, Args... args) {
    int per_sm = 0;
    GM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
    if (per_sm < 1) return GM_ERANGE;
    long long blocks = (long long)num_sms() * per_sm;
    const long long need = (work_items + threads - 1) / threads;
    if (blocks > need) blocks = need;
    // A cooperative grid needs all its CTAs resident at once; beside a saturating kernel on another stream the
    // scheduler has to drain that many SM slots first.  The fixpoints are latency bound (a grid barrier per sweep),
    // so a small grid costs them little and disturbs the neighbour less (GM_COOP_MAX_BLOCKS; measured in DESIGN.md).
    static const int max_blocks = gm_env_int("GM_COOP_MAX_BLOCKS", GM_COOP_DEFAULT_MAX_BLOCKS);
    if (max_blocks > 0 && blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    void* argv[] = {(void*)&args...};
    GM_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)blocks), dim3((unsigned)threads), argv, 0, s));
    gm_note_launches(1);
    return GM_OK;
}

__global__ void __launch_bounds__(256)
k_mask_inactive(const unsigned char* __restrict__ active, long long n, unsigned char* __restrict__ state) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !active[i]) state[i] = 2;
}

// Seam candidates of the cross-band merge (sharding.py): a box of this rank can overlap a box of ANOTHER rank only if its
// AABB, grown by `bound` on every side, reaches a rectangle where foreign box centres can lie (the safe regions of the
// foreign tiles, Detect_OBB.py:167-174) - given that no corner of any box anywhere is farther than `bound` (Chebyshev)
// from that box's centre: a foreign box lies inside centre +- bound, so if it meets this box's AABB its centre lies
// inside the grown AABB.  Every rank checks its own boxes and a violation travels with the exchange, so all ranks agree.
struct SeamCfg {
    int n_rects;
    float bound;
    float4 rects[8];           // closed rectangles (x0, y0, x1, y1) of foreign box centres
};

__global__ void __launch_bounds__(256)
k_seam_candidates(const float4* __restrict__ aabb, long long n, SeamCfg cfg, unsigned char* __restrict__ dfr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = aabb[i];
    bool c = false;
    for (int k = 0; k < cfg.n_rects; ++k) {
        const float4 r = cfg.rects[k];
        c |= (a.x - cfg.bound <= r.z) && (r.x <= a.z + cfg.bound) && (a.y - cfg.bound <= r.w) && (r.y <= a.w + cfg.bound);
    }
    dfr[i] = c ? 1 : 0;                                  // a non-finite AABB (blank row) compares false everywhere
}

// Shared engine: exact greedy NMS inside groups.  `major` (optional) is the leading sort key
// of the output order (tile id); `group` decides which boxes can suppress each other.
// Part 1: priorities, candidate pairs, fixpoint -> w.state (1 kept, 2 suppressed, 3 deferred) and the stable
// confidence order in `order`.  `seam` (optional): boxes that may overlap another rank's boxes are deferred (w.dfr).
int nms_resolve(const double* boxes, const int* group, unsigned int max_group, const int* major,
                unsigned int max_major, const float* conf, const unsigned char* active, long long n,
                double thr, long long cap, int* order, const SeamCfg* seam, MergeWs& w, cudaStream_t s) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_extent_init<<<1, 1, 0, s>>>(w.ext); gm_note_launches(1);
    k_prepare<<<blocks, 256, 0, s>>>(boxes, conf, major, n, w.qp, w.qw, w.aabb, w.ext, w.sort.ka, w.sort.va); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    int where = radix_sort_pairs(w.sort, n, 32 + (major ? bits_for(max_major) : 0), s);
    if (where < 0) return GM_EINVAL;
    k_ranks<<<blocks, 256, 0, s>>>(where ? w.sort.vb : w.sort.va, n, w.rank, order); gm_note_launches(1);
    const unsigned int inactive_group = max_group + 1u;
    k_cell_keys<<<blocks, 256, 0, s>>>(w.aabb, group, active, inactive_group, n, w.ext, w.sort.ka, w.sort.va); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    where = radix_sort_pairs(w.sort, n, 2 * CELL_BITS + bits_for(inactive_group), s);
    if (where < 0) return GM_EINVAL;
    const unsigned long long* skey = where ? w.sort.kb : w.sort.ka;
    const unsigned int* sidx = where ? w.sort.vb : w.sort.va;
    GM_CUDA_TRY(cudaMemsetAsync(w.state, 0, (size_t)n, s));
    GM_CUDA_TRY(cudaMemsetAsync(w.sup, 0, (size_t)n, s));
    GM_CUDA_TRY(cudaMemsetAsync(w.blk, 0, (size_t)n * sizeof(unsigned int), s));
    if (active) { k_mask_inactive<<<blocks, 256, 0, s>>>(active, n, w.state); gm_note_launches(1); }
    if (seam) { k_seam_candidates<<<blocks, 256, 0, s>>>(w.aabb, n, *seam, w.dfr); gm_note_launches(1); }
    k_discover<false><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(skey, sidx, n, w.qp, w.qw, w.aabb, boxes, w.rank, nullptr,
                                                                 inactive_group, thr, w.ext, w.edges, nullptr, cap, nullptr); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    {
        const Edge* e = w.edges; Extent* x = w.ext; unsigned char* st = w.state; unsigned char* sp = w.sup;
        unsigned int* bk = w.blk; unsigned char* df = seam ? w.dfr : nullptr;
        int rc = launch_cooperative(k_nms_fixpoint, 512, n > cap ? n : cap, s, e, cap, n, x, st, sp, bk, df);
        if (rc != GM_OK) return rc;
    }
    return GM_OK;
}

// Part 2: kept boxes (state 1) in the stable confidence order -> keep flags, compacted index list, count.
int nms_compact(const int* order, unsigned char* keep_out, int* kept_idx, long long* n_kept, long long n, long long cap,
                MergeWs& w, cudaStream_t s) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    // boxes excluded up front (inactive) must not be kept: state 1 only if active
    k_keep_flags<<<blocks, 256, 0, s>>>(order, w.state, n, w.flag, keep_out); gm_note_launches(1);
    int st = exclusive_scan_u32(w.flag, w.pos, n, w.sort.scan_tmp, w.total, s);
    if (st != GM_OK) return st;
    k_compact_kept<<<blocks, 256, 0, s>>>(order, w.flag, w.pos, n, kept_idx); gm_note_launches(1);
    k_finish_count<<<1, 1, 0, s>>>(w.ext, cap, w.total, n_kept); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

// ========================================================================================
// Per-tile NMS with a bounded number of detections per tile (the pipeline's per-tile stage: a detector emits at most
// max_det = 300 boxes per tile): one CTA per tile instead of the grid / sort / edge-list / fixpoint engine.
//   1. the active boxes are counted and scattered per tile (their order inside a tile is restored by step 2);
//   2. the CTA ranks its boxes by (confidence desc, input index asc) - the reference's stable in-place sort - by counting;
//   3. every same-class pair with overlapping AABBs gets the engine's float64-decided `IoU >= thr` test (fp32 first,
//      float64 within 1e-4 of the threshold or for a concave quad), one bit per ordered pair in shared memory;
//   4. one warp sweeps the score-sorted list: a box not yet removed is kept and ORs its row into the removed mask -
//      exactly the sequential greedy rule of merge_detections (Detect_OBB.py:183-198).
// A tile with more boxes than the bit matrix holds (TN_CAP) is still exact: the same CTA ranks by counting over global
// memory and runs the greedy rule box by box against the boxes kept so far (slow; only reached when the caller's bound
// does not hold).  Output = the engine's: kept input indices in (tile, confidence desc) order.
constexpr int TN_CAP = 320;
constexpr int TN_WORDS = TN_CAP / 32;
constexpr int TN_THREADS = 128;
constexpr int TN_CHUNK = 2048;            // candidate pairs held per round (a round covers as many boxes b as can never overflow it)

__global__ void __launch_bounds__(256)
k_tn_count(const int* __restrict__ tile_id, const unsigned char* __restrict__ active, long long n,
           unsigned int* __restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && active[i]) atomicAdd(&cnt[tile_id[i]], 1u);          // active implies a valid tile id (k_tile_remap)
}

__global__ void __launch_bounds__(256)
k_tn_scatter(const int* __restrict__ tile_id, const unsigned char* __restrict__ active, long long n,
             const unsigned int* __restrict__ off, unsigned int* __restrict__ cursor, unsigned int* __restrict__ idx_by_tile) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && active[i]) {
        const int t = tile_id[i];
        idx_by_tile[off[t] + atomicAdd(&cursor[t], 1u)] = (unsigned int)i;
    }
}

// the engine's decision for one pair (k_discover): `win` is the lower-priority box (the window), `pol` the higher one
__device__ __forceinline__ bool tn_pair_reaches(const QPoly* __restrict__ qp, const QWin* __restrict__ qw,
                                                const double* __restrict__ boxes, unsigned int win, unsigned int pol, double thr) {
    const QPoly A = qp[win];
    const QWin Aw = qw[win];
    const QPoly Pj = qp[pol];
    double v = (double)qbox_iou(Pj, A, Aw);
    if ((Pj.valid | A.valid) & 2) {
        v = DISC_F64_GENERAL(boxes + (long long)win * 8, boxes + (long long)pol * 8);
    } else if (fabs(v - thr) < 1e-4) {
        v = DISC_F64_CONVEX(boxes + (long long)win * 8, boxes + (long long)pol * 8);
        atomicAdd(&g_adjacent_stats[0], 1ull);
        if (fabs(v - thr) < 1e-5) atomicAdd(&g_adjacent_stats[1], 1ull);
    }
    return v >= thr;
}

__global__ void __launch_bounds__(TN_THREADS, 8)
k_tile_nms(const unsigned int* __restrict__ cnt, const unsigned int* __restrict__ off,
           const unsigned int* __restrict__ idx_by_tile, const unsigned long long* __restrict__ key,
           const int* __restrict__ group, const QPoly* __restrict__ qp, const QWin* __restrict__ qw,
           const float4* __restrict__ aabb, const double* __restrict__ boxes, double thr,
           int* __restrict__ order, unsigned int* __restrict__ flag) {
    __shared__ unsigned int u_idx[TN_CAP], s_idx[TN_CAP];
    __shared__ unsigned long long u_key[TN_CAP];
    __shared__ int s_grp[TN_CAP];
    __shared__ float4 s_aabb[TN_CAP];
    __shared__ unsigned int bits[TN_CAP][TN_WORDS];
    __shared__ unsigned int cand[TN_CHUNK];              // candidate pairs of one chunk: a | b << 16
    __shared__ unsigned int n_cand, seg_max;
    __shared__ unsigned short s_pos[TN_CAP], s_pos2[TN_CAP];   // output position of a box: in load order / in NMS order
    __shared__ int s_grp2[TN_CAP];                       // class key in NMS order
    __shared__ unsigned int n_key[TN_CAP];               // (class, output rank)
    __shared__ unsigned int rem[TN_WORDS];               // removed mask of the sweep
    const int m = (int)cnt[blockIdx.x];
    if (m == 0) return;
    const unsigned int base = off[blockIdx.x];
    const int tid = threadIdx.x;
    if (m <= TN_CAP) {
        // keys inside a tile: (confidence key, input index) in one 64-bit word - the tile part of the engine's sort key is
        // the same for all of them - so "before" is one compare
        for (int k = tid; k < m; k += TN_THREADS) {
            const unsigned int i = idx_by_tile[base + k];
            u_idx[k] = i;
            u_key[k] = (key[i] << 32) | (unsigned long long)i;
            s_grp[k] = group[i];                                  // unsorted for now (u_* order)
        }
        if (tid == 0) seg_max = 1u;
        if (tid < TN_WORDS) rem[tid] = 0u;
        __syncthreads();
        // rank 1 by counting: the place in the OUTPUT order (confidence desc, input index asc - the reference's stable sort)
        for (int k = tid; k < m; k += TN_THREADS) {
            const unsigned long long kk = u_key[k];
            int r_out = 0;
            for (int j = 0; j < m; ++j) r_out += u_key[j] < kk;
            s_pos[k] = (unsigned short)r_out;                     // still in u_* order
            order[base + r_out] = (int)u_idx[k];
        }
        __syncthreads();
        // rank 2: the place in the NMS order (class first, then the output order): classes are independent, so the greedy
        // rule may walk class by class, and same-class boxes are then neighbours - ~m * 8 pair tests instead of m * m.
        // (class, output rank) is a 32-bit key: classes of one tile differ by less than 2^15 (group = tile * C + class)
        const int g0 = s_grp[0];
        for (int k = tid; k < m; k += TN_THREADS) n_key[k] = ((unsigned int)(s_grp[k] - g0 + 32768) << 16) | (unsigned int)s_pos[k];
        __syncthreads();
        for (int k = tid; k < m; k += TN_THREADS) {
            const unsigned int nk = n_key[k];
            int r_nms = 0;
            for (int j = 0; j < m; ++j) r_nms += n_key[j] < nk;
            s_idx[r_nms] = u_idx[k];
            s_pos2[r_nms] = s_pos[k];
            s_grp2[r_nms] = s_grp[k];
        }
        __syncthreads();
        for (int r = tid; r < m; r += TN_THREADS) {
            s_aabb[r] = aabb[s_idx[r]];
#pragma unroll
            for (int w = 0; w < TN_WORDS; ++w) bits[r][w] = 0u;
            // length of the class segment ending at r (only its last box reports): bounds the candidates of a box
            if (r + 1 == m || s_grp2[r + 1] != s_grp2[r]) {
                int a = r;
                while (a > 0 && s_grp2[a - 1] == s_grp2[r]) --a;
                atomicMax(&seg_max, (unsigned int)(r - a + 1));
            }
        }
        __syncthreads();
        // candidates (a before b in one class segment, AABBs overlap) compacted per block of b's, then one IoU per thread
        const int nb = max(1, TN_CHUNK / (int)seg_max);
        for (int b0 = 0; b0 < m; b0 += nb) {
            if (tid == 0) n_cand = 0u;
            __syncthreads();
            const int b1 = min(m, b0 + nb);
            for (int bb = b0 + tid; bb < b1; bb += TN_THREADS) {
                const int gb = s_grp2[bb];
                const float4 ab = s_aabb[bb];
                for (int a = bb - 1; a >= 0 && s_grp2[a] == gb; --a)
                    if (aabb_overlap(ab, s_aabb[a])) cand[atomicAdd(&n_cand, 1u)] = (unsigned int)a | ((unsigned int)bb << 16);
            }
            __syncthreads();
            const int nc = (int)n_cand;
            for (int c = tid; c < nc; c += TN_THREADS) {
                const int a = (int)(cand[c] & 0xffffu), bb = (int)(cand[c] >> 16);
                if (tn_pair_reaches(qp, qw, boxes, s_idx[bb], s_idx[a], thr)) atomicOr(&bits[a][bb >> 5], 1u << (bb & 31));
            }
            __syncthreads();
        }
        // the greedy sweep, one thread per class segment (a row of the bit matrix only has bits inside its own segment, so the
        // segments never touch each other's bits of the shared removed mask)
        for (int r = tid; r < m; r += TN_THREADS) {
            if (r > 0 && s_grp2[r - 1] == s_grp2[r]) continue;      // not the head of a segment
            const int g = s_grp2[r];
            for (int a = r; a < m && s_grp2[a] == g; ++a) {
                const unsigned int word = *reinterpret_cast<volatile unsigned int*>(&rem[a >> 5]);
                if (!((word >> (a & 31)) & 1u)) {
                    flag[base + s_pos2[a]] = 1u;
                    for (int w = a >> 5; w < TN_WORDS; ++w) {
                        const unsigned int row = bits[a][w];
                        if (row) atomicOr(&rem[w], row);
                    }
                }
            }
        }
        return;
    }
    // ---- more boxes than the bit matrix holds: same rule, everything in global memory
    for (int k = tid; k < m; k += TN_THREADS) {
        const unsigned int ik = idx_by_tile[base + k];
        const unsigned long long kk = key[ik];
        int r = 0;
        for (int j = 0; j < m; ++j) {
            const unsigned int ij = idx_by_tile[base + j];
            const unsigned long long kj = key[ij];
            r += (kj < kk) || (kj == kk && ij < ik);
        }
        order[base + r] = (int)ik;
    }
    __syncthreads();
    for (int a = 0; a < m; ++a) {
        const unsigned int ia = (unsigned int)order[base + a];
        const int ga = group[ia];
        const float4 ba = aabb[ia];
        int hit = 0;
        for (int b = tid; b < a && !hit; b += TN_THREADS) {
            if (!flag[base + b]) continue;
            const unsigned int ib = (unsigned int)order[base + b];
            if (group[ib] == ga && aabb_overlap(ba, aabb[ib]) && tn_pair_reaches(qp, qw, boxes, ia, ib, thr)) hit = 1;
        }
        const int any = __syncthreads_or(hit);
        if (!any && tid == 0) flag[base + a] = 1u;
        __syncthreads();
    }
}

int tile_nms_bounded(const double* boxes, const int* group, const int* tile_id, const float* conf,
                     const unsigned char* active, long long n, int n_tiles, double thr, long long cap,
                     int* kept_idx, long long* n_kept, MergeWs& w, cudaStream_t s) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    unsigned int* cnt = w.blk;                 // [n_tiles + 1]  (callers guarantee n_tiles + 1 <= n)
    unsigned int* off = w.rank;                // exclusive scan of cnt
    unsigned int* cursor = w.sort.vb;
    unsigned int* idx_by_tile = w.sort.va;     // free once k_prepare has written its (unused) identity values
    k_extent_init<<<1, 1, 0, s>>>(w.ext); gm_note_launches(1);
    k_prepare<<<blocks, 256, 0, s>>>(boxes, conf, tile_id, n, w.qp, w.qw, w.aabb, w.ext, w.sort.ka, w.sort.va); gm_note_launches(1);
    GM_CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)(n_tiles + 1) * sizeof(unsigned int), s));
    GM_CUDA_TRY(cudaMemsetAsync(cursor, 0, (size_t)n_tiles * sizeof(unsigned int), s));
    GM_CUDA_TRY(cudaMemsetAsync(w.flag, 0, (size_t)n * sizeof(unsigned int), s));
    k_tn_count<<<blocks, 256, 0, s>>>(tile_id, active, n, cnt); gm_note_launches(1);
    int st = exclusive_scan_u32(cnt, off, n_tiles, w.sort.scan_tmp, nullptr, s);
    if (st != GM_OK) return st;
    k_tn_scatter<<<blocks, 256, 0, s>>>(tile_id, active, n, off, cursor, idx_by_tile); gm_note_launches(1);
    k_tile_nms<<<(unsigned)n_tiles, TN_THREADS, 0, s>>>(cnt, off, idx_by_tile, w.sort.ka, group, w.qp, w.qw, w.aabb, boxes, thr,
                                                        w.order_tmp, w.flag); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    st = exclusive_scan_u32(w.flag, w.pos, n, w.sort.scan_tmp, w.total, s);
    if (st != GM_OK) return st;
    k_compact_kept<<<blocks, 256, 0, s>>>(w.order_tmp, w.flag, w.pos, n, kept_idx); gm_note_launches(1);
    k_finish_count<<<1, 1, 0, s>>>(w.ext, cap, w.total, n_kept); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

int nms_engine(const double* boxes, const int* group, unsigned int max_group, const int* major,
               unsigned int max_major, const float* conf, const unsigned char* active, long long n,
               double thr, long long cap, int* order_out, unsigned char* keep_out, int* kept_idx,
               long long* n_kept, MergeWs& w, cudaStream_t s) {
    int* order = order_out ? order_out : w.order_tmp;
    int st = nms_resolve(boxes, group, max_group, major, max_major, conf, active, n, thr, cap, order, nullptr, w, s);
    if (st != GM_OK) return st;
    return nms_compact(order, keep_out, kept_idx, n_kept, n, cap, w, s);
}


// ========================================================================================
// Cross-band exchange records (multi-GPU merge, sharding.merge_bands_device): the per-tile-NMS survivors of a
// rank travel as fixed-size 80-byte records {corners double[8], class int32, confidence float, angle double};
// rows beyond the rank's count are blank (NaN corners, class -1, confidence -inf).  Three kernels replace the
// two dozen tensor operations that blanked, packed, sliced, masked and compacted the same data.

__global__ void __launch_bounds__(256)
k_band_pack(const double* __restrict__ boxes, const int* __restrict__ cls, const float* __restrict__ conf,
            const double* __restrict__ angle, long long n_in, const long long* __restrict__ count,
            long long capacity, unsigned char* __restrict__ rec) {
    // one thread per 8-byte word of a record: words 0..7 corners, word 8 = {class, confidence}, word 9 angle
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= capacity * 10) return;
    const long long row = g / 10;
    const int wd = (int)(g - row * 10);
    long long cnt = *count;
    if (cnt > n_in) cnt = n_in;
    const bool live = row < cnt;                              // a negative count (overflow upstream) blanks everything
    unsigned long long v;
    if (wd < 8) v = live ? (unsigned long long)__double_as_longlong(boxes[row * 8 + wd]) : 0x7ff8000000000000ULL;
    else if (wd == 8) {
        const unsigned int c = live ? (unsigned int)cls[row] : 0xffffffffu;
        const unsigned int f = live ? __float_as_uint(conf[row]) : 0xff800000u;
        v = (unsigned long long)c | ((unsigned long long)f << 32);
    } else v = live ? (unsigned long long)__double_as_longlong(angle ? angle[row] : 0.0) : 0ULL;
    reinterpret_cast<unsigned long long*>(rec)[g] = v;
}

__global__ void __launch_bounds__(256)
k_band_unpack(const unsigned char* __restrict__ rec, long long total, int world, int rank,
              double* __restrict__ boxes, int* __restrict__ cls, int* __restrict__ cls_owned,
              float* __restrict__ conf, double* __restrict__ angle, long long* __restrict__ n_valid) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int valid = 0;
    if (g < total * 10) {
        const long long row = g / 10;
        const int wd = (int)(g - row * 10);
        const unsigned long long v = reinterpret_cast<const unsigned long long*>(rec)[g];
        if (wd < 8) boxes[row * 8 + wd] = __longlong_as_double((long long)v);
        else if (wd == 8) {
            const int c = (int)(unsigned int)(v & 0xffffffffULL);
            cls[row] = c;
            cls_owned[row] = (c >= 0 && (c % world) == rank) ? c : -1;
            conf[row] = __uint_as_float((unsigned int)(v >> 32));
            valid = c >= 0;
        } else if (angle) angle[row] = __longlong_as_double((long long)v);
    }
    const unsigned int votes = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(reinterpret_cast<unsigned long long*>(n_valid), (unsigned long long)__popc(votes));
}

__global__ void __launch_bounds__(256)
k_band_mask_keep(unsigned char* __restrict__ keep, const int* __restrict__ cls_owned, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total && cls_owned[i] < 0) keep[i] = 0;
}

__global__ void __launch_bounds__(256)
k_band_flags(const int* __restrict__ order, const unsigned char* __restrict__ keep, long long total,
             unsigned int* __restrict__ flag) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < total) flag[r] = keep[order[r]] ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
k_band_extract(const int* __restrict__ order, const unsigned int* __restrict__ flag, const unsigned int* __restrict__ pos,
               long long total, const double* __restrict__ boxes, const int* __restrict__ cls,
               const float* __restrict__ conf, const double* __restrict__ angle,
               double* __restrict__ o_boxes, int* __restrict__ o_cls, float* __restrict__ o_conf,
               double* __restrict__ o_angle, long long* __restrict__ o_index) {
    // 8 threads per confidence rank: one corner coordinate each; the first of them also moves the scalars
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r = g >> 3;
    const int k = (int)(g & 7);
    if (r >= total || !flag[r]) return;
    const long long src = order[r], dst = pos[r];
    o_boxes[dst * 8 + k] = boxes[src * 8 + k];
    if (k == 0) {
        o_cls[dst] = cls[src];
        o_conf[dst] = conf[src];
        if (o_angle) o_angle[dst] = angle ? angle[src] : 0.0;
        o_index[dst] = src;
    }
}

__global__ void k_band_count(const unsigned int* __restrict__ total, long long* __restrict__ n_out) { *n_out = (long long)*total; }


// ========================================================================================
// Seam-band exchange (multi-GPU merge, sharding.merge_bands_seam_*).  A rank resolves the global NMS of its OWN band
// locally; only the boxes whose fate depends on another band (state 3 after the deferring fixpoint) travel, as 80-byte
// records {corners double[8] | class int32, confidence float | source row int32, 0}.  Row 0 of a rank's block is a header
// {int64 seam rows, int64 status bits, double largest box extent, int64 survivors}.

constexpr long long SEAM_WORDS = GM_BAND_RECORD_BYTES / 8;      // 10 eight-byte words per record

__global__ void __launch_bounds__(256)
k_band_blank(double* __restrict__ boxes, int* __restrict__ cls, float* __restrict__ conf, long long n_rows,
             const long long* __restrict__ count) {
    // rows at or beyond the device count are not data: NaN corners, class -1 (inactive group), confidence -inf
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows * 8) return;
    const long long row = g >> 3;
    long long cnt = *count;
    if (cnt < 0) cnt = 0;
    if (row < cnt) return;
    boxes[g] = __longlong_as_double(0x7ff8000000000000LL);
    if ((g & 7) == 0) { cls[row] = -1; conf[row] = __uint_as_float(0xff800000u); }
}

__global__ void __launch_bounds__(256)
k_state_flags(const unsigned char* __restrict__ state, const int* __restrict__ cls, long long n, unsigned char want,
              unsigned int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (state[i] == want && cls[i] >= 0) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
k_state_keep_masked(const unsigned char* __restrict__ state, const int* __restrict__ cls, long long n,
                    unsigned char* __restrict__ out) {
    // the local verdicts that outlive the engine workspace; blank rows (class -1) are never kept
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cls[i] < 0 ? 2 : state[i];
}

__global__ void __launch_bounds__(256)
k_seam_pack(const double* __restrict__ boxes, const int* __restrict__ cls, const float* __restrict__ conf,
            const unsigned int* __restrict__ flag, const unsigned int* __restrict__ pos, const unsigned int* __restrict__ total,
            long long n, long long seam_cap, const Extent* __restrict__ ext, long long edge_cap, float bound,
            const long long* __restrict__ count, unsigned long long* __restrict__ rec) {
    // thread per (record row 1..seam_cap or source row, word): first the blank tail + header, then the live rows
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)*total;
    if (g < (seam_cap + 1) * SEAM_WORDS) {
        const long long row = g / SEAM_WORDS;
        const int wd = (int)(g - row * SEAM_WORDS);
        if (row == 0) {
            unsigned long long v = 0ull;
            if (wd == 0) v = (unsigned long long)tot;
            else if (wd == 1) {
                unsigned long long st = 0ull;
                if ((long long)ext->edge_count > edge_cap) st |= GM_SEAM_EDGE_OVERFLOW;
                if (tot > seam_cap) st |= GM_SEAM_CAPACITY_OVERFLOW;
                if (dec_f32(ext->maxrad) > bound) st |= GM_SEAM_EXTENT_EXCEEDED;
                if (*count < 0) st |= GM_SEAM_INPUT_OVERFLOW;
                v = st;
            } else if (wd == 2) v = (unsigned long long)__double_as_longlong((double)dec_f32(ext->maxrad));
            else if (wd == 3) v = (unsigned long long)(*count < 0 ? 0 : *count);
            rec[g] = v;
        } else if (row - 1 >= tot) {
            rec[g] = wd < 8 ? 0x7ff8000000000000ULL : (wd == 8 ? (0xffffffffULL | (0xff800000ULL << 32)) : 0xffffffffULL);
        }
    }
    if (g < n * SEAM_WORDS) {
        const long long i = g / SEAM_WORDS;
        const int wd = (int)(g - i * SEAM_WORDS);
        if (flag[i] && (long long)pos[i] < seam_cap) {
            unsigned long long v;
            if (wd < 8) v = (unsigned long long)__double_as_longlong(boxes[i * 8 + wd]);
            else if (wd == 8) v = (unsigned long long)(unsigned int)cls[i] | ((unsigned long long)__float_as_uint(conf[i]) << 32);
            else v = (unsigned long long)(unsigned int)i;
            rec[((long long)pos[i] + 1) * SEAM_WORDS + wd] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
k_seam_headers(const unsigned long long* __restrict__ rec, int world, long long seam_cap, long long* __restrict__ meta) {
    // meta[1] |= status of every rank, meta[2] += seam rows of every rank, meta[3] += survivors of every rank
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= world) return;
    const unsigned long long* h = rec + (long long)r * (seam_cap + 1) * SEAM_WORDS;
    atomicAdd(reinterpret_cast<unsigned long long*>(meta + 2), h[0]);
    if (h[1]) atomicOr(reinterpret_cast<unsigned long long*>(meta + 1), h[1]);
    atomicAdd(reinterpret_cast<unsigned long long*>(meta + 3), h[3]);
}

__global__ void __launch_bounds__(256)
k_seam_unpack(const unsigned long long* __restrict__ rec, int block_begin, int n_blocks, long long seam_cap,
              double* __restrict__ boxes, int* __restrict__ cls, float* __restrict__ conf, int* __restrict__ src) {
    // the record rows (not the headers) of rank blocks [block_begin, block_begin + n_blocks) -> SoA rows, block-major
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rows_words = seam_cap * SEAM_WORDS;
    if (g >= (long long)n_blocks * rows_words) return;
    const long long b = g / rows_words;
    const long long in = g - b * rows_words;
    const long long row = in / SEAM_WORDS;
    const int wd = (int)(in - row * SEAM_WORDS);
    const unsigned long long v = rec[((long long)(block_begin + b) * (seam_cap + 1) + 1 + row) * SEAM_WORDS + wd];
    const long long u = b * seam_cap + row;
    if (wd < 8) boxes[u * 8 + wd] = __longlong_as_double((long long)v);
    else if (wd == 8) { cls[u] = (int)(unsigned int)(v & 0xffffffffULL); conf[u] = __uint_as_float((unsigned int)(v >> 32)); }
    else src[u] = (int)(unsigned int)(v & 0xffffffffULL);
}

__global__ void __launch_bounds__(256)
k_seam_apply(const unsigned char* __restrict__ state_u, const int* __restrict__ cls_u, const int* __restrict__ src_u,
             long long first, long long seam_cap, long long n_local, unsigned char* __restrict__ state_local,
             long long* __restrict__ meta) {
    // the seam verdict on this rank's own deferred rows.  A row still deferred (state 3: only possible when the seam set
    // was restricted to neighbouring ranks and a chain of overlaps leaves that neighbourhood) has no verdict: flagged.
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= seam_cap) return;
    const long long u = first + k;
    if (cls_u[u] < 0) return;
    const int i = src_u[u];
    const unsigned char st = state_u[u];
    if (st == 3) atomicOr(reinterpret_cast<unsigned long long*>(meta + 1), (unsigned long long)GM_SEAM_CHAIN_ESCAPES);
    if (i >= 0 && i < n_local) state_local[i] = st == 1 ? 1 : 2;
}

__global__ void k_seam_meta(const unsigned int* __restrict__ total, const Extent* __restrict__ ext, long long edge_cap,
                            long long* __restrict__ meta) {
    meta[0] = (long long)*total;
    if ((long long)ext->edge_count > edge_cap) meta[1] |= (long long)GM_SEAM_EDGE_OVERFLOW;
}

struct SeamWs {
    unsigned char* p_state;      // local verdicts, alive between the two calls
    int* p_order;                // local stable confidence order
    double* u_boxes; int* u_cls; float* u_conf; int* u_src;
    int* kept_tmp;
    void* engine;                // carve_merge region for max(local rows, gathered seam rows)
    size_t engine_bytes, bytes;
};

SeamWs carve_seam(void* ws, long long n, long long n_u, long long cap) {
    GmArena a(ws, ~(size_t)0);
    SeamWs w{};
    const size_t N = (size_t)(n > 0 ? n : 1), U = (size_t)(n_u > 0 ? n_u : 1);
    w.p_state = a.take<unsigned char>(N);
    w.p_order = a.take<int>(N);
    w.u_boxes = a.take<double>(8 * U);
    w.u_cls = a.take<int>(U);
    w.u_conf = a.take<float>(U);
    w.u_src = a.take<int>(U);
    w.kept_tmp = a.take<int>(N);
    a.off = gm_align_up(a.off, 256);
    w.engine = ws ? (void*)((uint8_t*)ws + a.off) : nullptr;
    w.engine_bytes = carve_merge(nullptr, (long long)(N > U ? N : U), cap, false, false).bytes;
    w.bytes = gm_align_up(a.off + w.engine_bytes, 256);
    return w;
}

}  // namespace

// ------------------------------------------------------------------------------------------

extern "C" int gm_threshold_adjacent_stats(uint64_t* counts2_host, int32_t reset) {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
    if (counts2_host) GM_CUDA_TRY(cudaMemcpyFromSymbol(counts2_host, g_adjacent_stats, 2 * sizeof(uint64_t)));
    if (reset) {
        const uint64_t z[2] = {0u, 0u};
        GM_CUDA_TRY(cudaMemcpyToSymbol(g_adjacent_stats, z, sizeof(z)));
    }
    return GM_OK;
}

extern "C" size_t gm_nms_workspace_bytes(int64_t n, int64_t edge_capacity) {
    if (n < 0) return 0;
    return carve_merge(nullptr, n, default_edge_cap(n, edge_capacity), false, false).bytes;
}

extern "C" int gm_nms_global(const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev, int64_t n,
                             int32_t max_class, double iou_thr, int64_t edge_capacity,
                             int32_t* order_dev, uint8_t* keep_dev, int32_t* kept_idx_dev, int64_t* n_kept_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (n < 0 || max_class < 0 || !n_kept_dev) return GM_EINVAL;
    cudaStream_t s = gm_stream(stream);
    if (n == 0) { GM_CUDA_TRY(cudaMemsetAsync(n_kept_dev, 0, sizeof(int64_t), s)); return GM_OK; }
    if (!boxes_dev || !cls_dev || !conf_dev || !kept_idx_dev || !workspace_dev) return GM_EINVAL;
    if (n > (1LL << 30)) return GM_ERANGE;
    const long long cap = default_edge_cap(n, edge_capacity);
    if (workspace_bytes < gm_nms_workspace_bytes(n, edge_capacity)) return GM_ENOSPC;
    MergeWs w = carve_merge(workspace_dev, n, cap, false, false);
    return nms_engine(boxes_dev, cls_dev, (unsigned)max_class, nullptr, 0u, conf_dev, nullptr, n, iou_thr, cap,
                      order_dev, keep_dev, kept_idx_dev, reinterpret_cast<long long*>(n_kept_dev), w, s);
}

extern "C" size_t gm_tile_postprocess_workspace_bytes(int64_t n, int64_t edge_capacity) {
    if (n < 0) return 0;
    return carve_merge(nullptr, n, default_edge_cap(n, edge_capacity), false, true).bytes;
}

extern "C" int gm_tile_postprocess(const float* boxes_local_dev, const int32_t* cls_dev, const float* conf_dev,
                                   const int32_t* tile_id_dev, int64_t n,
                                   const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_class,
                                   int32_t margin_px, int32_t angle_class, double iou_merge, int64_t edge_capacity,
                                   double* out_boxes_dev, int32_t* out_cls_dev, float* out_conf_dev,
                                   double* out_angle_dev, int32_t* out_src_dev, int64_t* out_count_dev,
                                   void* workspace_dev, size_t workspace_bytes, void* stream) {
    return gm_tile_postprocess_bounded(boxes_local_dev, cls_dev, conf_dev, tile_id_dev, n, tiles_dev, n_tiles, max_class,
                                       margin_px, angle_class, iou_merge, edge_capacity, 0, out_boxes_dev, out_cls_dev,
                                       out_conf_dev, out_angle_dev, out_src_dev, out_count_dev, workspace_dev,
                                       workspace_bytes, stream);
}

extern "C" int gm_tile_postprocess_bounded(const float* boxes_local_dev, const int32_t* cls_dev, const float* conf_dev,
                                           const int32_t* tile_id_dev, int64_t n,
                                           const gm_tile* tiles_dev, int32_t n_tiles, int32_t max_class,
                                           int32_t margin_px, int32_t angle_class, double iou_merge, int64_t edge_capacity,
                                           int32_t max_per_tile,
                                           double* out_boxes_dev, int32_t* out_cls_dev, float* out_conf_dev,
                                           double* out_angle_dev, int32_t* out_src_dev, int64_t* out_count_dev,
                                           void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (n < 0 || n_tiles < 0 || max_class < 0 || !out_count_dev) return GM_EINVAL;
    cudaStream_t s = gm_stream(stream);
    if (n == 0) { GM_CUDA_TRY(cudaMemsetAsync(out_count_dev, 0, sizeof(int64_t), s)); return GM_OK; }
    if (!boxes_local_dev || !cls_dev || !conf_dev || !tile_id_dev || !tiles_dev || !workspace_dev) return GM_EINVAL;
    if (!out_boxes_dev || !out_cls_dev || !out_conf_dev || !out_angle_dev || !out_src_dev) return GM_EINVAL;
    if (n_tiles == 0 || n > (1LL << 30)) return GM_ERANGE;
    if ((long long)n_tiles * (max_class + 1) >= (1LL << 31) - 1) return GM_ERANGE;
    const long long cap = default_edge_cap(n, edge_capacity);
    if (workspace_bytes < gm_tile_postprocess_workspace_bytes(n, edge_capacity)) return GM_ENOSPC;
    MergeWs w = carve_merge(workspace_dev, n, cap, false, true);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_tile_remap<<<blocks, 256, 0, s>>>(boxes_local_dev, cls_dev, tile_id_dev, n, tiles_dev, n_tiles, max_class,
                                        margin_px, angle_class, w.gbox, w.angle, w.group, w.active); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    // per-tile counts bounded by the caller (a detector's max_det): one CTA per tile, score-sorted bit-mask sweep; GM_TILE_NMS_FAST=0
    // and every unbounded call take the engine (same results: tests/test_gpu_geom.py runs both)
    const bool bounded = max_per_tile > 0 && max_per_tile <= TN_CAP && (long long)n_tiles + 1 <= n && max_class < 32768 &&
                         gm_env_int("GM_TILE_NMS_FAST", 1) != 0;
    int st = bounded
        ? tile_nms_bounded(w.gbox, w.group, tile_id_dev, conf_dev, w.active, n, n_tiles, iou_merge, cap, w.kept_tmp,
                           reinterpret_cast<long long*>(out_count_dev), w, s)
        : nms_engine(w.gbox, w.group, (unsigned)((long long)n_tiles * (max_class + 1)), tile_id_dev,
                     (unsigned)(n_tiles - 1), conf_dev, w.active, n, iou_merge, cap, nullptr, nullptr,
                     w.kept_tmp, reinterpret_cast<long long*>(out_count_dev), w, s);
    if (st != GM_OK) return st;
    k_gather_records<<<blocks, 256, 0, s>>>(w.kept_tmp, reinterpret_cast<long long*>(out_count_dev), w.gbox, cls_dev,
                                            conf_dev, w.angle, n, out_boxes_dev, out_cls_dev, out_conf_dev,
                                            out_angle_dev, out_src_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" size_t gm_fuse_workspace_bytes(int64_t n, int64_t edge_capacity) {
    if (n < 0) return 0;
    return carve_merge(nullptr, n, default_edge_cap(n, edge_capacity), true, false).bytes;
}

extern "C" int gm_fuse_scales(const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev,
                              const int32_t* scale_id_dev, int64_t n, int32_t n_scales, int32_t max_class,
                              double iou_partner, double conf_low, double conf_high, int64_t edge_capacity,
                              int32_t* kept_idx_dev, int64_t* n_kept_dev,
                              void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (n < 0 || n_scales < 1 || max_class < 0 || !n_kept_dev) return GM_EINVAL;
    cudaStream_t s = gm_stream(stream);
    if (n == 0) { GM_CUDA_TRY(cudaMemsetAsync(n_kept_dev, 0, sizeof(int64_t), s)); return GM_OK; }
    if (!kept_idx_dev) return GM_EINVAL;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (n_scales == 1) {        // Detect_OBB.py:357-358: passthrough, no confidence filter
        k_iota<<<blocks, 256, 0, s>>>(kept_idx_dev, n, reinterpret_cast<long long*>(n_kept_dev)); gm_note_launches(1);
        GM_LAUNCH_CHECK();
        return GM_OK;
    }
    if (!boxes_dev || !cls_dev || !conf_dev || !scale_id_dev || !workspace_dev) return GM_EINVAL;
    if (n > (1LL << 30)) return GM_ERANGE;
    const long long cap = default_edge_cap(n, edge_capacity);
    if (workspace_bytes < gm_fuse_workspace_bytes(n, edge_capacity)) return GM_ENOSPC;
    MergeWs w = carve_merge(workspace_dev, n, cap, true, false);
    k_extent_init<<<1, 1, 0, s>>>(w.ext); gm_note_launches(1);
    k_fuse_init<<<blocks, 256, 0, s>>>(conf_dev, n, conf_low, w.active, w.state, w.emit, w.degree, w.cursor); gm_note_launches(1);
    k_prepare<<<blocks, 256, 0, s>>>(boxes_dev, conf_dev, nullptr, n, w.qp, w.qw, w.aabb, w.ext, w.sort.ka, w.sort.va); gm_note_launches(1);
    const unsigned int inactive_group = (unsigned int)max_class + 1u;
    k_cell_keys<<<blocks, 256, 0, s>>>(w.aabb, cls_dev, w.active, inactive_group, n, w.ext, w.sort.ka, w.sort.va); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    const int where = radix_sort_pairs(w.sort, n, 2 * CELL_BITS + bits_for(inactive_group), s);
    if (where < 0) return GM_EINVAL;
    k_discover<true><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(where ? w.sort.kb : w.sort.ka, where ? w.sort.vb : w.sort.va,
                                                                n, w.qp, w.qw, w.aabb, boxes_dev, nullptr, scale_id_dev,
                                                                inactive_group, iou_partner, w.ext, w.edges, w.edge_iou,
                                                                cap, w.degree); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    int st = exclusive_scan_u32(w.degree, w.off, n, w.sort.scan_tmp, nullptr, s);
    if (st != GM_OK) return st;
    k_adj_fill<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(w.edges, w.edge_iou, cap, w.ext, w.off, w.cursor, w.adj); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    {
        const unsigned int* off = w.off; const unsigned int* deg = w.degree; const Adj* adj = w.adj;
        Extent* x = w.ext; unsigned char* done = w.state; unsigned char* ready = w.flag8; int* emit = w.emit;
        long long nn = n;
        int rc = launch_cooperative(k_fuse_fixpoint, 512, n, s, off, deg, adj, conf_dev, nn, cap, conf_high, x, done,
                                    ready, emit);
        if (rc != GM_OK) return rc;
    }
    k_emit_flags<<<blocks, 256, 0, s>>>(w.emit, n, w.flag); gm_note_launches(1);
    st = exclusive_scan_u32(w.flag, w.pos, n, w.sort.scan_tmp, w.total, s);
    if (st != GM_OK) return st;
    k_emit_compact<<<blocks, 256, 0, s>>>(w.emit, w.pos, n, kept_idx_dev); gm_note_launches(1);
    k_finish_count<<<1, 1, 0, s>>>(w.ext, cap, w.total, reinterpret_cast<long long*>(n_kept_dev)); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

// ------------------------------------------------------------------------------------------
// cross-band exchange records

extern "C" int gm_band_pack(const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev, const double* angle_dev,
                            int64_t n_rows, const int64_t* count_dev, int64_t capacity, uint8_t* records_dev, void* stream) {
    if (capacity < 0 || n_rows < 0) return GM_EINVAL;
    if (capacity == 0) return GM_OK;
    if (!boxes_dev || !cls_dev || !conf_dev || !count_dev || !records_dev) return GM_EINVAL;
    const long long words = (long long)capacity * 10;
    k_band_pack<<<(unsigned)((words + 255) / 256), 256, 0, gm_stream(stream)>>>(boxes_dev, cls_dev, conf_dev, angle_dev, n_rows,
        reinterpret_cast<const long long*>(count_dev), capacity, records_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_band_unpack(const uint8_t* records_dev, int64_t total, int32_t world, int32_t rank, double* boxes_dev,
                              int32_t* cls_dev, int32_t* cls_owned_dev, float* conf_dev, double* angle_dev,
                              int64_t* n_valid_dev, void* stream) {
    if (total < 0 || world < 1 || rank < 0 || rank >= world || !n_valid_dev) return GM_EINVAL;
    cudaStream_t s = gm_stream(stream);
    GM_CUDA_TRY(cudaMemsetAsync(n_valid_dev, 0, sizeof(int64_t), s));
    if (total == 0) return GM_OK;
    if (!records_dev || !boxes_dev || !cls_dev || !cls_owned_dev || !conf_dev) return GM_EINVAL;
    const long long words = (long long)total * 10;
    k_band_unpack<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(records_dev, total, world, rank, boxes_dev, cls_dev, cls_owned_dev,
        conf_dev, angle_dev, reinterpret_cast<long long*>(n_valid_dev)); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_band_mask_keep(uint8_t* keep_dev, const int32_t* cls_owned_dev, int64_t total, void* stream) {
    if (total < 0) return GM_EINVAL;
    if (total == 0) return GM_OK;
    if (!keep_dev || !cls_owned_dev) return GM_EINVAL;
    k_band_mask_keep<<<(unsigned)((total + 255) / 256), 256, 0, gm_stream(stream)>>>(keep_dev, cls_owned_dev, total); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" size_t gm_band_extract_workspace_bytes(int64_t total) {
    if (total < 0) return 0;
    GmArena a(nullptr, ~(size_t)0);
    a.take<unsigned int>((size_t)total);
    a.take<unsigned int>((size_t)total);
    a.take<unsigned int>((size_t)scan_blocks(total > 0 ? total : 1));
    a.take<unsigned int>(1);
    return gm_align_up(a.off, 256);
}

extern "C" int gm_band_extract(const int32_t* order_dev, const uint8_t* keep_dev, int64_t total, const double* boxes_dev,
                               const int32_t* cls_dev, const float* conf_dev, const double* angle_dev, double* out_boxes_dev,
                               int32_t* out_cls_dev, float* out_conf_dev, double* out_angle_dev, int64_t* out_index_dev,
                               int64_t* n_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (total < 0 || !n_out_dev) return GM_EINVAL;
    cudaStream_t s = gm_stream(stream);
    if (total == 0) { GM_CUDA_TRY(cudaMemsetAsync(n_out_dev, 0, sizeof(int64_t), s)); return GM_OK; }
    if (!order_dev || !keep_dev || !boxes_dev || !cls_dev || !conf_dev || !out_boxes_dev || !out_cls_dev || !out_conf_dev ||
        !out_index_dev || !workspace_dev) return GM_EINVAL;
    if (workspace_bytes < gm_band_extract_workspace_bytes(total)) return GM_ENOSPC;
    GmArena a(workspace_dev, workspace_bytes);
    unsigned int* flag = a.take<unsigned int>((size_t)total);
    unsigned int* pos = a.take<unsigned int>((size_t)total);
    unsigned int* tmp = a.take<unsigned int>((size_t)scan_blocks(total));
    unsigned int* tot = a.take<unsigned int>(1);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    k_band_flags<<<blocks, 256, 0, s>>>(order_dev, keep_dev, total, flag); gm_note_launches(1);
    int st = exclusive_scan_u32(flag, pos, total, tmp, tot, s);
    if (st != GM_OK) return st;
    k_band_extract<<<(unsigned)((total * 8 + 255) / 256), 256, 0, s>>>(order_dev, flag, pos, total, boxes_dev, cls_dev, conf_dev, angle_dev,
        out_boxes_dev, out_cls_dev, out_conf_dev, out_angle_dev, reinterpret_cast<long long*>(out_index_dev)); gm_note_launches(1);
    k_band_count<<<1, 1, 0, s>>>(tot, reinterpret_cast<long long*>(n_out_dev)); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

// ------------------------------------------------------------------------------------------
// seam-band exchange

extern "C" size_t gm_band_merge_workspace_bytes(int64_t n_rows, int32_t world, int64_t seam_capacity, int64_t edge_capacity) {
    if (n_rows < 0 || world < 1 || seam_capacity < 0) return 0;
    const long long nu = (long long)world * seam_capacity;
    const long long nm = n_rows > nu ? n_rows : nu;
    return carve_seam(nullptr, n_rows, nu, default_edge_cap(nm, edge_capacity)).bytes;
}

extern "C" int gm_band_merge_local(double* boxes_dev, int32_t* cls_dev, float* conf_dev, int64_t n_rows,
                                   const int64_t* count_dev, int32_t max_class, double iou_thr, int64_t edge_capacity,
                                   const float* foreign_rects_host, int32_t n_rects, float extent_bound,
                                   int32_t world, int64_t seam_capacity, uint8_t* seam_records_dev,
                                   void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (n_rows < 0 || max_class < 0 || world < 1 || seam_capacity < 0 || n_rects < 0 || n_rects > 8 || !count_dev) return GM_EINVAL;
    if (!seam_records_dev || !workspace_dev || (n_rects > 0 && !foreign_rects_host)) return GM_EINVAL;
    if (n_rows > 0 && (!boxes_dev || !cls_dev || !conf_dev)) return GM_EINVAL;
    if (n_rows > (1LL << 30) || (long long)world * seam_capacity > (1LL << 30)) return GM_ERANGE;
    if (workspace_bytes < gm_band_merge_workspace_bytes(n_rows, world, seam_capacity, edge_capacity)) return GM_ENOSPC;
    cudaStream_t s = gm_stream(stream);
    const long long nu = (long long)world * seam_capacity;
    const long long n = n_rows > 0 ? n_rows : 0;
    const long long cap = default_edge_cap(n > nu ? n : nu, edge_capacity);
    SeamWs sw = carve_seam(workspace_dev, n, nu, cap);
    MergeWs w = carve_merge(sw.engine, n > 0 ? n : 1, cap, false, false);
    SeamCfg cfg{};
    cfg.n_rects = n_rects;
    cfg.bound = extent_bound;
    for (int k = 0; k < n_rects; ++k)
        cfg.rects[k] = make_float4(foreign_rects_host[4 * k], foreign_rects_host[4 * k + 1], foreign_rects_host[4 * k + 2],
                                   foreign_rects_host[4 * k + 3]);
    const long long words = ((seam_capacity + 1) > n ? (seam_capacity + 1) : n) * SEAM_WORDS;
    if (n == 0) {
        k_extent_init<<<1, 1, 0, s>>>(w.ext); gm_note_launches(1);
        GM_CUDA_TRY(cudaMemsetAsync(w.total, 0, sizeof(unsigned int), s));
    } else {
        k_band_blank<<<(unsigned)((n * 8 + 255) / 256), 256, 0, s>>>(boxes_dev, cls_dev, conf_dev, n, (const long long*)count_dev); gm_note_launches(1);
        int st = nms_resolve(boxes_dev, cls_dev, (unsigned)max_class, nullptr, 0u, conf_dev, nullptr, n, iou_thr, cap, sw.p_order,
                             &cfg, w, s);
        if (st != GM_OK) return st;
        k_state_keep_masked<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w.state, cls_dev, n, sw.p_state); gm_note_launches(1);
        k_state_flags<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w.state, cls_dev, n, 3, w.flag); gm_note_launches(1);
        st = exclusive_scan_u32(w.flag, w.pos, n, w.sort.scan_tmp, w.total, s);
        if (st != GM_OK) return st;
    }
    k_seam_pack<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(boxes_dev, cls_dev, conf_dev, w.flag, w.pos, w.total, n, seam_capacity,
        w.ext, cap, extent_bound, (const long long*)count_dev, reinterpret_cast<unsigned long long*>(seam_records_dev)); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_band_merge_finish(const uint8_t* gathered_dev, int32_t world, int32_t rank, int64_t seam_capacity,
                                    int32_t block_begin, int32_t block_count, const float* outside_rects_host, int32_t n_rects,
                                    float extent_bound,
                                    const double* boxes_dev, const int32_t* cls_dev, const float* conf_dev,
                                    const double* angle_dev, int64_t n_rows, int32_t max_class, double iou_thr,
                                    int64_t edge_capacity, double* out_boxes_dev, int32_t* out_cls_dev, float* out_conf_dev,
                                    double* out_angle_dev, int32_t* out_src_dev, int64_t* meta_dev,
                                    void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (n_rows < 0 || max_class < 0 || world < 1 || rank < 0 || rank >= world || seam_capacity < 0 || !meta_dev) return GM_EINVAL;
    if (!gathered_dev || !workspace_dev) return GM_EINVAL;
    if (block_count <= 0) { block_begin = 0; block_count = world; }                 // the seam boxes of ALL ranks
    if (block_begin < 0 || block_begin + block_count > world || rank < block_begin || rank >= block_begin + block_count) return GM_EINVAL;
    if (n_rects < 0 || n_rects > 8 || (n_rects > 0 && !outside_rects_host)) return GM_EINVAL;
    if (n_rows > 0 && (!boxes_dev || !cls_dev || !conf_dev || !out_boxes_dev || !out_cls_dev || !out_conf_dev || !out_src_dev)) return GM_EINVAL;
    if (workspace_bytes < gm_band_merge_workspace_bytes(n_rows, world, seam_capacity, edge_capacity)) return GM_ENOSPC;
    cudaStream_t s = gm_stream(stream);
    const long long nu_all = (long long)world * seam_capacity;
    const long long nu = (long long)block_count * seam_capacity;
    const long long n = n_rows;
    const long long cap = default_edge_cap(n > nu_all ? n : nu_all, edge_capacity);
    SeamWs sw = carve_seam(workspace_dev, n, nu_all, cap);
    GM_CUDA_TRY(cudaMemsetAsync(meta_dev, 0, 4 * sizeof(int64_t), s));
    k_seam_headers<<<(unsigned)((world + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(gathered_dev), world,
        seam_capacity, reinterpret_cast<long long*>(meta_dev)); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    MergeWs w = carve_merge(sw.engine, (n > nu_all ? n : nu_all) > 0 ? (n > nu_all ? n : nu_all) : 1, cap, false, false);
    if (nu > 0) {
        const long long words = nu * SEAM_WORDS;
        k_seam_unpack<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(gathered_dev), block_begin,
            block_count, seam_capacity, sw.u_boxes, sw.u_cls, sw.u_conf, sw.u_src); gm_note_launches(1);
        GM_LAUNCH_CHECK();
        // the seam boxes of the chosen ranks, in rank order = list order: the same exact greedy NMS.  When the set leaves
        // ranks out, a box that may overlap a box of an absent rank is tainted the way seam candidates are (deferral
        // propagates down its chains); if none of THIS rank's boxes ends up deferred, their verdicts equal the full set's.
        SeamCfg cfg{};
        const bool restricted = block_count < world && n_rects > 0;
        if (restricted) {
            cfg.n_rects = n_rects;
            cfg.bound = extent_bound;
            for (int k = 0; k < n_rects; ++k)
                cfg.rects[k] = make_float4(outside_rects_host[4 * k], outside_rects_host[4 * k + 1], outside_rects_host[4 * k + 2],
                                           outside_rects_host[4 * k + 3]);
        } else if (block_count < world) return GM_EINVAL;      // a restricted set without the rectangles of the absent ranks
        int st = nms_resolve(sw.u_boxes, sw.u_cls, (unsigned)max_class, nullptr, 0u, sw.u_conf, nullptr, nu, iou_thr, cap,
                             w.order_tmp, restricted ? &cfg : nullptr, w, s);
        if (st != GM_OK) return st;
        if (n > 0 && seam_capacity > 0) {
            k_seam_apply<<<(unsigned)((seam_capacity + 255) / 256), 256, 0, s>>>(w.state, sw.u_cls, sw.u_src,
                (long long)(rank - block_begin) * seam_capacity, seam_capacity, n, sw.p_state, reinterpret_cast<long long*>(meta_dev)); gm_note_launches(1);
        }
    } else {
        k_extent_init<<<1, 1, 0, s>>>(w.ext); gm_note_launches(1);
    }
    if (n > 0) {
        // kept rows of this band in the stable confidence order (state 1 after the seam verdicts)
        const unsigned blocks = (unsigned)((n + 255) / 256);
        k_keep_flags<<<blocks, 256, 0, s>>>(sw.p_order, sw.p_state, n, w.flag, nullptr); gm_note_launches(1);
        int st = exclusive_scan_u32(w.flag, w.pos, n, w.sort.scan_tmp, w.total, s);
        if (st != GM_OK) return st;
        k_compact_kept<<<blocks, 256, 0, s>>>(sw.p_order, w.flag, w.pos, n, sw.kept_tmp); gm_note_launches(1);
        k_seam_meta<<<1, 1, 0, s>>>(w.total, w.ext, cap, reinterpret_cast<long long*>(meta_dev)); gm_note_launches(1);
        k_gather_records<<<blocks, 256, 0, s>>>(sw.kept_tmp, reinterpret_cast<long long*>(meta_dev), boxes_dev, cls_dev, conf_dev,
                                                angle_dev, n, out_boxes_dev, out_cls_dev, out_conf_dev, out_angle_dev, out_src_dev); gm_note_launches(1);
    } else {
        GM_CUDA_TRY(cudaMemsetAsync(w.total, 0, sizeof(unsigned int), s));
        k_seam_meta<<<1, 1, 0, s>>>(w.total, w.ext, cap, reinterpret_cast<long long*>(meta_dev)); gm_note_launches(1);
    }
    GM_LAUNCH_CHECK();
    return GM_OK;
}
