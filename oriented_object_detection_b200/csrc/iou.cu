// a10: rotated IoU kernels (pair list, dense matrix, dense checksum) + FFMA peak probe.
//
// Reference: compute_polygon_iou (Detect_OBB.py:144-154).  FP32 pipe bound: nothing here is a dense contraction, so no
// tensor cores.  Dense form: the boxes are prepared once per call (k_iou_prepare: polygon record + window functionals),
// a CTA stages 256 row records in shared memory (broadcast reads) and each thread keeps one column box - the window -
// in registers and walks the rows, two per step with packed f32x2 arithmetic in the default form.
#include "gm_common.cuh"
#include "geom.cuh"

namespace {

constexpr int IOU_THREADS = 128;    // columns (boxes b) per CTA, one per thread
// Rows (boxes a) staged in shared memory per CTA: 256 by default.  (When every CTA still prepared its own boxes the row
// count also set the prologue per pair - 19 instructions at 64 rows, 9 at 256; with prepared records it only sets how often
// a thread reloads its window.)  GM_IOU_VARIANT selects the other shapes / forms for tuning.
#ifndef GM_IOU_DEFAULT_VARIANT
#define GM_IOU_DEFAULT_VARIANT 0
#endif

__global__ void __launch_bounds__(256)
k_iou_pairs(const double* __restrict__ boxes_a, const double* __restrict__ boxes_b,
            const int* __restrict__ idx_a, const int* __restrict__ idx_b, long long n_pairs,
            float* __restrict__ iou) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const long long ia = idx_a ? idx_a[p] : p;
    const long long ib = idx_b ? idx_b[p] : p;
    QPoly A, B;
    QWin Aw, Bw;
    qbox_from_corners(boxes_a + ia * 8, A, Aw);
    qbox_from_corners(boxes_b + ib * 8, B, Bw);
    qpoly_mark_concave(boxes_a + ia * 8, A);
    qpoly_mark_concave(boxes_b + ib * 8, B);
    // concave simple quads (valid for shapely, Detect_OBB.py:148-151) go through the float64 piecewise clip
    iou[p] = ((A.valid | B.valid) & 2) ? (float)iou_f64_general(boxes_a + ia * 8, boxes_b + ib * 8) : qbox_iou(A, B, Bw);
}

// Same pair list in float64: the arithmetic the reference itself performs (shapely on Python floats), concave simple
// quads included.  The decision paths (NMS, fusion, evaluation) use it for threshold-adjacent pairs; as an entry point it
// serves callers that need the reference's float64 value for a list of pairs and the error report of bench.py.
__global__ void __launch_bounds__(128)
k_iou_pairs_f64(const double* __restrict__ boxes_a, const double* __restrict__ boxes_b,
                const int* __restrict__ idx_a, const int* __restrict__ idx_b, long long n_pairs,
                double* __restrict__ iou) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const long long ia = idx_a ? idx_a[p] : p;
    const long long ib = idx_b ? idx_b[p] : p;
    iou[p] = iou_f64_from_corners(boxes_a + ia * 8, boxes_b + ib * 8);
}

// Dense n x m matrix.  Boxes are prepared ONCE by k_iou_prepare (polygon record + window functionals from the raw
// float64 corners, ~800 mostly-FP64 instructions per box) into a stream-ordered scratch buffer; the first form prepared
// its 256 rows and 128 columns again in every CTA (48x redundant on 8192 x 8192 and a serial FP64 prologue in front of
// every CTA's loop).  A thread keeps its column box as the window (in registers), the CTA's row boxes are staged in
// shared memory as 64-byte polygon records and read back as broadcast 16-byte loads.  Records of invalid boxes are
// zeroed (area 0, corners 0): the clamp of the intersection to min(area) then returns 0 without a validity test in the
// loop (qbox_iou_rect<false>).
__global__ void __launch_bounds__(128)
k_iou_prepare(const double* __restrict__ boxes_a, int n, int n_pad, QPoly* __restrict__ qa,
              const double* __restrict__ boxes_b, int m, QPoly* __restrict__ qb, QWin* __restrict__ wb,
              double* __restrict__ col_sum) {
    // one launch for both box lists: threads [0, n_pad) take the (padded) rows, the next m threads the columns; the
    // column threads also clear the checksum the matrix kernel accumulates into
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad + m) return;
    const bool is_b = i >= n_pad;
    const int k = is_b ? i - n_pad : i;
    QPoly p = QPoly{};
    QWin w = QWin{};
    if (is_b || k < n) {
        qbox_from_corners((is_b ? boxes_b : boxes_a) + (long long)k * 8, p, w);
        if (!(p.valid & 1)) {               // invalid (or concave: float64 path only): contributes IoU 0 everywhere
            p = QPoly{};
            w = QWin{};
        }
    }
    if (is_b) {
        qb[k] = p;
        wb[k] = w;
        if (col_sum) col_sum[k] = 0.0;
    } else {
        qa[k] = p;
    }
}

#ifndef GM_IOU_MINB_SCALAR
#define GM_IOU_MINB_SCALAR 1
#endif
template <bool kStore, int IOU_ROWS, int UNROLL>
__global__ void __launch_bounds__(IOU_THREADS, GM_IOU_MINB_SCALAR)
k_iou_matrix(const QPoly* __restrict__ qa, int n, const QPoly* __restrict__ qb, const QWin* __restrict__ wb, int m,
             float* __restrict__ iou, double* __restrict__ col_sum) {
    __shared__ __align__(16) QPoly rows[IOU_ROWS];
    const int j = blockIdx.x * IOU_THREADS + threadIdx.x;
    const int i0 = blockIdx.y * IOU_ROWS;
    {
        // qa is padded to a multiple of IOU_ROWS records (zero records behind n): plain 16-byte copies, no bounds test
        const uint4* src = reinterpret_cast<const uint4*>(qa + i0);
        uint4* dst = reinterpret_cast<uint4*>(rows);
#pragma unroll
        for (int k = threadIdx.x; k < IOU_ROWS * 4; k += IOU_THREADS) dst[k] = src[k];
    }
    QPoly B = QPoly{};
    QWin Bw = QWin{};
    if (j < m) { B = qb[j]; Bw = wb[j]; }
    __syncthreads();
    const int nr = min(IOU_ROWS, n - i0);
    float acc = 0.f;
    // the window type is a property of the thread's column box: the choice is made once, outside the row loop
    // (rectangles - every box of the pipeline - take the slab form; a warp with both kinds runs both loops)
    if (Bw.rect) {
#pragma unroll UNROLL
        for (int r = 0; r < nr; ++r) {
            const float v = qbox_iou_rect<false>(rows[r], B, Bw);
            if (kStore) {
                if (j < m) iou[(long long)(i0 + r) * m + j] = v;
            } else {
                acc += v;
            }
        }
    } else {
        for (int r = 0; r < nr; ++r) {
            const float v = qbox_iou_quad(rows[r], B, Bw);
            if (kStore) {
                if (j < m) iou[(long long)(i0 + r) * m + j] = v;
            } else {
                acc += v;
            }
        }
    }
    if (!kStore && j < m) atomicAdd(&col_sum[j], (double)acc);
}

// Packed form: a thread still owns one column box (the window) but takes TWO row boxes per step, every FFMA / FADD / FMUL
// of the pair issued once as FFMA2 / FADD2 / FMUL2 (geom.cuh: qbox_iou_rect2).  The CTA's rows are staged pairwise
// interleaved (QPoly2: field k = (row 2p, row 2p + 1)) so a pair arrives as broadcast 16-byte shared loads of ready-made
// register pairs.  Windows that are not parallelograms (general convex quads) take the scalar general form per row.
#ifndef GM_IOU_MINB
#define GM_IOU_MINB 1                     // minimum resident CTAs per SM asked of the register allocator (tuning)
#endif
template <bool kStore, int IOU_ROWS>
__global__ void __launch_bounds__(IOU_THREADS, GM_IOU_MINB)
k_iou_matrix2(const QPoly* __restrict__ qa, int n, const QPoly* __restrict__ qb, const QWin* __restrict__ wb, int m,
              float* __restrict__ iou, double* __restrict__ col_sum) {
    static_assert(IOU_ROWS % 2 == 0, "rows are staged in pairs");
    __shared__ __align__(16) float rows2[IOU_ROWS / 2][16][2];           // QPoly2 records, 128 B each
    const int j = blockIdx.x * IOU_THREADS + threadIdx.x;
    const int i0 = blockIdx.y * IOU_ROWS;
    // interleave the (padded, prepared) row records pairwise: word k of row r -> rows2[r / 2][k][r & 1]
    for (int t = threadIdx.x; t < IOU_ROWS * 16; t += IOU_THREADS) {
        const int r = t >> 4, k = t & 15;
        rows2[r >> 1][k][r & 1] = reinterpret_cast<const float*>(qa + i0)[t];
    }
    QPoly B = QPoly{};
    QWin Bw = QWin{};
    if (j < m) { B = qb[j]; Bw = wb[j]; }
    __syncthreads();
    const int nr = min(IOU_ROWS, n - i0);
    const int np = (nr + 1) >> 1;
    float acc = 0.f;
    if (Bw.rect) {
        QWin2 W;
        qwin2_from(B, Bw, W);
        const QPoly2* rows = reinterpret_cast<const QPoly2*>(&rows2[0][0][0]);
        float acc1 = 0.f;
        for (int p = 0; p < np; ++p) {
            float v0, v1;
            qbox_iou_rect2<false>(rows[p], W, B.valid, B.area, v0, v1);
            if (kStore) {
                if (j < m) {
                    iou[(long long)(i0 + 2 * p) * m + j] = v0;
                    if (2 * p + 1 < nr) iou[(long long)(i0 + 2 * p + 1) * m + j] = v1;
                }
            } else {
                acc += v0;
                acc1 += v1;
            }
        }
        acc += acc1;
    } else {
        for (int r = 0; r < nr; ++r) {
            const float (*src)[2] = rows2[r >> 1];
            const int h = r & 1;
            QPoly A;
            A.chx = src[0][h]; A.clx = src[1][h]; A.chy = src[2][h]; A.cly = src[3][h];
#pragma unroll
            for (int k = 0; k < 4; ++k) { A.lx[k] = src[4 + k][h]; A.ly[k] = src[8 + k][h]; }
            A.area = src[12][h];
            A.valid = __float_as_int(src[13][h]);
            const float v = qbox_iou_quad(A, B, Bw);
            if (kStore) {
                if (j < m) iou[(long long)(i0 + r) * m + j] = v;
            } else {
                acc += v;
            }
        }
    }
    if (!kStore && j < m) atomicAdd(&col_sum[j], (double)acc);
}

__global__ void k_zero_f64(double* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// FP32 peak probe: 8 independent FFMA chains per thread, 2 flops each.  The multiplier and the addend are kernel
// arguments (constant-bank operands of the FFMA: nothing to re-materialise inside the loop - the first version's
// literal constants cost one HFMA2 per trip on the FMA pipe) and the loop is unrolled 32x, so a trip is 256 FFMA
// plus 3 loop instructions: the probe can read up to 98.8 % of the pipe's peak (the first version: 32 of 36 issue
// slots = 88.9 %).  `iters` counts 8-FFMA groups and must be a multiple of 32.
__global__ void __launch_bounds__(256)
k_ffma_peak(int iters, float b, float c, float* sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    for (int i = 0; i < iters; i += 32) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    const float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456f) sink[0] = s;
}

// variant -> (rows per CTA, unroll of the row loop)
template <bool kStore>
int launch_iou_matrix(const double* a, int n, const double* b, int m, float* iou, double* col_sum, cudaStream_t s) {
    const int variant = gm_env_int("GM_IOU_VARIANT", GM_IOU_DEFAULT_VARIANT);
    const int rows = (variant == 1 || variant == 4) ? 64 : (variant == 2 ? 128 : 256);
    dim3 grid((unsigned)((m + IOU_THREADS - 1) / IOU_THREADS), (unsigned)((n + rows - 1) / rows));
    if (grid.y > 65535u) return GM_ERANGE;
    // stream-ordered scratch for the prepared records (rows padded to whole CTAs): the pool hands the same block back
    // call after call, and the free is ordered behind the matrix kernel on the same stream
    const size_t n_pad = (size_t)grid.y * (size_t)rows;
    const size_t bytes_a = gm_align_up(n_pad * sizeof(QPoly), 256);
    const size_t bytes_b = gm_align_up((size_t)m * sizeof(QPoly), 256);
    const size_t bytes_w = gm_align_up((size_t)m * sizeof(QWin), 256);
    {
        // keep what the pool has handed out across synchronisation points (the default threshold of 0 returns it to the
        // driver at every synchronize, and the next call would pay a fresh allocation); once per device
        static bool pool_ready[64] = {};
        int dev = 0;
        GM_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = 256ull << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            (void)cudaGetLastError();
            pool_ready[dev] = true;
        }
    }
    uint8_t* scratch = nullptr;
    GM_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), bytes_a + bytes_b + bytes_w, s));
    QPoly* qa = reinterpret_cast<QPoly*>(scratch);
    QPoly* qb = reinterpret_cast<QPoly*>(scratch + bytes_a);
    QWin* wb = reinterpret_cast<QWin*>(scratch + bytes_a + bytes_b);
    k_iou_prepare<<<(unsigned)((n_pad + (size_t)m + 127) / 128), 128, 0, s>>>(a, n, (int)n_pad, qa, b, m, qb, wb, kStore ? nullptr : col_sum);
    switch (variant) {
        case 1: k_iou_matrix<kStore, 64, 1><<<grid, IOU_THREADS, 0, s>>>(qa, n, qb, wb, m, iou, col_sum); break;
        case 2: k_iou_matrix<kStore, 128, 1><<<grid, IOU_THREADS, 0, s>>>(qa, n, qb, wb, m, iou, col_sum); break;
        case 3: k_iou_matrix<kStore, 256, 2><<<grid, IOU_THREADS, 0, s>>>(qa, n, qb, wb, m, iou, col_sum); break;
        case 4: k_iou_matrix<kStore, 64, 2><<<grid, IOU_THREADS, 0, s>>>(qa, n, qb, wb, m, iou, col_sum); break;
        case 5: k_iou_matrix<kStore, 256, 1><<<grid, IOU_THREADS, 0, s>>>(qa, n, qb, wb, m, iou, col_sum); break;   // scalar form
        default: k_iou_matrix2<kStore, 256><<<grid, IOU_THREADS, 0, s>>>(qa, n, qb, wb, m, iou, col_sum); break;    // packed f32x2 form
    }
    gm_note_launches(2);
    const cudaError_t launch_err = cudaGetLastError();
    GM_CUDA_TRY(cudaFreeAsync(scratch, s));
    if (launch_err != cudaSuccess) return (int)launch_err;
    return GM_OK;
}

}  // namespace

extern "C" int gm_rotated_iou_pairs(const double* boxes_a_dev, const double* boxes_b_dev,
                                    const int32_t* idx_a_dev, const int32_t* idx_b_dev, int64_t n_pairs,
                                    float* iou_dev, void* stream) {
    if (n_pairs == 0) return GM_OK;
    if (!boxes_a_dev || !boxes_b_dev || !iou_dev || n_pairs < 0) return GM_EINVAL;
    const long long blocks = (n_pairs + 255) / 256;
    if (blocks > 0x7fffffffLL) return GM_ERANGE;
    k_iou_pairs<<<(unsigned)blocks, 256, 0, gm_stream(stream)>>>(boxes_a_dev, boxes_b_dev, idx_a_dev, idx_b_dev,
                                                               n_pairs, iou_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_rotated_iou_pairs_f64(const double* boxes_a_dev, const double* boxes_b_dev,
                                        const int32_t* idx_a_dev, const int32_t* idx_b_dev, int64_t n_pairs,
                                        double* iou_dev, void* stream) {
    if (n_pairs == 0) return GM_OK;
    if (!boxes_a_dev || !boxes_b_dev || !iou_dev || n_pairs < 0 || ((idx_a_dev == nullptr) != (idx_b_dev == nullptr))) return GM_EINVAL;
    const long long blocks = (n_pairs + 127) / 128;
    if (blocks > 0x7fffffffLL) return GM_ERANGE;
    k_iou_pairs_f64<<<(unsigned)blocks, 128, 0, gm_stream(stream)>>>(boxes_a_dev, boxes_b_dev, idx_a_dev, idx_b_dev,
                                                                     n_pairs, iou_dev); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_rotated_iou_matrix(const double* boxes_a_dev, int32_t n, const double* boxes_b_dev, int32_t m,
                                     float* iou_dev, void* stream) {
    if (n == 0 || m == 0) return GM_OK;
    if (!boxes_a_dev || !boxes_b_dev || !iou_dev || n < 0 || m < 0) return GM_EINVAL;
    const int st = launch_iou_matrix<true>(boxes_a_dev, n, boxes_b_dev, m, iou_dev, nullptr, gm_stream(stream));
    if (st != GM_OK) return st;
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_rotated_iou_matrix_sum(const double* boxes_a_dev, int32_t n, const double* boxes_b_dev, int32_t m,
                                         double* col_sum_dev, void* stream) {
    if (m == 0) return GM_OK;
    if (!boxes_a_dev || !boxes_b_dev || !col_sum_dev || n < 0 || m < 0) return GM_EINVAL;
    if (n == 0) {
        k_zero_f64<<<(m + 255) / 256, 256, 0, gm_stream(stream)>>>(col_sum_dev, m); gm_note_launches(1);
        GM_LAUNCH_CHECK();
        return GM_OK;
    }
    const int st = launch_iou_matrix<false>(boxes_a_dev, n, boxes_b_dev, m, nullptr, col_sum_dev, gm_stream(stream));
    if (st != GM_OK) return st;
    GM_LAUNCH_CHECK();
    return GM_OK;
}

extern "C" int gm_ffma_peak(int32_t iters, double* tflops_host, void* stream) {
    if (!tflops_host || iters <= 0) return GM_EINVAL;
    iters = (iters + 31) / 32 * 32;
    cudaStream_t s = gm_stream(stream);
    float* sink = nullptr;
    GM_CUDA_TRY(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    GM_CUDA_TRY(cudaEventCreate(&e0));
    GM_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = GM_NUM_SMS_B200 * 8;
    k_ffma_peak<<<blocks, 256, 0, s>>>(iters, 1.0000001f, 1e-7f, sink); gm_note_launches(1);           // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, s);
        k_ffma_peak<<<blocks, 256, 0, s>>>(iters, 1.0000001f, 1e-7f, sink); gm_note_launches(1);
        cudaEventRecord(e1, s);
        GM_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    GM_LAUNCH_CHECK();
    const double flops = (double)blocks * 256.0 * (double)iters * 8.0 * 2.0;
    *tflops_host = flops / ((double)best * 1e-3) / 1e12;
    return GM_OK;
}
