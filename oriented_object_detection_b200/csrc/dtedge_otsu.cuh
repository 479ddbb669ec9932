// DT-Edge, Otsu binarisation (DT_BIN_METHOD = "otsu": Detect_OBB.py:109-111, Train_OBB.py:633-635):
//     acc8  = cv2.normalize(acc, None, 0, 255, NORM_MINMAX).astype(np.uint8)
//     edges = cv2.threshold(acc8, 0, 255, THRESH_BINARY + THRESH_OTSU)
// The per-tile scalar arithmetic, as __host__ __device__ functions so that tests/host_harness/otsu_host.cu can run
// the very same code against the oracle without a GPU; k_otsu_grad (dtedge.cu) adds the min/max reduction and the
// 256-bin histogram around them.  The library calls are OpenCV's (not in the reference tree): restated in
// oracle/pixel.py (normalize_minmax, otsu_threshold_u8) and pinned there against cv2 4.13 and the lifted reference.
//
// Everything float64 is written with explicit single-rounding operations: nvcc contracts a*b+c into an FMA by
// default, OpenCV's x86-64 baseline build does not.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>

namespace otsu {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float f_sqrt(float x) { return __fsqrt_rn(x); }
#else
inline double d_mul(double a, double b) { volatile double r = a * b; return r; }
inline double d_add(double a, double b) { volatile double r = a + b; return r; }
inline double d_sub(double a, double b) { volatile double r = a - b; return r; }
inline double d_div(double a, double b) { volatile double r = a / b; return r; }
inline float f_sqrt(float x) { return sqrtf(x); }
#endif

// cv2.normalize(src, dst, a, b, NORM_MINMAX) for a float source: scale = (b - a) * (1 / (smax - smin)) (0 when the
// range is <= DBL_EPSILON), shift = a - smin * scale, both float64, handed to convertTo as fp32.
__host__ __device__ inline void normalize_constants(float smin, float smax, double a, double b, float* fs, float* fh) {
    const double rng = d_sub((double)smax, (double)smin);
    const double inv = (rng > DBL_EPSILON) ? d_div(1.0, rng) : 0.0;
    const double scale = d_mul(d_sub(b, a), inv);
    const double shift = d_sub(a, d_mul((double)smin, scale));
    *fs = (float)scale;
    *fh = (float)shift;
}

// acc8 of one pixel: cv2.magnitude (correctly rounded fp32 sqrt of the integer sum), convertTo's fp32 fma, numpy's
// truncating cast.  Monotone non-decreasing in S (fs >= 0).
__host__ __device__ __forceinline__ unsigned int acc8_of(unsigned int S, float fs, float fh) {
    const float v = fmaf(f_sqrt((float)S), fs, fh);
    const int q = (int)v;                       // truncation toward zero; v is in [-tiny, 255 + tiny]
    return (unsigned int)(q < 0 ? 0 : (q > 255 ? 255 : q));
}

// OpenCV's getThreshVal_Otsu_8u: between-class variance maximised over the 256 bins, float64, first maximum wins.
__host__ __device__ inline int threshold_from_hist(const unsigned int* hist, long long n) {
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = d_add(mu, d_mul((double)i, (double)hist[i]));      // exact: integers < 2^53
    const double scale = d_div(1.0, (double)n);
    mu = d_mul(mu, scale);
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        const double p_i = d_mul((double)hist[i], scale);
        mu1 = d_mul(mu1, q1);
        q1 = d_add(q1, p_i);
        const double q2 = d_sub(1.0, q1);
        const double qlo = q1 < q2 ? q1 : q2, qhi = q1 < q2 ? q2 : q1;
        if (qlo < (double)FLT_EPSILON || qhi > d_sub(1.0, (double)FLT_EPSILON)) continue;
        mu1 = d_div(d_add(mu1, d_mul((double)i, p_i)), q1);
        const double mu2 = d_div(d_sub(mu, d_mul(q1, mu1)), q2);
        const double dm = d_sub(mu1, mu2);
        const double sigma = d_mul(d_mul(d_mul(q1, q2), dm), dm);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return max_val;
}

// edges = acc8 > thr8  <=>  S >= s_thr: the smallest S in (kmin, kmax] whose acc8 exceeds thr8, by bisection on the
// monotone map; 0xffffffff when no pixel of the tile does (flat tile, thr8 = 255).
__host__ __device__ inline unsigned int s_threshold(unsigned int kmin, unsigned int kmax, int thr8, float fs, float fh) {
    if (acc8_of(kmax, fs, fh) <= (unsigned int)thr8) return 0xffffffffu;
    if (acc8_of(kmin, fs, fh) > (unsigned int)thr8) return 0u;        // cannot happen (acc8(kmin) == 0 <= thr8); kept total
    unsigned int a = kmin, b = kmax;                                  // acc8(a) <= thr8 < acc8(b)
    while (b - a > 1u) {
        const unsigned int m = a + ((b - a) >> 1);
        if (acc8_of(m, fs, fh) > (unsigned int)thr8) b = m; else a = m;
    }
    return b;
}

}  // namespace otsu
