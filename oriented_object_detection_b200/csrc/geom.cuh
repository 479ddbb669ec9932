// Rotated-box geometry (a10).
//
// Reference: compute_polygon_iou (Detect_OBB.py:144-154): shapely polygons from the four
// corners, 0.0 if either is invalid, inter / (a1 + a2 - inter).  The polygon overlay itself
// is shapely/GEOS code outside the reference tree; restated in oracle/geometry.py.
//
// Formulation.  Both quads are made counter-clockwise and expressed relative to box A's
// centroid (pair-local coordinates: mandatory in fp32).  A is clipped against
// the four half-planes of B (Sutherland-Hodgman, inclusive side test, same order of operations
// as the oracle) and the area of the resulting polygon (<= 8 vertices) is its shoelace sum.
// The growing vertex list is the only dynamically indexed state; it lives in a caller-provided
// scratch column (shared memory, one column per thread, stride = threads per CTA, so every
// access is bank-conflict free) and never in local memory.  A clip always yields a closed
// polygon, so near-coincident edges can only cost a sliver of area, never a topology error.
#pragma once
#include <cuda_runtime.h>

#define GEOM_SCRATCH_WORDS 32   // per thread: 2 buffers x 8 vertices x (x, y)

// A box prepared once: centroid in map coordinates (always float64: the reference's corner
// coordinates are Python floats, and fp32(local + tile offset) would already move a 16384-px
// map coordinate by 5e-4 px), CCW corners relative to the centroid in T.
template <typename T>
struct PBox {
    double cx, cy;
    T lx[4], ly[4];
    T area;
    int valid;      // convex, non-zero area
};

template <typename T>
__host__ __device__ __forceinline__ void pbox_from_corners(const double* __restrict__ b, PBox<T>& p) {
    T x[4], y[4];
    p.cx = 0.25 * ((b[0] + b[2]) + (b[4] + b[6]));
    p.cy = 0.25 * ((b[1] + b[3]) + (b[5] + b[7]));
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = (T)(b[2 * i] - p.cx); y[i] = (T)(b[2 * i + 1] - p.cy); }
    T s = (T)0;
    bool pos = false, neg = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3, k = (i + 2) & 3;
        s += x[i] * y[j] - x[j] * y[i];
        const T cr = (x[j] - x[i]) * (y[k] - y[j]) - (y[j] - y[i]) * (x[k] - x[j]);
        pos |= cr > (T)0;
        neg |= cr < (T)0;
    }
    p.valid = (s != (T)0) && !(pos && neg);
    const bool flip = s < (T)0;
    p.lx[0] = x[0]; p.ly[0] = y[0];
    p.lx[1] = flip ? x[3] : x[1]; p.ly[1] = flip ? y[3] : y[1];
    p.lx[2] = x[2]; p.ly[2] = y[2];
    p.lx[3] = flip ? x[1] : x[3]; p.ly[3] = flip ? y[1] : y[3];
    p.area = (T)0.5 * (flip ? -s : s);
}

// Area of (CCW quad A) n (CCW convex quad B), both in the same local frame.
// scratch: GEOM_SCRATCH_WORDS values of T, element e at scratch[e * stride].
template <typename T>
__host__ __device__ __forceinline__ T clip_area(const T (&ax)[4], const T (&ay)[4],
                                                const T (&bx)[4], const T (&by)[4],
                                                T* __restrict__ scratch, int stride) {
#define GEOM_AT(buf, slot, c) scratch[(((buf) * 8 + (slot)) * 2 + (c)) * stride]
#pragma unroll
    for (int i = 0; i < 4; ++i) { GEOM_AT(0, i, 0) = ax[i]; GEOM_AT(0, i, 1) = ay[i]; }
    int n = 4, cur = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int k1 = (k + 1) & 3;
        const T qx = bx[k], qy = by[k];
        const T ex = bx[k1] - qx, ey = by[k1] - qy;
        int m = 0;
        if (n > 0) {
            T px = GEOM_AT(cur, 0, 0), py = GEOM_AT(cur, 0, 1);
            const T fx = px, fy = py;
            T dp = ex * (py - qy) - ey * (px - qx);
            const T d_first = dp;
            for (int i = 0; i < n; ++i) {
                T nx, ny, dq;
                if (i + 1 < n) {
                    nx = GEOM_AT(cur, i + 1, 0); ny = GEOM_AT(cur, i + 1, 1);
                    dq = ex * (ny - qy) - ey * (nx - qx);
                } else {
                    nx = fx; ny = fy; dq = d_first;
                }
                if (dp >= (T)0) {
                    GEOM_AT(cur ^ 1, m, 0) = px; GEOM_AT(cur ^ 1, m, 1) = py; ++m;
                    if (dq < (T)0) {
                        const T t = dp / (dp - dq);
                        GEOM_AT(cur ^ 1, m, 0) = px + t * (nx - px); GEOM_AT(cur ^ 1, m, 1) = py + t * (ny - py); ++m;
                    }
                } else if (dq >= (T)0) {
                    const T t = dp / (dp - dq);
                    GEOM_AT(cur ^ 1, m, 0) = px + t * (nx - px); GEOM_AT(cur ^ 1, m, 1) = py + t * (ny - py); ++m;
                }
                px = nx; py = ny; dp = dq;
            }
        }
        n = m; cur ^= 1;
    }
    T s = (T)0;
    if (n >= 3) {
        const T fx = GEOM_AT(cur, 0, 0), fy = GEOM_AT(cur, 0, 1);
        T px = fx, py = fy;
        for (int i = 1; i <= n; ++i) {
            const T nx = (i < n) ? GEOM_AT(cur, i, 0) : fx;
            const T ny = (i < n) ? GEOM_AT(cur, i, 1) : fy;
            s += px * ny - nx * py;
            px = nx; py = ny;
        }
    }
#undef GEOM_AT
    s = s < (T)0 ? -s : s;
    return (T)0.5 * s;
}

template <typename T>
__host__ __device__ __forceinline__ T pbox_iou(const PBox<T>& A, const PBox<T>& B, T* __restrict__ scratch, int stride) {
    if (!A.valid || !B.valid) return (T)0;
    const T dx = (T)(B.cx - A.cx), dy = (T)(B.cy - A.cy);
    T bx[4], by[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bx[i] = B.lx[i] + dx; by[i] = B.ly[i] + dy; }
    const T inter = clip_area<T>(A.lx, A.ly, bx, by, scratch, stride);
    const T uni = A.area + B.area - inter;
    return uni > (T)0 ? inter / uni : (T)0;
}

// float64 IoU from raw corners (rare path: thread-private scratch in local memory).
__host__ __device__ inline double iou_f64_from_corners(const double* __restrict__ rawA, const double* __restrict__ rawB) {
    PBox<double> a, b;
    pbox_from_corners<double>(rawA, a);
    pbox_from_corners<double>(rawB, b);
    double scratch[GEOM_SCRATCH_WORDS];
    return pbox_iou<double>(a, b, scratch, 1);
}

// IoU >= thr with the reference's float64 decision: fp32 first, and only pairs whose fp32
// value is within 1e-4 of the threshold are recomputed in float64 from the raw corners.
__host__ __device__ __forceinline__ bool iou_reaches(const PBox<float>& A, const PBox<float>& B,
                                                     const double* __restrict__ rawA, const double* __restrict__ rawB,
                                                     double thr, float* __restrict__ scratch, int stride,
                                                     double* iou_out = nullptr) {
    const float v = pbox_iou<float>(A, B, scratch, stride);
    double r = (double)v;
    if (fabs(r - thr) < 1e-4) r = iou_f64_from_corners(rawA, rawB);
    if (iou_out) *iou_out = r;
    return r >= thr;
}
