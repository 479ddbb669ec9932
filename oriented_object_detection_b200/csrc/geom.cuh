// Rotated-box geometry in registers (a10).
//
// Reference: compute_polygon_iou (Detect_OBB.py:144-154): shapely polygons from the four
// corners, 0.0 if either is invalid, inter / (a1 + a2 - inter).  The polygon overlay itself
// is shapely/GEOS code outside the reference tree; restated in oracle/geometry.py.
//
// Formulation.  Both quads are made counter-clockwise and expressed relative to box A's
// centroid (pair-local coordinates: mandatory in fp32).  The area of A n B
// is the boundary integral 1/2 * sum cross(P, Q) over the pieces of A's edges that lie inside
// B plus the pieces of B's edges that lie inside A; each piece comes from clipping one edge
// parametrically against the four half-planes of the other quad.  Fixed trip counts (8 edges
// x 4 planes), no vertex lists, no local memory.  Coincident edges are counted once (from A,
// and only when both quads lie on the same side of the shared line), so identical boxes give
// IoU 1 and boxes touching along an edge give 0.
#pragma once
#include <cuda_runtime.h>

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

// Sum of cross(P(t0), P(t1)) over the edges of CCW polygon P clipped to CCW convex polygon Q.
// kSubjectIsA: coincident same-direction edges are kept (A's turn); otherwise dropped (B's turn).
template <typename T, bool kSubjectIsA>
__host__ __device__ __forceinline__ T boundary_inside(const T (&px)[4], const T (&py)[4],
                                             const T (&qx)[4], const T (&qy)[4]) {
    T ex[4], ey[4], d[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int k1 = (k + 1) & 3;
        ex[k] = qx[k1] - qx[k];
        ey[k] = qy[k1] - qy[k];
#pragma unroll
        for (int i = 0; i < 4; ++i) d[k][i] = ex[k] * (py[i] - qy[k]) - ey[k] * (px[i] - qx[k]);
    }
    T acc = (T)0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        const T dxe = px[j] - px[i], dye = py[j] - py[i];
        T t0 = (T)0, t1 = (T)1;
        bool empty = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const T d0 = d[k][i], d1 = d[k][j];
            bool in0, in1;
            if (kSubjectIsA) {
                in0 = d0 >= (T)0; in1 = d1 >= (T)0;
                if (d0 == (T)0 && d1 == (T)0 && (dxe * ex[k] + dye * ey[k]) <= (T)0) empty = true;
            } else {
                in0 = d0 > (T)0; in1 = d1 > (T)0;
            }
            if (in0 != in1) {
                const T t = d0 / (d0 - d1);
                if (in0) t1 = t < t1 ? t : t1;
                else     t0 = t > t0 ? t : t0;
            } else if (!in0) {
                empty = true;
            }
        }
        if (!empty && t0 < t1) {
            const T ax = px[i] + t0 * dxe, ay = py[i] + t0 * dye;
            const T bx = px[i] + t1 * dxe, by = py[i] + t1 * dye;
            acc += ax * by - bx * ay;
        }
    }
    return acc;
}

template <typename T>
__host__ __device__ __forceinline__ T pbox_iou(const PBox<T>& A, const PBox<T>& B) {
    if (!A.valid || !B.valid) return (T)0;
    const T dx = (T)(B.cx - A.cx), dy = (T)(B.cy - A.cy);
    T bx[4], by[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bx[i] = B.lx[i] + dx; by[i] = B.ly[i] + dy; }
    T inter = (T)0.5 * (boundary_inside<T, true>(A.lx, A.ly, bx, by) +
                        boundary_inside<T, false>(bx, by, A.lx, A.ly));
    inter = inter > (T)0 ? inter : (T)0;
    const T uni = A.area + B.area - inter;
    return uni > (T)0 ? inter / uni : (T)0;
}

// IoU >= thr with the reference's float64 decision: fp32 first, and only pairs whose fp32
// value is within 1e-4 of the threshold are recomputed in float64 from the raw corners.
__host__ __device__ __forceinline__ bool iou_reaches(const PBox<float>& A, const PBox<float>& B,
                                            const double* __restrict__ rawA, const double* __restrict__ rawB,
                                            double thr, double* iou_out = nullptr) {
    const float v = pbox_iou<float>(A, B);
    double r = (double)v;
    if (fabs(r - thr) < 1e-4) {
        PBox<double> a, b;
        pbox_from_corners<double>(rawA, a);
        pbox_from_corners<double>(rawB, b);
        r = pbox_iou<double>(a, b);
    }
    if (iou_out) *iou_out = r;
    return r >= thr;
}
