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
#include <cstring>

#define GEOM_SCRATCH_WORDS 32   // per thread: 2 buffers x 8 vertices x (x, y)
#ifndef GEOM_RECT_EPS
#define GEOM_RECT_EPS 5e-6       // |alpha - 1|, |beta - 1| below which a window takes the slab (parallelogram) form
#endif

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

// kSnap: the reference point is the centroid ROUNDED TO fp32 (any point near the box serves: only the corners relative
// to it and differences of reference points enter the arithmetic).  The difference of two fp32 reference points is a
// single correctly rounded fp32 subtraction, so the window forms need no float-float low parts.
template <typename T, bool kSnap = false>
__host__ __device__ __forceinline__ void pbox_from_corners(const double* __restrict__ b, PBox<T>& p) {
    T x[4], y[4];
    p.cx = 0.25 * ((b[0] + b[2]) + (b[4] + b[6]));
    p.cy = 0.25 * ((b[1] + b[3]) + (b[5] + b[7]));
    if (kSnap) { p.cx = (double)(float)p.cx; p.cy = (double)(float)p.cy; }
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

// ------------------------------------------------------------------------------------------
// Branch-free fp32 IoU (the throughput path).
//
// The vertex list of a Sutherland-Hodgman clip forces dynamic indexing and divergent loops.  The
// area of A n B is instead taken as the boundary integral (Green) over the pieces of A's edges that
// lie inside B plus the pieces of B's edges that lie inside A:
//     area = 1/2 [ sum_i w_i cross(p_i, d_i)  +  sum_k len_k cross(b_k, e_k) ].
// That formulation is only robust if every inside/outside decision about one point and one line is
// taken once, from one number.  So box B (the "window") is prepared as six affine functionals of a
// point (x, y): its canonical coordinates X, Y in the frame that maps b0 -> (0,0), b1 -> (1,0),
// b3 -> (0,1) (then the half-planes of edges 0 and 3 are Y >= 0 and X >= 0, and those two edges pass
// through the origin, so their boundary terms vanish), the half-plane functions H1, H2 of the other
// two edges, and the normalised positions U1, U2 along those edges.  The four vertices of A are pushed
// through the functionals once (24 FMA); each edge of A is then cut by the four half-planes from the
// per-vertex values (one reciprocal each), and a window edge's piece inside A is the span between the
// two places where A's boundary changes sides of that line - read from the same per-vertex values and
// the same cut parameters.  A vertex exactly on a line is "inside" for both uses, an edge of A lying on
// a window line is counted once (as a piece of A), so coincident edges cannot be double counted or
// lost.  ~330 instructions per pair, no branches, no local or shared memory.  IoU is affine invariant,
// so the canonical-frame area is scaled back by |u x v| only at the end.

struct QPoly {                      // a box as the polygon being cut: 64 bytes
    float chx, clx, chy, cly;       // reference point (map coordinates): the centroid rounded to fp32 (chx, chy); clx = cly = 0
                                    // (kept for the record layout: the first forms carried a float-float centroid)
    float lx[4], ly[4];             // CCW corners relative to the reference point
    float area;
    int valid;                      // 1: convex, non-zero area; 2: concave simple quad (IoU only through iou_f64_general); 0: invalid
    float pad[2];
};

struct QWin {                       // the same box as the window: 96 bytes
    float f[6][3];                  // X, Y, H1, H2, U1, U2 : f(x, y) = a x + b y + c, (x, y) relative to the centroid
    float alpha, beta;              // canonical image of corner 2; also cross(b_k, e_k) of edges 2 and 1
    float scale;                    // |u x v|: canonical area -> map area
    int rect;                       // the window is a parallelogram to GEOM_RECT_EPS (alpha ~ beta ~ 1): qbox_iou_rect applies
    float ea, eb;                   // alpha - 1, beta - 1 (rounded from float64: the deviation itself keeps full precision)
};

static_assert(sizeof(QPoly) == 64 && sizeof(QWin) == 96, "record sizes are part of the staging / exchange layouts");

__host__ __device__ inline void qbox_from_corners(const double* __restrict__ b, QPoly& P, QWin& Wn) {
    PBox<float> pb;
    pbox_from_corners<float, true>(b, pb);       // orientation, validity and area as in the clip path; fp32 reference point
    P.chx = (float)pb.cx; P.clx = 0.f;           // exact: pb.cx is an fp32 value
    P.chy = (float)pb.cy; P.cly = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { P.lx[i] = pb.lx[i]; P.ly[i] = pb.ly[i]; }
    P.area = pb.area;
    P.valid = pb.valid;
    P.pad[0] = P.pad[1] = 0.f;
    // window frame at the corner with the largest |edge x edge| (a valid quad may have one straight corner)
    double x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = (double)pb.lx[i]; y[i] = (double)pb.ly[i]; }
    int k0 = 0;
    double best = -1.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int n = (k + 1) & 3, p = (k + 3) & 3;
        const double cr = (x[n] - x[k]) * (y[p] - y[k]) - (y[n] - y[k]) * (x[p] - x[k]);
        if (cr > best) { best = cr; k0 = k; }
    }
    double qx[4], qy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // (k0 + i) & 3 without dynamic indexing of x[], y[]
        const int s = (k0 + i) & 3;
        qx[i] = s == 0 ? x[0] : s == 1 ? x[1] : s == 2 ? x[2] : x[3];
        qy[i] = s == 0 ? y[0] : s == 1 ? y[1] : s == 2 ? y[2] : y[3];
    }
    const double ux = qx[1] - qx[0], uy = qy[1] - qy[0], vx = qx[3] - qx[0], vy = qy[3] - qy[0];
    const double det = ux * vy - uy * vx;
    double F[6][3];
    double alpha = 1.0, beta = 1.0;
    if (!(det > 0.0) || !pb.valid) {
        P.valid = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k][0] = F[k][1] = F[k][2] = 0.0;
    } else {
        const double inv = 1.0 / det;
        F[0][0] = vy * inv;  F[0][1] = -vx * inv; F[0][2] = -(F[0][0] * qx[0] + F[0][1] * qy[0]);     // X
        F[1][0] = -uy * inv; F[1][1] = ux * inv;  F[1][2] = -(F[1][0] * qx[0] + F[1][1] * qy[0]);     // Y
        alpha = F[0][0] * qx[2] + F[0][1] * qy[2] + F[0][2];
        beta = F[1][0] * qx[2] + F[1][1] * qy[2] + F[1][2];
        // H1 = -beta X + (alpha - 1) Y + beta ;  H2 = -(1 - beta) X - alpha Y + alpha
        const double h1x = -beta, h1y = alpha - 1.0, h1c = beta;
        const double h2x = -(1.0 - beta), h2y = -alpha, h2c = alpha;
        // U1 = ((X - 1) e1x + Y e1y) / |e1|^2, e1 = (alpha - 1, beta);  U2 = ((X - alpha) e2x + (Y - beta) e2y) / |e2|^2, e2 = (-alpha, 1 - beta)
        const double e1x = alpha - 1.0, e1y = beta, n1 = 1.0 / (e1x * e1x + e1y * e1y);
        const double e2x = -alpha, e2y = 1.0 - beta, n2 = 1.0 / (e2x * e2x + e2y * e2y);
        const double u1x = e1x * n1, u1y = e1y * n1, u1c = -e1x * n1;
        const double u2x = e2x * n2, u2y = e2y * n2, u2c = -(alpha * e2x + beta * e2y) * n2;
        const double gx[4] = {h1x, h2x, u1x, u2x}, gy[4] = {h1y, h2y, u1y, u2y}, gc[4] = {h1c, h2c, u1c, u2c};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            F[2 + k][0] = gx[k] * F[0][0] + gy[k] * F[1][0];
            F[2 + k][1] = gx[k] * F[0][1] + gy[k] * F[1][1];
            F[2 + k][2] = gx[k] * F[0][2] + gy[k] * F[1][2] + gc[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) { Wn.f[k][0] = (float)F[k][0]; Wn.f[k][1] = (float)F[k][1]; Wn.f[k][2] = (float)F[k][2]; }
    Wn.alpha = (float)alpha; Wn.beta = (float)beta;
    Wn.scale = (float)(det > 0.0 ? det : 0.0);
    // Every box the detector emits is a rectangle; its canonical image of corner 2 is (1, 1) up to the rounding of
    // the corners (fp32 tile-local corners + integer tile offset, the reference's tuples: |alpha - 1| <= 4e-6 over
    // 100 k boxes of 11..100 px; fp32 GLOBAL corners at 2000 px: 95 % below 5e-6).
    Wn.rect = (det > 0.0 && pb.valid && fabs(alpha - 1.0) < GEOM_RECT_EPS && fabs(beta - 1.0) < GEOM_RECT_EPS) ? 1 : 0;
    Wn.ea = (float)(alpha - 1.0);
    Wn.eb = (float)(beta - 1.0);
}

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float q_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
inline float q_rcp(float x) { return 1.0f / x; }
#endif

// Cuts the parameter interval [t0, t1] of an edge whose end points have half-plane values hp, hq
// (inside = h >= 0).  Returns the parameter where the edge meets the line; `up` = h increases along the edge.
__host__ __device__ __forceinline__ float q_cut(float hp, float hq, float& t0, float& t1, bool& up) {
    // The slope gets a 1e-10 relative share of hp: nothing where the cut matters (|hp| <= |hq - hp| there),
    // but a parallel edge (hq == hp) is then unrestricted when inside (hp > 0: tc = -1e10 as a lower bound),
    // empty when outside (hp < 0: tc = -1e10 as an upper bound), and an edge lying on the line (hp == hq == 0)
    // yields NaN, which fmaxf / fminf ignore: inside, inclusive.  Two FMA-pipe instructions, no bit tricks.
    const float dh = fmaf(hp, 1e-10f, hq - hp);
    const float tc = -hp * q_rcp(dh);
    up = dh > 0.f;
    t0 = fmaxf(t0, up ? tc : -3e38f);
    t1 = fminf(t1, up ? 3e38f : tc);
    return tc;
}

// Parallelogram window (rotated rectangles: every box of the pipeline).  In the canonical frame the window is the
// unit square, so the four half-planes are the two SLABS 0 <= X <= 1 and 0 <= Y <= 1: one reciprocal per slab gives
// both cut parameters, the two remaining functionals are H1 = 1 - X, H2 = 1 - Y, and the positions along window edges
// 1 and 2 are U1 = Y, U2 = 1 - X - only X and Y of A's vertices are needed (16 FMA instead of 48).
// Degenerate edges by IEEE arithmetic instead of a perturbed slope: dX == 0 gives r = +-inf, the cut parameters become
// +-inf (strictly inside / outside the slab) or NaN (the edge lies ON a slab line), `up` follows the sign of r, and
// fmaxf / fminf drop the NaN - inside, inclusive, exactly the convention of q_cut.
//
// Pieces of window edges 1 (X = 1) and 2 (Y = 1) inside A, as a SIGNED SUM instead of selected span ends: A's boundary
// leaves the half-plane X > 1 once and enters it once (A is convex and CCW: it leaves at the upper end of the span and
// enters at the lower one), so span = sum over A's edges of s * sat(u), with s = [X_i > 1] - [X_j > 1] in {-1, 0, +1}
// and u the position of the edge's crossing along the window edge, saturated to [0, 1].  The saturation is the FFMA's
// own .SAT modifier, which also turns the NaN / inf of an edge parallel to the line (s == 0 there) into 0; the per-vertex
// indicators are four FSETs.  This moves ~30 selects / predicate operations per pair from the 16-lane ALU pipe (the
// pipe that limited the previous form) to the FMA pipe and drops the span clamps of the tail.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ float q_sat(float x) { return __saturatef(x); }
#else
inline float q_sat(float x) { return fminf(fmaxf(x, 0.f), 1.f); }          // NaN -> 0, like the .SAT modifier
#endif

// kChecked = false: the caller guarantees that an invalid record has area 0 (gm_iou_prepare writes such records), so
// the clamp to min(area) already yields 0 and the validity test drops out of the loop.
template <bool kChecked = true>
__host__ __device__ __forceinline__ float qbox_iou_rect(const QPoly& A, const QPoly& Bp, const QWin& Bw) {
    const float dx = A.chx - Bp.chx;            // reference points are fp32 values: one correctly rounded subtraction
    const float dy = A.chy - Bp.chy;
    float X[4], Y[4], OX[4], OY[4], GX[4], GY[4];
    {
        const float cx = fmaf(Bw.f[0][0], dx, fmaf(Bw.f[0][1], dy, Bw.f[0][2]));
        const float cy = fmaf(Bw.f[1][0], dx, fmaf(Bw.f[1][1], dy, Bw.f[1][2]));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            X[i] = fmaf(Bw.f[0][0], A.lx[i], fmaf(Bw.f[0][1], A.ly[i], cx));
            Y[i] = fmaf(Bw.f[1][0], A.lx[i], fmaf(Bw.f[1][1], A.ly[i], cy));
            OX[i] = 1.f - X[i];
            OY[i] = 1.f - Y[i];
            GX[i] = X[i] >= 1.f ? 1.f : 0.f;                      // on the line counts as beyond it (see below)
            GY[i] = Y[i] >= 1.f ? 1.f : 0.f;
        }
    }
    // The true window has corner 2 at (1 + ea, 1 + eb): against the unit square it gains (loses) a sliver along edge 1 of
    // thickness ea * U1 and one along edge 2 of thickness eb * (1 - U2).  Where those edges run inside A the slivers are
    // inside A too, to first order in ea, eb (<= GEOM_RECT_EPS): their areas over the inside spans [lo, hi] are
    // ea (hi1^2 - lo1^2) / 2 and eb (len2 - (hi2^2 - lo2^2) / 2), i.e. every signed crossing term s*u gets the weight
    // (1 + ea u) on edge 1 and (1 + 2 eb - eb u) on edge 2.  Measured against float64 (tests/test_geom_host.py): same
    // error as the general form (<= 7e-7) on overlapping, contained, touching and identical boxes; up to
    // 0.7 * GEOM_RECT_EPS = 3.5e-6 only when an edge of A runs INSIDE a sliver, i.e. within 5e-6 of a side of an edge of
    // B over its length (jittered copies of the same box, IoU ~ 1: far from any threshold).
    //
    // Near / far cut of a slab = min / max of the parameters on its two lines (two FMNMX instead of a compare and two
    // selects).  IEEE does the degenerate edges: ex == 0 gives rx = inf and the parameters -X*inf, (1-X)*inf are
    // (+inf, +inf) left of the slab, (-inf, +inf) inside, (-inf, -inf) right of it - empty, unrestricted, empty - and a
    // NaN exactly ON a line (0 * inf), which min / max drop: an edge of A lying on X = 0 gets (inf, inf), one on X = 1
    // gets (-inf, -inf); both are EMPTY as pieces of A.  That is consistent: the boundary term of a piece on X = 0 (or
    // Y = 0) vanishes anyway (those window edges pass through the canonical origin), and a piece on X = 1 is picked up
    // by the window-edge span instead, because a vertex ON the line counts as beyond it (G = [X >= 1]): A's boundary then
    // leaves the half-plane at one end of the coincident edge and enters at the other, and the span between them is the
    // edge.  Coincident edges are counted exactly once either way.
    const float k2a = fmaf(2.f, Bw.eb, 1.f), k2b = -Bw.eb;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        const float ex = X[j] - X[i], ey = Y[j] - Y[i];
        const float rx = q_rcp(ex), ry = q_rcp(ey);
        const float ax = -X[i] * rx, bx = OX[i] * rx;                 // edge parameter on the lines X = 0 and X = 1
        const float ay = -Y[i] * ry, by = OY[i] * ry;
        const float t0 = fmaxf(fmaxf(0.f, fminf(ax, bx)), fminf(ay, by));
        const float t1 = fminf(fminf(1.f, fmaxf(ax, bx)), fmaxf(ay, by));
        const float w = q_sat(t1 - t0);                               // t0 >= 0 and t1 <= 1: sat == max(., 0); NaN (inf - inf) -> 0
        acc = fmaf(w, X[i] * ey - Y[i] * ex, acc);
        const float u1 = q_sat(fmaf(bx, ey, Y[i]));                   // U1 = Y where the edge meets X = 1
        const float u2 = q_sat(fmaf(-by, ex, OX[i]));                 // U2 = 1 - X where the edge meets Y = 1
        const float s1 = (GX[i] - GX[j]) * u1;
        const float s2 = (GY[i] - GY[j]) * u2;
        acc = fmaf(s1, fmaf(Bw.ea, u1, 1.f), acc);
        acc = fmaf(s2, fmaf(k2b, u2, k2a), acc);
    }
    float inter = 0.5f * Bw.scale * acc;
    inter = fminf(fmaxf(inter, 0.f), fminf(A.area, Bp.area));
    const float uni = A.area + Bp.area - inter;
    const bool ok = (!kChecked || (A.valid & Bp.valid & 1)) && uni > 0.f;   // valid == 2: concave simple quad, float64 path only
    return ok ? inter * q_rcp(uni) : 0.f;
}

// ------------------------------------------------------------------------------------------
// Two polygons per call with Blackwell's packed fp32 arithmetic (fma/add/sub/mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2:
// one issue slot for the same operation on two independent pairs).  qbox_iou_rect is issue bound - ~182 slots per pair,
// ~119 of them FFMA / FADD / FMUL - so two row boxes share every one of those slots; what stays scalar is what has no
// packed form: the reciprocals (MUFU), the side selects and min / max of the cuts (ALU pipe), the indicator FSETs and
// the three saturating operations per edge (.sat does not exist for f32x2).  Per lane the operations and their order
// are those of qbox_iou_rect except that -X and -Y are formed once per vertex instead of as operand negations, so the
// two forms agree to rounding (tests/test_geom_host.py runs this one on the host through the struct emulation below).
#ifdef __CUDA_ARCH__
typedef unsigned long long qf2;
__device__ __forceinline__ qf2 q2_pack(float lo, float hi) { qf2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float q2_lo(qf2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float q2_hi(qf2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
// GM_Q2_SCALAR_{FMA,ADD,MUL}: issue that operation as two scalar instructions on the halves instead (tuning: the packed
// forms save issue slots but occupy the FP32 pipe per LANE-operation - measured in DESIGN.md section 4.3).
#ifdef GM_Q2_SCALAR_FMA
__device__ __forceinline__ qf2 q2_fma(qf2 a, qf2 b, qf2 c) { return q2_pack(fmaf(q2_lo(a), q2_lo(b), q2_lo(c)), fmaf(q2_hi(a), q2_hi(b), q2_hi(c))); }
#else
__device__ __forceinline__ qf2 q2_fma(qf2 a, qf2 b, qf2 c) { qf2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
#endif
#ifdef GM_Q2_SCALAR_ADD
__device__ __forceinline__ qf2 q2_add(qf2 a, qf2 b) { return q2_pack(q2_lo(a) + q2_lo(b), q2_hi(a) + q2_hi(b)); }
__device__ __forceinline__ qf2 q2_sub(qf2 a, qf2 b) { return q2_pack(q2_lo(a) - q2_lo(b), q2_hi(a) - q2_hi(b)); }
#else
__device__ __forceinline__ qf2 q2_add(qf2 a, qf2 b) { qf2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ qf2 q2_sub(qf2 a, qf2 b) { qf2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#endif
#ifdef GM_Q2_SCALAR_MUL
__device__ __forceinline__ qf2 q2_mul(qf2 a, qf2 b) { return q2_pack(q2_lo(a) * q2_lo(b), q2_hi(a) * q2_hi(b)); }
#else
__device__ __forceinline__ qf2 q2_mul(qf2 a, qf2 b) { qf2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#endif
#else
struct qf2 { float lo, hi; };
inline qf2 q2_pack(float lo, float hi) { qf2 r; r.lo = lo; r.hi = hi; return r; }
inline float q2_lo(qf2 v) { return v.lo; }
inline float q2_hi(qf2 v) { return v.hi; }
inline qf2 q2_fma(qf2 a, qf2 b, qf2 c) { return q2_pack(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
inline qf2 q2_add(qf2 a, qf2 b) { return q2_pack(a.lo + b.lo, a.hi + b.hi); }
inline qf2 q2_sub(qf2 a, qf2 b) { return q2_pack(a.lo - b.lo, a.hi - b.hi); }
inline qf2 q2_mul(qf2 a, qf2 b) { return q2_pack(a.lo * b.lo, a.hi * b.hi); }
#endif
__host__ __device__ __forceinline__ qf2 q2_dup(float v) { return q2_pack(v, v); }

// Two polygon records side by side: every field holds (polygon 0, polygon 1).  128 bytes.
struct QPoly2 {
    qf2 chx, clx, chy, cly;
    qf2 lx[4], ly[4];
    qf2 area;
    qf2 valid;                      // the int `valid` of each record, bit-cast
    qf2 pad[2];
};

// The window's constants, duplicated into both lanes once per window (outside the pair loop).
struct QWin2 {
    qf2 f00, f01, f02, f10, f11, f12;
    qf2 bhx, bhy;                   // window reference point
    qf2 ea, k2a, k2b, one, hscale, barea;
};

__host__ __device__ __forceinline__ void qwin2_from(const QPoly& Bp, const QWin& Bw, QWin2& W) {
    W.f00 = q2_dup(Bw.f[0][0]); W.f01 = q2_dup(Bw.f[0][1]); W.f02 = q2_dup(Bw.f[0][2]);
    W.f10 = q2_dup(Bw.f[1][0]); W.f11 = q2_dup(Bw.f[1][1]); W.f12 = q2_dup(Bw.f[1][2]);
    W.bhx = q2_dup(Bp.chx); W.bhy = q2_dup(Bp.chy);
    W.ea = q2_dup(Bw.ea); W.k2a = q2_dup(fmaf(2.f, Bw.eb, 1.f)); W.k2b = q2_dup(-Bw.eb);
    W.one = q2_dup(1.f); W.hscale = q2_dup(0.5f * Bw.scale); W.barea = q2_dup(Bp.area);
}

// IoU of the two polygons of A against the parallelogram window B; results in (out0, out1).
template <bool kChecked = true>
__host__ __device__ __forceinline__ void qbox_iou_rect2(const QPoly2& A, const QWin2& W, int b_valid, float b_area,
                                                        float& out0, float& out1) {
    const qf2 dx = q2_sub(A.chx, W.bhx);                 // fp32 reference points: exact to one rounding (qbox_iou_rect)
    const qf2 dy = q2_sub(A.chy, W.bhy);
    const qf2 cx = q2_fma(W.f00, dx, q2_fma(W.f01, dy, W.f02));
    const qf2 cy = q2_fma(W.f10, dx, q2_fma(W.f11, dy, W.f12));
    const qf2 zero = q2_dup(0.f);
    qf2 X[4], Y[4], NX[4], NY[4], OX[4], OY[4], GX[4], GY[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        X[i] = q2_fma(W.f00, A.lx[i], q2_fma(W.f01, A.ly[i], cx));
        Y[i] = q2_fma(W.f10, A.lx[i], q2_fma(W.f11, A.ly[i], cy));
        NX[i] = q2_sub(zero, X[i]);
        NY[i] = q2_sub(zero, Y[i]);
        OX[i] = q2_sub(W.one, X[i]);
        OY[i] = q2_sub(W.one, Y[i]);
        GX[i] = q2_pack(q2_lo(X[i]) >= 1.f ? 1.f : 0.f, q2_hi(X[i]) >= 1.f ? 1.f : 0.f);
        GY[i] = q2_pack(q2_lo(Y[i]) >= 1.f ? 1.f : 0.f, q2_hi(Y[i]) >= 1.f ? 1.f : 0.f);
    }
    qf2 acc = zero;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        const qf2 ex = q2_sub(X[j], X[i]), ey = q2_sub(Y[j], Y[i]);
        const float rx0 = q_rcp(q2_lo(ex)), rx1 = q_rcp(q2_hi(ex)), ry0 = q_rcp(q2_lo(ey)), ry1 = q_rcp(q2_hi(ey));
        const qf2 rx = q2_pack(rx0, rx1), ry = q2_pack(ry0, ry1);
        const qf2 ax = q2_mul(NX[i], rx), bx = q2_mul(OX[i], rx);
        const qf2 ay = q2_mul(NY[i], ry), by = q2_mul(OY[i], ry);
        float w0, w1, u10, u11, u20, u21;
        {
            const float a_x = q2_lo(ax), b_x = q2_lo(bx), a_y = q2_lo(ay), b_y = q2_lo(by);
            const float t0 = fmaxf(fmaxf(0.f, fminf(a_x, b_x)), fminf(a_y, b_y));
            const float t1 = fminf(fminf(1.f, fmaxf(a_x, b_x)), fmaxf(a_y, b_y));
            w0 = q_sat(t1 - t0);
            u10 = q_sat(fmaf(b_x, q2_lo(ey), q2_lo(Y[i])));
            u20 = q_sat(fmaf(-b_y, q2_lo(ex), q2_lo(OX[i])));
        }
        {
            const float a_x = q2_hi(ax), b_x = q2_hi(bx), a_y = q2_hi(ay), b_y = q2_hi(by);
            const float t0 = fmaxf(fmaxf(0.f, fminf(a_x, b_x)), fminf(a_y, b_y));
            const float t1 = fminf(fminf(1.f, fmaxf(a_x, b_x)), fmaxf(a_y, b_y));
            w1 = q_sat(t1 - t0);
            u11 = q_sat(fmaf(b_x, q2_hi(ey), q2_hi(Y[i])));
            u21 = q_sat(fmaf(-b_y, q2_hi(ex), q2_hi(OX[i])));
        }
        const qf2 w = q2_pack(w0, w1), u1 = q2_pack(u10, u11), u2 = q2_pack(u20, u21);
        acc = q2_fma(w, q2_fma(NY[i], ex, q2_mul(X[i], ey)), acc);
        const qf2 s1 = q2_mul(q2_sub(GX[i], GX[j]), u1);
        const qf2 s2 = q2_mul(q2_sub(GY[i], GY[j]), u2);
        acc = q2_fma(s1, q2_fma(W.ea, u1, W.one), acc);
        acc = q2_fma(s2, q2_fma(W.k2b, u2, W.k2a), acc);
    }
    const qf2 raw = q2_mul(W.hscale, acc);
    const float a0 = q2_lo(A.area), a1 = q2_hi(A.area);
    const float i0 = fminf(fmaxf(q2_lo(raw), 0.f), fminf(a0, b_area));
    const float i1 = fminf(fmaxf(q2_hi(raw), 0.f), fminf(a1, b_area));
    const qf2 uni = q2_sub(q2_add(A.area, W.barea), q2_pack(i0, i1));
    const float un0 = q2_lo(uni), un1 = q2_hi(uni);
#ifdef __CUDA_ARCH__
    const int v0 = __float_as_int(q2_lo(A.valid)), v1 = __float_as_int(q2_hi(A.valid));
#else
    int v0, v1; { float t0 = q2_lo(A.valid), t1 = q2_hi(A.valid); memcpy(&v0, &t0, 4); memcpy(&v1, &t1, 4); }
#endif
    out0 = ((!kChecked || (v0 & b_valid & 1)) && un0 > 0.f) ? i0 * q_rcp(un0) : 0.f;
    out1 = ((!kChecked || (v1 & b_valid & 1)) && un1 > 0.f) ? i1 * q_rcp(un1) : 0.f;
}

// IoU of polygon A against window B (Bp: the polygon record of the same box B), any convex window.
__host__ __device__ __forceinline__ float qbox_iou_quad(const QPoly& A, const QPoly& Bp, const QWin& Bw) {
    const float dx = A.chx - Bp.chx;
    const float dy = A.chy - Bp.chy;
    float V[6][4];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float c = fmaf(Bw.f[k][0], dx, fmaf(Bw.f[k][1], dy, Bw.f[k][2]));
#pragma unroll
        for (int i = 0; i < 4; ++i) V[k][i] = fmaf(Bw.f[k][0], A.lx[i], fmaf(Bw.f[k][1], A.ly[i], c));
    }
    float acc = 0.f;
    float ulo1 = 2.f, uhi1 = -1.f, ulo2 = 2.f, uhi2 = -1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        float t0 = 0.f, t1 = 1.f;
        bool up;
        q_cut(V[0][i], V[0][j], t0, t1, up);
        q_cut(V[1][i], V[1][j], t0, t1, up);
        bool up1, up2;
        const float tc1 = q_cut(V[2][i], V[2][j], t0, t1, up1);
        const float tc2 = q_cut(V[3][i], V[3][j], t0, t1, up2);
        const float w = fmaxf(t1 - t0, 0.f);
        const float ex = V[0][j] - V[0][i], ey = V[1][j] - V[1][i];
        acc = fmaf(w, V[0][i] * ey - V[1][i] * ex, acc);
        // pieces of window edges 1 and 2 inside A: between the places where A's boundary changes sides
        const bool x1 = (V[2][i] < 0.f) != (V[2][j] < 0.f);
        const bool x2 = (V[3][i] < 0.f) != (V[3][j] < 0.f);
        const float uc1 = fmaf(tc1, V[4][j] - V[4][i], V[4][i]);
        const float uc2 = fmaf(tc2, V[5][j] - V[5][i], V[5][i]);
        uhi1 = (x1 && up1) ? uc1 : uhi1;   ulo1 = (x1 && !up1) ? uc1 : ulo1;
        uhi2 = (x2 && up2) ? uc2 : uhi2;   ulo2 = (x2 && !up2) ? uc2 : ulo2;
    }
    const float len1 = fmaxf(fminf(uhi1, 1.f) - fmaxf(ulo1, 0.f), 0.f);
    const float len2 = fmaxf(fminf(uhi2, 1.f) - fmaxf(ulo2, 0.f), 0.f);
    float inter = 0.5f * Bw.scale * (acc + Bw.beta * len1 + Bw.alpha * len2);
    inter = fminf(fmaxf(inter, 0.f), fminf(A.area, Bp.area));
    const float uni = A.area + Bp.area - inter;
    const bool ok = (A.valid & Bp.valid & 1) && uni > 0.f;          // valid == 2: concave simple quad, float64 path only
    return ok ? inter * q_rcp(uni) : 0.f;
}

// IoU of polygon A against window B: the slab form for parallelogram windows, the general form otherwise.
__host__ __device__ __forceinline__ float qbox_iou(const QPoly& A, const QPoly& Bp, const QWin& Bw) {
    return Bw.rect ? qbox_iou_rect<true>(A, Bp, Bw) : qbox_iou_quad(A, Bp, Bw);
}

// ------------------------------------------------------------------------------------------
// General SIMPLE quads in float64 (shapely's ``is_valid`` semantics, Detect_OBB.py:148-151, :631-636): a concave quad
// whose ring does not touch or cross itself is valid for shapely and is intersected correctly by it; only zero-area
// and self-intersecting (bow-tie, spike, self-touching) rings are invalid.  The detector emits rectangles, so this is
// reachable only with hand-made boxes and ground-truth labels (the evaluation path).  A concave quad is split along
// the diagonal through its reflex vertex into two counter-clockwise triangles; A n B is then the sum over piece pairs
// of convex clips (pieces of one quad are interior-disjoint).  Convex pairs never come here: their arithmetic is the
// single quad-quad clip above, unchanged.

// a*b - c*d with every operation rounded once: nvcc would contract it into an FMA, whose result for a == c, b == d is
// the rounding error of the product instead of 0 - and the zero / sign tests below decide validity (a collapsed label
// must have area exactly 0, as it has for the reference's numpy/shapely arithmetic).
__host__ __device__ inline double gq_det(double a, double b, double c, double d) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(__dmul_rn(a, b), __dmul_rn(c, d));
#else
    volatile double p = a * b, q = c * d;      // volatile: no contraction under -ffp-contract=fast either
    return p - q;
#endif
}
__host__ __device__ inline double gq_orient(double ax, double ay, double bx, double by, double cx, double cy) {
    return gq_det(bx - ax, cy - ay, by - ay, cx - ax);
}
// c collinear with a-b: does it lie on the closed segment?
__host__ __device__ inline bool gq_on_seg(double ax, double ay, double bx, double by, double cx, double cy) {
    return fmin(ax, bx) <= cx && cx <= fmax(ax, bx) && fmin(ay, by) <= cy && cy <= fmax(ay, by);
}
// closed segments a-b and c-d share a point
__host__ __device__ inline bool gq_segs_meet(double ax, double ay, double bx, double by, double cx, double cy, double dx,
                                             double dy) {
    const double d1 = gq_orient(cx, cy, dx, dy, ax, ay), d2 = gq_orient(cx, cy, dx, dy, bx, by);
    const double d3 = gq_orient(ax, ay, bx, by, cx, cy), d4 = gq_orient(ax, ay, bx, by, dx, dy);
    if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0))) return true;
    if (d1 == 0 && gq_on_seg(cx, cy, dx, dy, ax, ay)) return true;
    if (d2 == 0 && gq_on_seg(cx, cy, dx, dy, bx, by)) return true;
    if (d3 == 0 && gq_on_seg(ax, ay, bx, by, cx, cy)) return true;
    if (d4 == 0 && gq_on_seg(ax, ay, bx, by, dx, dy)) return true;
    return false;
}

struct GQuad {
    double x[4], y[4];      // counter-clockwise ring (the input order, reversed as 0,3,2,1 when clockwise)
    double area;
    int kind;               // 0 invalid, 1 convex, 2 concave simple
    int reflex;             // kind 2: index of the reflex vertex in x[], y[]
};

__host__ __device__ inline void gquad_from_corners(const double* __restrict__ b, GQuad& q) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        s += gq_det(b[2 * i], b[2 * j + 1], b[2 * j], b[2 * i + 1]);
    }
    q.kind = 0; q.reflex = 0; q.area = 0.5 * fabs(s);
    const bool flip = s < 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = (flip && (i & 1)) ? (i ^ 2) : i;
        q.x[i] = b[2 * k]; q.y[i] = b[2 * k + 1];
    }
    if (s == 0.0) return;
    int nneg = 0, r = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = (i + 3) & 3, n = (i + 1) & 3;
        const double cr = gq_det(q.x[i] - q.x[p], q.y[n] - q.y[i], q.y[i] - q.y[p], q.x[n] - q.x[i]);
        if (cr < 0.0) { ++nneg; r = i; }
    }
    if (nneg == 0) { q.kind = 1; return; }
    if (gq_segs_meet(q.x[0], q.y[0], q.x[1], q.y[1], q.x[2], q.y[2], q.x[3], q.y[3])) return;
    if (gq_segs_meet(q.x[1], q.y[1], q.x[2], q.y[2], q.x[3], q.y[3], q.x[0], q.y[0])) return;
    if (nneg != 1) return;
    q.kind = 2; q.reflex = r;
}

// Convex pieces of a valid GQuad as degenerate quads (a triangle repeats its last vertex), relative to (ox, oy).
__host__ __device__ inline int gquad_pieces(const GQuad& q, double ox, double oy, double (&px)[2][4], double (&py)[2][4]) {
    if (q.kind == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { px[0][i] = q.x[i] - ox; py[0][i] = q.y[i] - oy; }
        return 1;
    }
    double x[4], y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = (q.reflex + i) & 3;
        x[i] = (k == 0 ? q.x[0] : k == 1 ? q.x[1] : k == 2 ? q.x[2] : q.x[3]) - ox;
        y[i] = (k == 0 ? q.y[0] : k == 1 ? q.y[1] : k == 2 ? q.y[2] : q.y[3]) - oy;
    }
    px[0][0] = x[0]; py[0][0] = y[0]; px[0][1] = x[1]; py[0][1] = y[1]; px[0][2] = x[2]; py[0][2] = y[2]; px[0][3] = x[2]; py[0][3] = y[2];
    px[1][0] = x[0]; py[1][0] = y[0]; px[1][1] = x[2]; py[1][1] = y[2]; px[1][2] = x[3]; py[1][2] = y[3]; px[1][3] = x[3]; py[1][3] = y[3];
    return 2;
}

__host__ __device__ inline double iou_f64_general(const double* __restrict__ rawA, const double* __restrict__ rawB) {
    GQuad a, b;
    gquad_from_corners(rawA, a);
    gquad_from_corners(rawB, b);
    if (a.kind == 0 || b.kind == 0) return 0.0;
    const double ox = 0.25 * ((rawA[0] + rawA[2]) + (rawA[4] + rawA[6])), oy = 0.25 * ((rawA[1] + rawA[3]) + (rawA[5] + rawA[7]));
    double ax[2][4], ay[2][4], bx[2][4], by[2][4];
    const int na = gquad_pieces(a, ox, oy, ax, ay), nb = gquad_pieces(b, ox, oy, bx, by);
    double scratch[GEOM_SCRATCH_WORDS];
    double inter = 0.0;
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) inter += clip_area<double>(ax[i], ay[i], bx[j], by[j], scratch, 1);
    const double uni = a.area + b.area - inter;
    return uni > 0.0 ? inter / uni : 0.0;
}

// Strict interior test (shapely ``Polygon.contains(Point)``) for a concave simple quad: inside one of the two closed
// triangles and not on the ring.
__host__ __device__ inline bool gquad_contains_concave(const GQuad& q, double px, double py) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        if (gq_orient(q.x[i], q.y[i], q.x[j], q.y[j], px, py) == 0.0 && gq_on_seg(q.x[i], q.y[i], q.x[j], q.y[j], px, py)) return false;
    }
    double tx[2][4], ty[2][4];
    gquad_pieces(q, 0.0, 0.0, tx, ty);
    for (int t = 0; t < 2; ++t) {
        const bool in = gq_orient(tx[t][0], ty[t][0], tx[t][1], ty[t][1], px, py) >= 0.0 &&
                        gq_orient(tx[t][1], ty[t][1], tx[t][2], ty[t][2], px, py) >= 0.0 &&
                        gq_orient(tx[t][2], ty[t][2], tx[t][0], ty[t][0], px, py) >= 0.0;
        if (in) return true;
    }
    return false;
}

// Marks a polygon record whose box is a concave SIMPLE quad (valid for shapely): the fp32 forms still return 0 for it,
// callers that decide thresholds (k_discover, k_iou_pairs) see valid == 2 and take iou_f64_general instead.
__host__ __device__ inline void qpoly_mark_concave(const double* __restrict__ b, QPoly& P) {
    if (P.valid) return;
    GQuad g;
    gquad_from_corners(b, g);
    if (g.kind == 2) P.valid = 2;
}

// float64 IoU from raw corners (rare path: thread-private scratch in local memory).
__host__ __device__ inline double iou_f64_from_corners(const double* __restrict__ rawA, const double* __restrict__ rawB) {
    PBox<double> a, b;
    pbox_from_corners<double>(rawA, a);
    pbox_from_corners<double>(rawB, b);
    if (!a.valid || !b.valid) return iou_f64_general(rawA, rawB);      // concave simple quads are valid for shapely
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
