/* Plain-C restatement of the geometry half of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithms, same float64 operation order as oracle/geometry.py (compiled with
 * -ffp-contract=off so no FMA changes a rounding), for the 10^5-box cases the pure-Python
 * oracle cannot finish:
 *   orc_quad_iou   Detect_OBB.py:144-154  (shapely 2.0.7 overlay restated: S-H clip, float64)
 *   orc_nms        Detect_OBB.py:176-200  (stable conf-desc sort + greedy class-wise NMS)
 *   orc_fuse       Detect_OBB.py:347-423  (dual-scale late fusion)
 * The only liberty taken: pairs whose axis-aligned bounding boxes are disjoint skip the clip
 * (their IoU is exactly 0 in the reference too).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y; } pt;

static double shoelace2(const pt* p, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const pt a = p[i], b = p[(i + 1) % n];
        s += a.x * b.y - b.x * a.y;
    }
    return s;
}

static int quad_valid(const pt* p) {
    if (shoelace2(p, 4) == 0.0) return 0;
    int pos = 0, neg = 0;
    for (int i = 0; i < 4; ++i) {
        const pt a = p[i], b = p[(i + 1) % 4], c = p[(i + 2) % 4];
        const double cr = (b.x - a.x) * (c.y - b.y) - (b.y - a.y) * (c.x - b.x);
        if (cr > 0) pos = 1; else if (cr < 0) neg = 1;
    }
    return !(pos && neg);
}

static int clip_convex(const pt* subject, int ns, const pt* clipper, int nc, pt* out) {
    pt buf[2][16];
    int n = ns, cur = 0;
    memcpy(buf[0], subject, sizeof(pt) * ns);
    for (int i = 0; i < nc && n > 0; ++i) {
        const pt a = clipper[i], b = clipper[(i + 1) % nc];
        const double ex = b.x - a.x, ey = b.y - a.y;
        const pt* in = buf[cur];
        pt* o = buf[cur ^ 1];
        int m = 0;
        for (int k = 0; k < n; ++k) {
            const pt p = in[k], q = in[(k + 1) % n];
            const double dp = ex * (p.y - a.y) - ey * (p.x - a.x);
            const double dq = ex * (q.y - a.y) - ey * (q.x - a.x);
            if (dp >= 0) {
                o[m++] = p;
                if (dq < 0) { const double t = dp / (dp - dq); o[m].x = p.x + t * (q.x - p.x); o[m].y = p.y + t * (q.y - p.y); ++m; }
            } else if (dq >= 0) {
                const double t = dp / (dp - dq); o[m].x = p.x + t * (q.x - p.x); o[m].y = p.y + t * (q.y - p.y); ++m;
            }
        }
        n = m; cur ^= 1;
    }
    memcpy(out, buf[cur], sizeof(pt) * n);
    return n;
}

/* ---- concave simple quads (valid for shapely): same steps as oracle/geometry.py quad_classify / _general_inter_area */
static double orient(pt a, pt b, pt c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); }
static int on_seg(pt a, pt b, pt c) {
    return fmin(a.x, b.x) <= c.x && c.x <= fmax(a.x, b.x) && fmin(a.y, b.y) <= c.y && c.y <= fmax(a.y, b.y);
}
static int segs_meet(pt a, pt b, pt c, pt d) {
    const double d1 = orient(c, d, a), d2 = orient(c, d, b), d3 = orient(a, b, c), d4 = orient(a, b, d);
    if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0))) return 1;
    return (d1 == 0 && on_seg(c, d, a)) || (d2 == 0 && on_seg(c, d, b)) || (d3 == 0 && on_seg(a, b, c)) || (d4 == 0 && on_seg(a, b, d));
}
/* kind: 0 invalid, 1 convex, 2 concave simple; q = CCW ring; *reflex = reflex vertex of kind 2 */
static int quad_classify(const pt* p, pt* q, int* reflex) {
    const double s = shoelace2(p, 4);
    *reflex = 0;
    for (int i = 0; i < 4; ++i) q[i] = p[i];
    if (s == 0.0) return 0;
    if (s < 0) { q[1] = p[3]; q[3] = p[1]; }
    int nneg = 0;
    for (int i = 0; i < 4; ++i)
        if (orient(q[(i + 3) % 4], q[i], q[(i + 1) % 4]) < 0) { ++nneg; *reflex = i; }
    if (nneg == 0) return 1;
    if (segs_meet(q[0], q[1], q[2], q[3]) || segs_meet(q[1], q[2], q[3], q[0]) || nneg != 1) return 0;
    return 2;
}
static int pieces(int kind, const pt* q, int r, pt out[2][4], int* np) {
    if (kind == 1) { for (int i = 0; i < 4; ++i) out[0][i] = q[i]; np[0] = 4; return 1; }
    const pt a = q[r % 4], b = q[(r + 1) % 4], c = q[(r + 2) % 4], d = q[(r + 3) % 4];
    out[0][0] = a; out[0][1] = b; out[0][2] = c; np[0] = 3;
    out[1][0] = a; out[1][1] = c; out[1][2] = d; np[1] = 3;
    return 2;
}
static double quad_iou_general(const pt* p1, const pt* p2) {
    pt q1[4], q2[4], a[2][4], b[2][4], sa[4], sb[4], poly[16];
    int r1, r2, na[2], nb[2];
    const int k1 = quad_classify(p1, q1, &r1), k2 = quad_classify(p2, q2, &r2);
    if (k1 == 0 || k2 == 0) return 0.0;
    const double ox = 0.25 * ((p1[0].x + p1[1].x) + (p1[2].x + p1[3].x)), oy = 0.25 * ((p1[0].y + p1[1].y) + (p1[2].y + p1[3].y));
    const int ma = pieces(k1, q1, r1, a, na), mb = pieces(k2, q2, r2, b, nb);
    double inter = 0.0;
    for (int i = 0; i < ma; ++i)
        for (int j = 0; j < mb; ++j) {
            for (int k = 0; k < na[i]; ++k) { sa[k].x = a[i][k].x - ox; sa[k].y = a[i][k].y - oy; }
            for (int k = 0; k < nb[j]; ++k) { sb[k].x = b[j][k].x - ox; sb[k].y = b[j][k].y - oy; }
            const int n = clip_convex(sa, na[i], sb, nb[j], poly);
            inter += n >= 3 ? fabs(shoelace2(poly, n)) * 0.5 : 0.0;
        }
    const double a1 = fabs(shoelace2(p1, 4)) * 0.5, a2 = fabs(shoelace2(p2, 4)) * 0.5;
    const double uni = a1 + a2 - inter;
    return uni > 0 ? inter / uni : 0.0;
}

double orc_quad_iou(const double* b1, const double* b2) {
    pt p1[4], p2[4], q1[4], q2[4], poly[16];
    for (int i = 0; i < 4; ++i) { p1[i].x = b1[2 * i]; p1[i].y = b1[2 * i + 1]; p2[i].x = b2[2 * i]; p2[i].y = b2[2 * i + 1]; }
    if (!quad_valid(p1) || !quad_valid(p2)) return quad_iou_general(p1, p2);
    const double s1 = shoelace2(p1, 4), s2 = shoelace2(p2, 4);
    if (s1 < 0) { pt t = p1[0]; p1[0] = p1[3]; p1[3] = t; t = p1[1]; p1[1] = p1[2]; p1[2] = t; }
    if (s2 < 0) { pt t = p2[0]; p2[0] = p2[3]; p2[3] = t; t = p2[1]; p2[1] = p2[2]; p2[2] = t; }
    const double ox = p1[0].x, oy = p1[0].y;
    for (int i = 0; i < 4; ++i) { q1[i].x = p1[i].x - ox; q1[i].y = p1[i].y - oy; q2[i].x = p2[i].x - ox; q2[i].y = p2[i].y - oy; }
    const int n = clip_convex(q1, 4, q2, 4, poly);
    const double inter = n >= 3 ? fabs(shoelace2(poly, n)) * 0.5 : 0.0;
    const double a1 = fabs(s1) * 0.5, a2 = fabs(s2) * 0.5;
    const double uni = a1 + a2 - inter;
    return uni > 0 ? inter / uni : 0.0;
}

static void aabb_of(const double* b, double* o) {
    o[0] = o[2] = b[0]; o[1] = o[3] = b[1];
    for (int i = 1; i < 4; ++i) {
        if (b[2 * i] < o[0]) o[0] = b[2 * i];
        if (b[2 * i] > o[2]) o[2] = b[2 * i];
        if (b[2 * i + 1] < o[1]) o[1] = b[2 * i + 1];
        if (b[2 * i + 1] > o[3]) o[3] = b[2 * i + 1];
    }
}

static int disjoint(const double* a, const double* b) { return a[0] > b[2] || b[0] > a[2] || a[1] > b[3] || b[1] > a[3]; }

/* stable merge sort of indices by conf descending */
static void sort_desc(const float* conf, int* idx, int* tmp, int n) {
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (conf[idx[j]] > conf[idx[i]]) ? idx[j++] : idx[i++];
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, sizeof(int) * n);
    }
}

/* order_out[n]: stable conf-desc permutation; kept_out[<=n]: kept input indices in order. */
int orc_nms(const double* boxes, const int* cls, const float* conf, int n, double thr, int* order_out, int* kept_out) {
    if (n <= 0) return 0;
    int* tmp = (int*)malloc(sizeof(int) * n);
    double* bb = (double*)malloc(sizeof(double) * 4 * n);
    for (int i = 0; i < n; ++i) { order_out[i] = i; aabb_of(boxes + 8 * (size_t)i, bb + 4 * (size_t)i); }
    sort_desc(conf, order_out, tmp, n);
    int nk = 0;
    for (int r = 0; r < n; ++r) {
        const int i = order_out[r];
        int keep = 1;
        for (int k = 0; k < nk; ++k) {
            const int j = kept_out[k];
            if (cls[j] != cls[i] || disjoint(bb + 4 * (size_t)i, bb + 4 * (size_t)j)) continue;
            if (orc_quad_iou(boxes + 8 * (size_t)i, boxes + 8 * (size_t)j) >= thr) { keep = 0; break; }
        }
        if (keep) kept_out[nk++] = i;
    }
    free(tmp); free(bb);
    return nk;
}

/* Detections concatenated in ascending-scale order; scale[n] non-decreasing ids. */
int orc_fuse(const double* boxes, const int* cls, const float* conf, const int* scale, int n, int n_scales,
             double iou_partner, double conf_low, double conf_high, int* kept_out) {
    if (n_scales == 1) { for (int i = 0; i < n; ++i) kept_out[i] = i; return n; }
    char* seen = (char*)calloc(n > 0 ? n : 1, 1);
    double* bb = (double*)malloc(sizeof(double) * 4 * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { aabb_of(boxes + 8 * (size_t)i, bb + 4 * (size_t)i); if (!((double)conf[i] >= conf_low)) seen[i] = 2; }
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (seen[i]) continue;
        int best = -1; double bconf = -1.0, biou = 0.0;
        for (int j = 0; j < n; ++j) {
            if (seen[j] || scale[j] == scale[i] || cls[j] != cls[i]) continue;
            if (disjoint(bb + 4 * (size_t)i, bb + 4 * (size_t)j)) continue;
            const double v = orc_quad_iou(boxes + 8 * (size_t)i, boxes + 8 * (size_t)j);
            if (v >= iou_partner) {
                const double cp = (double)conf[j];
                if (cp > bconf || (cp == bconf && v > biou)) { best = j; bconf = cp; biou = v; }
            }
        }
        seen[i] = 1;
        if (best < 0 || bconf < conf_low) { if ((double)conf[i] >= conf_high) kept_out[nk++] = i; continue; }
        kept_out[nk++] = ((double)conf[i] >= bconf) ? i : best;
        seen[best] = 1;
    }
    free(seen); free(bb);
    return nk;
}

void orc_iou_pairs(const double* a, const double* b, long long n, double* out) {
    for (long long i = 0; i < n; ++i) out[i] = orc_quad_iou(a + 8 * i, b + 8 * i);
}

/* ------------------------------------------------------------------------------------------------
 * Grid-indexed variants for the 10^6-box cases (BASELINE config 4).  Same sequential semantics as
 * orc_nms / orc_fuse; only the candidate lookup changes: a pair whose bounding boxes are disjoint has
 * IoU 0 in the reference, and two boxes with overlapping bounding boxes have centres at most one cell
 * apart when the cell is at least the largest box extent, so the 3x3 cell neighbourhood holds every
 * candidate.  tests/test_oracle_geom.py checks both variants equal to the plain loops. */

typedef struct { double minx, miny, cell; int nx, ny; } grid_t;

static grid_t grid_for(const double* bb, int n) {
    grid_t g; g.minx = g.miny = 0; g.cell = 1; g.nx = g.ny = 1;
    if (n <= 0) return g;
    double maxx = bb[2], maxy = bb[3], ext = 0;
    g.minx = bb[0]; g.miny = bb[1];
    for (int i = 0; i < n; ++i) {
        const double* b = bb + 4 * (size_t)i;
        if (b[0] < g.minx) g.minx = b[0];
        if (b[1] < g.miny) g.miny = b[1];
        if (b[2] > maxx) maxx = b[2];
        if (b[3] > maxy) maxy = b[3];
        if (b[2] - b[0] > ext) ext = b[2] - b[0];
        if (b[3] - b[1] > ext) ext = b[3] - b[1];
    }
    g.cell = ext * 1.001 + 1e-9;
    const double span = (maxx - g.minx > maxy - g.miny) ? maxx - g.minx : maxy - g.miny;
    if (g.cell < span / 2000.0) g.cell = span / 2000.0;
    g.nx = (int)((maxx - g.minx) / g.cell) + 1;
    g.ny = (int)((maxy - g.miny) / g.cell) + 1;
    return g;
}

static void cell_of(const grid_t* g, const double* b, int* cx, int* cy) {
    int x = (int)((0.5 * (b[0] + b[2]) - g->minx) / g->cell), y = (int)((0.5 * (b[1] + b[3]) - g->miny) / g->cell);
    if (x < 0) x = 0;
    if (x >= g->nx) x = g->nx - 1;
    if (y < 0) y = 0;
    if (y >= g->ny) y = g->ny - 1;
    *cx = x; *cy = y;
}

int orc_nms_grid(const double* boxes, const int* cls, const float* conf, int n, double thr, int* order_out, int* kept_out) {
    if (n <= 0) return 0;
    int* tmp = (int*)malloc(sizeof(int) * n);
    double* bb = (double*)malloc(sizeof(double) * 4 * n);
    for (int i = 0; i < n; ++i) { order_out[i] = i; aabb_of(boxes + 8 * (size_t)i, bb + 4 * (size_t)i); }
    sort_desc(conf, order_out, tmp, n);
    const grid_t g = grid_for(bb, n);
    int* head = (int*)malloc(sizeof(int) * (size_t)g.nx * g.ny);
    int* next = (int*)malloc(sizeof(int) * n);
    for (size_t c = 0; c < (size_t)g.nx * g.ny; ++c) head[c] = -1;
    int nk = 0;
    for (int r = 0; r < n; ++r) {
        const int i = order_out[r];
        int cx, cy, keep = 1;
        cell_of(&g, bb + 4 * (size_t)i, &cx, &cy);
        for (int y = cy - 1; y <= cy + 1 && keep; ++y) {
            if (y < 0 || y >= g.ny) continue;
            for (int x = cx - 1; x <= cx + 1 && keep; ++x) {
                if (x < 0 || x >= g.nx) continue;
                for (int j = head[(size_t)y * g.nx + x]; j >= 0; j = next[j]) {      /* kept boxes only */
                    if (cls[j] != cls[i] || disjoint(bb + 4 * (size_t)i, bb + 4 * (size_t)j)) continue;
                    if (orc_quad_iou(boxes + 8 * (size_t)i, boxes + 8 * (size_t)j) >= thr) { keep = 0; break; }
                }
            }
        }
        if (keep) { kept_out[nk++] = i; next[i] = head[(size_t)cy * g.nx + cx]; head[(size_t)cy * g.nx + cx] = i; }
    }
    free(tmp); free(bb); free(head); free(next);
    return nk;
}

static int cmp_int(const void* a, const void* b) { return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b); }

int orc_fuse_grid(const double* boxes, const int* cls, const float* conf, const int* scale, int n, int n_scales,
                  double iou_partner, double conf_low, double conf_high, int* kept_out) {
    if (n_scales == 1) { for (int i = 0; i < n; ++i) kept_out[i] = i; return n; }
    if (n <= 0) return 0;
    char* seen = (char*)calloc(n, 1);
    double* bb = (double*)malloc(sizeof(double) * 4 * n);
    for (int i = 0; i < n; ++i) { aabb_of(boxes + 8 * (size_t)i, bb + 4 * (size_t)i); if (!((double)conf[i] >= conf_low)) seen[i] = 2; }
    const grid_t g = grid_for(bb, n);
    const size_t nc = (size_t)g.nx * g.ny;
    int* start = (int*)calloc(nc + 1, sizeof(int));
    int* member = (int*)malloc(sizeof(int) * n);
    int* cellid = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) { int cx, cy; cell_of(&g, bb + 4 * (size_t)i, &cx, &cy); cellid[i] = cy * g.nx + cx; start[cellid[i] + 1]++; }
    for (size_t c = 0; c < nc; ++c) start[c + 1] += start[c];
    int* fill = (int*)malloc(sizeof(int) * nc);
    for (size_t c = 0; c < nc; ++c) fill[c] = start[c];
    for (int i = 0; i < n; ++i) member[fill[cellid[i]]++] = i;          /* ascending index inside a cell */
    int cap = 64, *cand = (int*)malloc(sizeof(int) * cap);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (seen[i]) continue;
        const int cx = cellid[i] % g.nx, cy = cellid[i] / g.nx;
        int m = 0;
        for (int y = cy - 1; y <= cy + 1; ++y) {
            if (y < 0 || y >= g.ny) continue;
            for (int x = cx - 1; x <= cx + 1; ++x) {
                if (x < 0 || x >= g.nx) continue;
                const size_t c = (size_t)y * g.nx + x;
                for (int k = start[c]; k < start[c + 1]; ++k) {
                    const int j = member[k];
                    if (seen[j] || scale[j] == scale[i] || cls[j] != cls[i]) continue;
                    if (disjoint(bb + 4 * (size_t)i, bb + 4 * (size_t)j)) continue;
                    if (m == cap) { cap *= 2; cand = (int*)realloc(cand, sizeof(int) * cap); }
                    cand[m++] = j;
                }
            }
        }
        qsort(cand, m, sizeof(int), cmp_int);                           /* the reference walks j in list order */
        int best = -1; double bconf = -1.0, biou = 0.0;
        for (int k = 0; k < m; ++k) {
            const int j = cand[k];
            const double v = orc_quad_iou(boxes + 8 * (size_t)i, boxes + 8 * (size_t)j);
            if (v >= iou_partner) {
                const double cp = (double)conf[j];
                if (cp > bconf || (cp == bconf && v > biou)) { best = j; bconf = cp; biou = v; }
            }
        }
        seen[i] = 1;
        if (best < 0 || bconf < conf_low) { if ((double)conf[i] >= conf_high) kept_out[nk++] = i; continue; }
        kept_out[nk++] = ((double)conf[i] >= bconf) ? i : best;
        seen[best] = 1;
    }
    free(seen); free(bb); free(start); free(member); free(cellid); free(fill); free(cand);
    return nk;
}
