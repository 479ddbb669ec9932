// Host-side harness: runs the __host__ __device__ geometry of csrc/geom.cuh on the CPU so
// the formulation can be checked against the float64 oracle without a GPU (tests only).
// stdin: n, then n*16 doubles (box A corners, box B corners).  stdout: per pair the fp32 clip IoU, the fp64 IoU, the fp32 window (boundary-integral) IoU as dispatched,
// the general-quad window IoU, whether the window took the parallelogram (slab) path, and the packed two-polygon form (3 columns).
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstring>
#include "../../oriented_object_detection_b200/csrc/geom.cuh"

int main() {
    long long n = 0;
    if (fread(&n, sizeof(n), 1, stdin) != 1) return 1;
    std::vector<double> buf((size_t)n * 16);
    if (fread(buf.data(), sizeof(double), buf.size(), stdin) != buf.size()) return 2;
    std::vector<double> out((size_t)n * 8);
    for (long long i = 0; i < n; ++i) {
        const double* a = &buf[(size_t)i * 16];
        const double* b = a + 8;
        PBox<float> fa, fb;
        pbox_from_corners<float>(a, fa);
        pbox_from_corners<float>(b, fb);
        float sf[GEOM_SCRATCH_WORDS];
        out[8 * i] = (double)pbox_iou<float>(fa, fb, sf, 1);
        out[8 * i + 1] = iou_f64_from_corners(a, b);
        QPoly pa, pb2; QWin wa, wb;
        qbox_from_corners(a, pa, wa);
        qbox_from_corners(b, pb2, wb);
        out[8 * i + 2] = (double)qbox_iou(pa, pb2, wb);
        out[8 * i + 3] = (double)qbox_iou_quad(pa, pb2, wb);
        out[8 * i + 4] = (double)wb.rect;
        // packed two-polygon form (qbox_iou_rect2, struct emulation of the f32x2 arithmetic): lane 0 = A_i, lane 1 = the
        // next pair's A against THIS window; column 7 = the scalar form on (A_next, B_i) for comparison
        const double* a2 = &buf[(size_t)((i + 1) % n) * 16];
        QPoly pn; QWin wn;
        qbox_from_corners(a2, pn, wn);
        out[8 * i + 5] = out[8 * i + 6] = out[8 * i + 7] = NAN;
        if (wb.rect) {
            QPoly2 two;
            two.chx = q2_pack(pa.chx, pn.chx); two.clx = q2_pack(pa.clx, pn.clx);
            two.chy = q2_pack(pa.chy, pn.chy); two.cly = q2_pack(pa.cly, pn.cly);
            for (int k = 0; k < 4; ++k) { two.lx[k] = q2_pack(pa.lx[k], pn.lx[k]); two.ly[k] = q2_pack(pa.ly[k], pn.ly[k]); }
            two.area = q2_pack(pa.area, pn.area);
            float va, vn; memcpy(&va, &pa.valid, 4); memcpy(&vn, &pn.valid, 4);
            two.valid = q2_pack(va, vn);
            QWin2 w2;
            qwin2_from(pb2, wb, w2);
            float r0, r1;
            qbox_iou_rect2(two, w2, pb2.valid, pb2.area, r0, r1);
            out[8 * i + 5] = r0; out[8 * i + 6] = r1;
            out[8 * i + 7] = (double)qbox_iou_rect(pn, pb2, wb);
        }
    }
    fwrite(out.data(), sizeof(double), out.size(), stdout);
    return 0;
}
