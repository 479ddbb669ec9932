// Host-side harness: runs the __host__ __device__ geometry of csrc/geom.cuh on the CPU so
// the formulation can be checked against the float64 oracle without a GPU (tests only).
// stdin: n, then n*16 doubles (box A corners, box B corners).  stdout: per pair the fp32 clip IoU, the fp64 IoU, the fp32 window (boundary-integral) IoU as dispatched,
// the general-quad window IoU and whether the window took the parallelogram (slab) path.
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../oriented_object_detection_b200/csrc/geom.cuh"

int main() {
    long long n = 0;
    if (fread(&n, sizeof(n), 1, stdin) != 1) return 1;
    std::vector<double> buf((size_t)n * 16);
    if (fread(buf.data(), sizeof(double), buf.size(), stdin) != buf.size()) return 2;
    std::vector<double> out((size_t)n * 5);
    for (long long i = 0; i < n; ++i) {
        const double* a = &buf[(size_t)i * 16];
        const double* b = a + 8;
        PBox<float> fa, fb;
        pbox_from_corners<float>(a, fa);
        pbox_from_corners<float>(b, fb);
        float sf[GEOM_SCRATCH_WORDS];
        out[5 * i] = (double)pbox_iou<float>(fa, fb, sf, 1);
        out[5 * i + 1] = iou_f64_from_corners(a, b);
        QPoly pa, pb2; QWin wa, wb;
        qbox_from_corners(a, pa, wa);
        qbox_from_corners(b, pb2, wb);
        out[5 * i + 2] = (double)qbox_iou(pa, pb2, wb);
        out[5 * i + 3] = (double)qbox_iou_quad(pa, pb2, wb);
        out[5 * i + 4] = (double)wb.rect;
    }
    fwrite(out.data(), sizeof(double), out.size(), stdout);
    return 0;
}
