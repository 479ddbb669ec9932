// Host-side harness: runs the __host__ __device__ per-tile arithmetic of csrc/dtedge_otsu.cuh (what thread 0 of
// k_otsu_grad executes, plus the per-pixel acc8 map its histogram pass applies) on the CPU, so the Otsu
// binarisation can be checked against the oracle without a GPU (tests only).
// stdin: n_tiles, then per tile: n (int64) and n uint32 values of S.
// stdout per tile: s_thr, thr8, bits of the 0..255 constants (fs, fh), bits of the 0..1 constants (ns, nh)  (6 x uint32).
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../../oriented_object_detection_b200/csrc/dtedge_otsu.cuh"

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main() {
    long long nt = 0;
    if (fread(&nt, sizeof(nt), 1, stdin) != 1) return 1;
    for (long long t = 0; t < nt; ++t) {
        long long n = 0;
        if (fread(&n, sizeof(n), 1, stdin) != 1) return 2;
        std::vector<uint32_t> S((size_t)n);
        if (fread(S.data(), 4, S.size(), stdin) != S.size()) return 3;
        const uint32_t kmin = *std::min_element(S.begin(), S.end()), kmax = *std::max_element(S.begin(), S.end());
        float fs, fh, ns, nh;
        otsu::normalize_constants(otsu::f_sqrt((float)kmin), otsu::f_sqrt((float)kmax), 0.0, 255.0, &fs, &fh);
        unsigned int hist[256] = {0};
        for (uint32_t v : S) hist[otsu::acc8_of(v, fs, fh)]++;
        const int thr8 = otsu::threshold_from_hist(hist, n);
        const uint32_t s_thr = otsu::s_threshold(kmin, kmax, thr8, fs, fh);
        otsu::normalize_constants(otsu::f_sqrt((float)kmin), otsu::f_sqrt((float)kmax), 0.0, 1.0, &ns, &nh);
        const uint32_t out[6] = {s_thr, (uint32_t)thr8, bits(fs), bits(fh), bits(ns), bits(nh)};
        fwrite(out, 4, 6, stdout);
    }
    return 0;
}
