// Shared helpers for the geomap_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/geomap_b200.h"

#define GM_CUDA_TRY(expr)                          \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

#define GM_LAUNCH_CHECK()                          \
    do {                                           \
        cudaError_t _e = cudaGetLastError();       \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

static inline cudaStream_t gm_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline size_t gm_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves a caller-provided workspace into aligned sub-buffers.
struct GmArena {
    uint8_t* base;
    size_t off;
    size_t cap;
    __host__ GmArena(void* p, size_t bytes) : base(static_cast<uint8_t*>(p)), off(0), cap(bytes) {}
    template <typename T>
    __host__ T* take(size_t count) {
        off = gm_align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    __host__ bool ok() const { return off <= cap; }
};

// cv2 BORDER_REFLECT_101 with repeated reflection (n == 1 -> 0).
__device__ __forceinline__ int gm_reflect101(int i, int n) {
    if (static_cast<unsigned>(i) < static_cast<unsigned>(n)) return i;
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    int m = i % period;
    if (m < 0) m += period;
    return m >= n ? period - m : m;
}

__device__ __forceinline__ int gm_lane() { return threadIdx.x & 31; }

constexpr int GM_NUM_SMS_B200 = 148;

// Integer tuning knob from the environment (read once per call site by the caller).
#include <cstdlib>
static inline int gm_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Kernel-launch bookkeeping for bench.py's `gpu_launches` (host side, relaxed atomic).
void gm_note_launches(int n);
