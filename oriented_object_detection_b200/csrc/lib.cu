// Library-level entry points and the synchronous host-buffer conveniences.
#include <atomic>
#include <mutex>
#include "gm_common.cuh"
#include "geom.cuh"

static std::atomic<long long> g_launches{0};
void gm_note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t gm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int gm_version(void) { return 100; }

extern "C" const char* gm_status_string(int status) {
    switch (status) {
        case GM_OK: return "ok";
        case GM_EINVAL: return "invalid argument";
        case GM_ENOSPC: return "caller buffer or workspace too small";
        case GM_ERANGE: return "size outside the supported range";
        case GM_ENODEV: return "no sm_100 CUDA device";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

extern "C" int gm_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GM_ENODEV;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return GM_ENODEV;
    return major == 10 ? GM_OK : GM_ENODEV;
}

namespace {

std::mutex g_host_mutex;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return GM_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return (int)e;
        cap = bytes;
        return GM_OK;
    }
};

// scratch arena of the *_host helpers, one per device (a pointer from cudaMalloc belongs to the device that was
// current when it was allocated: a process that switches devices between calls must not reuse it on another GPU)
constexpr int GM_MAX_DEVICES = 64;
struct HostArena { DevBuf in, out, ws, tiles; };
HostArena g_arena[GM_MAX_DEVICES];

HostArena* current_arena() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= GM_MAX_DEVICES) return nullptr;
    return &g_arena[dev];
}

__global__ void k_iou_f64(const double* a, const double* b, double* out) {
    *out = iou_f64_from_corners(a, b);
}

}  // namespace

extern "C" int gm_build_multich_host(const uint8_t* bgr_host, int32_t h, int32_t w, int32_t out_channels,
                                     const gm_dtedge_params* params, uint8_t* out_host) {
    if (!bgr_host || !out_host || h <= 0 || w <= 0) return GM_EINVAL;
    if (out_channels != 3 && out_channels != 4) return GM_EINVAL;
    if (out_channels == 4 && !params) return GM_EINVAL;
    const int mt = h > w ? h : w;
    if (out_channels == 4 && mt > GM_MAX_TILE) return GM_ERANGE;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    HostArena* ar = current_arena();
    if (!ar) return GM_ENODEV;
    DevBuf &g_in = ar->in, &g_out = ar->out, &g_ws = ar->ws, &g_tiles = ar->tiles;
    const size_t in_bytes = (size_t)h * w * 3, out_bytes = (size_t)h * w * out_channels;
    int st;
    if ((st = g_in.reserve(in_bytes + 16)) != GM_OK) return st;
    if ((st = g_out.reserve(out_bytes + 16)) != GM_OK) return st;
    if ((st = g_tiles.reserve(sizeof(gm_tile))) != GM_OK) return st;
    gm_tile t; t.y0 = 0; t.x0 = 0; t.h = h; t.w = w; t.px_off = 0;
    GM_CUDA_TRY(cudaMemcpy(g_in.p, bgr_host, in_bytes, cudaMemcpyHostToDevice));
    GM_CUDA_TRY(cudaMemcpy(g_tiles.p, &t, sizeof(t), cudaMemcpyHostToDevice));
    if (out_channels == 3) {
        st = gm_tile_gather_u8((const uint8_t*)g_in.p, h, w, (const gm_tile*)g_tiles.p, 1, mt, (uint8_t*)g_out.p, nullptr);
    } else {
        const size_t wsb = gm_dtedge_workspace_bytes((int64_t)h * w, 1);
        if ((st = g_ws.reserve(wsb)) != GM_OK) return st;
        st = gm_dtedge_build_u8((const uint8_t*)g_in.p, h, w, (const gm_tile*)g_tiles.p, 1, mt, (int64_t)h * w, params,
                                (uint8_t*)g_out.p, g_ws.p, wsb, nullptr);
    }
    if (st != GM_OK) return st;
    GM_CUDA_TRY(cudaMemcpy(out_host, g_out.p, out_bytes, cudaMemcpyDeviceToHost));
    return GM_OK;
}

extern "C" int gm_polygon_iou_host(const double* box1_host, const double* box2_host, double* iou_host) {
    if (!box1_host || !box2_host || !iou_host) return GM_EINVAL;
    std::lock_guard<std::mutex> lock(g_host_mutex);
    HostArena* ar = current_arena();
    if (!ar) return GM_ENODEV;
    DevBuf& g_in = ar->in;
    int st;
    if ((st = g_in.reserve(17 * sizeof(double))) != GM_OK) return st;
    double* d = (double*)g_in.p;
    GM_CUDA_TRY(cudaMemcpy(d, box1_host, 8 * sizeof(double), cudaMemcpyHostToDevice));
    GM_CUDA_TRY(cudaMemcpy(d + 8, box2_host, 8 * sizeof(double), cudaMemcpyHostToDevice));
    k_iou_f64<<<1, 1>>>(d, d + 8, d + 16); gm_note_launches(1);
    GM_LAUNCH_CHECK();
    GM_CUDA_TRY(cudaMemcpy(iou_host, d + 16, sizeof(double), cudaMemcpyDeviceToHost));
    return GM_OK;
}
