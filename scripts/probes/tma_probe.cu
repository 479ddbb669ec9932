// Stand-alone probe of the TMA load used by k_grad_fast<true>: one cp.async.bulk.tensor.2d of a BOXW x BOXH byte box from a
// uint8 [H][PITCH] tensor into shared memory, copied back for comparison.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
#ifndef BOXW
#define BOXW 144
#endif
#ifndef BOXH
#define BOXH 80
#endif
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, unsigned char* out) {
    __shared__ __align__(128) unsigned char buf[BOXW * BOXH];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BOXW * BOXH) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BOXW * BOXH; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int H = 700, W3 = 1968;
    std::vector<unsigned char> h((size_t)H * W3);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, BOXW * BOXH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {(cuuint64_t)W3, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W3};
    cuuint32_t box[2] = {BOXW, BOXH}, es[2] = {1, 1};
    CUresult r = ((enc_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d (query %d)\n", (int)r, (int)q);
    // the inner coordinate must be a multiple of 16 bytes: (-24, -8) and (7, 3) trap with 'illegal instruction' (measured)
    const int cs[5][2] = {{48, 16}, {-32, -8}, {W3 - 96, H - 30}, {0, 3}, {-160, -80}};
    for (int t = 0; t < 5; ++t) {
        k<<<1, 256>>>(tm, cs[t][0], cs[t][1], o);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<unsigned char> got(BOXW * BOXH);
        cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int y = 0; y < BOXH; ++y) for (int x = 0; x < BOXW; ++x) {
            const int gx = cs[t][0] + x, gy = cs[t][1] + y;
            const unsigned char want = (gx >= 0 && gx < W3 && gy >= 0 && gy < H) ? h[(size_t)gy * W3 + gx] : 0;
            bad += got[y * BOXW + x] != want;
        }
        printf("box %dx%d at (%d,%d): %s, mismatches %ld\n", BOXW, BOXH, cs[t][0], cs[t][1], cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
