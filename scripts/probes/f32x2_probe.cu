// Throughput of the FP32 pipe on sm_100a by instruction form (events, full grid): scalar FFMA with constant-bank
// operands, scalar FFMA with three distinct register operands, and the packed FFMA2 / FADD2 / FMUL2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float b, float c, float* sink) {
    float a[8]; f2 p[8];
    float bb[8], cc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = pk(a[i], a[i] + 0.5f); bb[i] = b + i * 1e-9f; cc[i] = c + i * 1e-9f; }
    const f2 pb = pk(b, b + 1e-9f), pc = pk(c, c + 1e-9f);
    f2 pbb[8], pcc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { pbb[i] = pk(bb[i], bb[i]); pcc[i] = pk(cc[i], cc[i]); }
    for (int it = 0; it < ITER; it += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fmaf(a[i], b, c);                                   // FFMA R, R, c[], c[]
                if (MODE == 1) a[i] = fmaf(a[i], bb[i], cc[i]);                            // three distinct registers
                if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pbb[i]), "l"(pcc[i]));
                if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pbb[i]));
                if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pbb[i]));
                if (MODE == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));   // shared operand pairs
                if (MODE == 6) a[i] = a[i] + bb[i];                                       // FADD two registers
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + lo(p[i]);
    if (s == 123.456f) sink[0] = s;
}
template <int MODE> void run(const char* name, double lane_ops_per_inst, float* sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(1.0000001f, 1e-7f, sink);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(1.0000001f, 1e-7f, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double insts = (double)blocks * 8 /*warps*/ * ITER * 8.0;                      // warp instructions
    const double per_smsp_clk = insts / (148.0 * 4.0) / (best * 1e-3 * 1.965e9);
    printf("%-34s %8.3f ms  %6.3f warp-inst/clk/SMSP  %6.1f lane-ops/clk/SM  (%.1f Tlane-op/s)\n", name, best, per_smsp_clk,
           per_smsp_clk * 4 * 32 * lane_ops_per_inst, insts * 32 * lane_ops_per_inst / (best * 1e-3) / 1e12);
}
int main() {
    float* sink; cudaMalloc(&sink, 4);
    run<0>("FFMA  reg, const, const", 1, sink);
    run<1>("FFMA  3 distinct registers", 1, sink);
    run<6>("FADD  2 registers", 1, sink);
    run<2>("FFMA2 3 distinct register pairs", 2, sink);
    run<5>("FFMA2 shared b, c pairs", 2, sink);
    run<3>("FADD2 2 register pairs", 2, sink);
    run<4>("FMUL2 2 register pairs", 2, sink);
    return 0;
}
