// k_grad_fast: the gradient-energy stage of the DT-Edge build for the reference's configured
// scale stack MS_SIGMAS = (0, 0.6, 1.2, 2.4) (Detect_OBB.py:29, Train_OBB.py:765), i.e. one
// unblurred scale plus three Gaussian scales of radius 2, 4 and 7 whose 8.8 fixed-point taps all
// fit a byte.  Any other stack takes the generic k_grad in dtedge.cu.  Same arithmetic as the
// generic kernel; what changes is how it is issued:
//
//   gray      4 pixels (12 map bytes) per thread from aligned 32-bit loads; the BGR weights are
//             16-bit, so each pixel is two IDP.2A (u16 x u8 dot products) on the raw words.
//   pass 1    horizontal 8.8 taps as IDP.4A on aligned words of the gray rows: the taps of output
//             column 4q+e are pre-shifted by the host into the byte lanes of words q..q+4, so no
//             byte is ever extracted.  One thread produces 4 columns of all three scales from the
//             same five words and stores them TRANSPOSED (column-major u16) for pass 2.
//   pass 2    vertical taps as IDP.2A on words holding two consecutive rows of one column
//             (pre-shifted tap pairs again); the rounding constant 32768 is the accumulator seed.
//             No intermediate rounding exists between the passes in cv2 (the horizontal sums are
//             exact u16, the vertical sums exact u32), so the order of the passes is free.
//   Scharr    mixed-sign IDP.4A (u8 pixels x s8 weights) on 3-byte windows of the blurred rows,
//             4 pixels per thread, running max over the 4 scales in registers, one 16-byte store.
//
// 32x64 output pixels per CTA (GM_GRAD_BH rows), 256 threads, 27.6 KB of shared memory, 7 CTAs per SM at 32 registers; the
// raw BGR patch arrives by one cp.async.bulk.tensor.2d (TMA) when the map's row pitch is a multiple of 16 bytes.  Grid =
// (column blocks, row blocks, tiles): the blocks of one tile are resident together, so their halo overlap is served by L2.
#pragma once
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "gm_common.cuh"

#ifndef GM_GRAD_BH
#define GM_GRAD_BH 64                   // rows per CTA: 64 halves the vertical halo overhead of 32 (measured 0.89 -> see DESIGN.md)
#endif

#ifndef GM_GRAD_MINB
#define GM_GRAD_MINB 7                  // minimum resident CTAs per SM asked of the register allocator: 32 registers, no spills (measured: 1 -> 6: 0.862 -> 0.828 ms; 6 -> 7: 0.685 -> 0.676 ms on c3, 21.56 -> 21.28 ms in the c5 step; 8: the same)
#endif

namespace gradfast {

constexpr int BW = 32, BH = GM_GRAD_BH;  // output block (columns x rows)
constexpr int HALO = 8;                 // max radius 7 + 1 (Scharr)
constexpr int PH = BH + 2 * HALO;       // gray patch rows
constexpr int PWW = 13;                 // gray patch row pitch in words (52 bytes, odd -> conflict-free column walks)
constexpr int THREADS = 256;
constexpr int NCOL = BW + 2;            // blurred columns kept (x = -1 .. 32 of the block)
constexpr int NROW = BH + 2;
constexpr int BLUR_PITCH = 36;          // bytes per blurred row

__host__ __device__ constexpr int radius_of(int s) { return s == 0 ? 2 : (s == 1 ? 4 : 7); }
__host__ __device__ constexpr int hrows_of(int R) { return BH + 2 + 2 * R; }                // rows of the horizontal pass
__host__ __device__ constexpr int hpitch_words_of(int R) { return ((hrows_of(R) + 1) / 2) | 1; }   // odd word pitch
constexpr int NCOL_PAD = 36;            // columns the horizontal pass WRITES (9 groups of 4): two more than are read, so its stores need no bounds test
__host__ __device__ constexpr int hwords_total(int R) { return NCOL_PAD * hpitch_words_of(R); }

constexpr int OFF_GRAY = 0;
constexpr int OFF_H0 = OFF_GRAY + PH * PWW;                       // word offsets
constexpr int OFF_H1 = OFF_H0 + hwords_total(2);
constexpr int OFF_H2 = OFF_H1 + hwords_total(4);
constexpr int OFF_B0 = OFF_H2 + hwords_total(7);
constexpr int BLUR_WORDS = NROW * BLUR_PITCH / 4;
constexpr int SMEM_WORDS = OFF_B0 + 3 * BLUR_WORDS;

// TMA variant: the raw BGR patch (PH rows x 48 pixels x 3 bytes, row pitch 144 B) lands in the region the horizontal
// pass later fills (dead until then), 128-byte aligned as cp.async.bulk.tensor requires.
// The box must START on a 16-byte boundary of the map row as well (cp.async.bulk.tensor traps otherwise - measured with
// scripts/probes/tma_probe.cu), so it begins at the patch's first byte rounded down to 16 and is 16 bytes wider: the patch
// sits at byte offset delta = (3 x_first) & 15 of every row, the same for all threads of the block.
constexpr int BGR_ROW_BYTES = (BW + 2 * HALO) * 3 + 16;            // 160: a multiple of 16, as the tensor-map box needs
constexpr int BGR_ROW_WORDS = BGR_ROW_BYTES / 4;                   // 40
constexpr int OFF_BGR = (OFF_H0 + 31) & ~31;
constexpr int BGR_BYTES = PH * BGR_ROW_BYTES;
static_assert(OFF_BGR + BGR_BYTES / 4 <= OFF_B0, "the BGR patch must fit the horizontal-pass region it aliases");
static_assert(BGR_ROW_BYTES % 16 == 0 && BGR_ROW_BYTES <= 256 && PH <= 256, "tensor-map box limits");

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One 2-D tile of a uint8 tensor (dim 0 = bytes of a map row, dim 1 = map rows) -> shared memory; coordinates may be
// negative or run past the tensor: the TMA unit fills what lies outside with zeros.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

// Host-packed tap words (kernel parameter -> constant bank operands of the IDP instructions).
struct Coef {
    unsigned int h[3][4][5];    // pass 1: [scale][e = column within the group of 4][data word q+k]
    unsigned int v[3][2][4];    // pass 2: [scale][e = row within the pair][tap bytes for data words 2m, 2m+1]
};

// taps[s] has 2R+1 entries (R = 2, 4, 7), each <= 255.
inline void pack_coef(const unsigned short taps[3][15], Coef* c) {
    for (int s = 0; s < 3; ++s) {
        const int R = radius_of(s);
        for (int e = 0; e < 4; ++e)
            for (int k = 0; k < 5; ++k) {
                unsigned int word = 0;
                for (int j = 0; j < 4; ++j) {
                    const int ti = 4 * k + j - (e + 7 - R);      // gray byte 4q+4k+j feeds output 4q+e with tap ti
                    if (ti >= 0 && ti <= 2 * R) word |= (unsigned int)(taps[s][ti] & 255u) << (8 * j);
                }
                c->h[s][e][k] = word;
            }
        for (int e = 0; e < 2; ++e)
            for (int m = 0; m < 4; ++m) {
                unsigned int word = 0;
                for (int j = 0; j < 4; ++j) {
                    const int ti = 4 * m + j - e;                // row 2p+4m+j of the column feeds output row 2p+e
                    if (ti >= 0 && ti <= 2 * R) word |= (unsigned int)(taps[s][ti] & 255u) << (8 * j);
                }
                c->v[s][e][m] = word;
            }
    }
}

__device__ __forceinline__ int dp4a_us(unsigned int a, int b, int c) {       // u8 data x s8 weights
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// gray of 4 consecutive pixels whose 12 bytes are v0 v1 v2 (B G R B G R ...), packed into one word
__device__ __forceinline__ unsigned int gray4(unsigned int v0, unsigned int v1, unsigned int v2) {
    constexpr unsigned int CB = 3735u, CG = 19235u, CR = 9798u, RND = 16384u;
    // pixel 0: bytes 0,1,2 of v0
    unsigned int g0 = __dp2a_lo(CB | (CG << 16), v0, RND);
    g0 = __dp2a_hi(CR, v0, g0);
    // pixel 1: byte 3 of v0, bytes 0,1 of v1
    unsigned int g1 = __dp2a_hi(CB << 16, v0, RND);
    g1 = __dp2a_lo(CG | (CR << 16), v1, g1);
    // pixel 2: bytes 2,3 of v1, byte 0 of v2
    unsigned int g2 = __dp2a_hi(CB | (CG << 16), v1, RND);
    g2 = __dp2a_lo(CR, v2, g2);
    // pixel 3: bytes 1,2,3 of v2
    unsigned int g3 = __dp2a_lo(CB << 16, v2, RND);
    g3 = __dp2a_hi(CG | (CR << 16), v2, g3);
    return (g0 >> 15) | ((g1 >> 15) << 8) | ((g2 >> 15) << 16) | ((g3 >> 15) << 24);
}

template <int S>
__device__ __forceinline__ void hpass_scale(const unsigned int (&w)[5], const Coef& c, unsigned short* hT,
                                            int p, int q) {
    constexpr int R = radius_of(S);
    constexpr int PITCH = 2 * hpitch_words_of(R);              // u16 units
    const int i = p - (HALO - 1) + R;                          // row index inside this scale's band
    if (i < 0 || i >= hrows_of(R)) return;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        unsigned int acc = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (k >= ((e + 7 - R) >> 2) && k <= ((e + 7 + R) >> 2)) acc = __dp4a(w[k], c.h[S][e][k], acc);
        const int o = 4 * q + e;
        hT[o * PITCH + i] = (unsigned short)acc;               // o < NCOL_PAD always
    }
}

// One task = one column and FOUR consecutive rows (two row pairs) of one scale: the R + 2 words of the
// column are loaded once and feed 4 (R + 1) dot products; the last task of a column holds rows 32..35, of
// which only 32 and 33 exist (its extra words are padding or the next column's, never stored).
template <int S>
__device__ __forceinline__ void vpass_scale(const unsigned int* hT, const Coef& c, unsigned char* blur, int j, int o) {
    constexpr int R = radius_of(S);
    constexpr int PW = hpitch_words_of(R);                     // rows 4j .. 4j+3 of the kept rows, column o
    const unsigned int* col = hT + o * PW + 2 * j;
    unsigned int d[R + 2];
#pragma unroll
    for (int k = 0; k < R + 2; ++k) d[k] = col[k];
    unsigned int a[4] = {32768u, 32768u, 32768u, 32768u};
#pragma unroll
    for (int k = 0; k <= R; ++k) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (k & 1) {
                a[e] = __dp2a_hi(d[k], c.v[S][e][k >> 1], a[e]);
                a[2 + e] = __dp2a_hi(d[k + 1], c.v[S][e][k >> 1], a[2 + e]);
            } else {
                a[e] = __dp2a_lo(d[k], c.v[S][e][k >> 1], a[e]);
                a[2 + e] = __dp2a_lo(d[k + 1], c.v[S][e][k >> 1], a[2 + e]);
            }
        }
    }
    unsigned char* dst = blur + (4 * j) * BLUR_PITCH + o;
    dst[0] = (unsigned char)(a[0] >> 16);
    dst[BLUR_PITCH] = (unsigned char)(a[1] >> 16);
    if (4 * j + 2 < NROW) {
        dst[2 * BLUR_PITCH] = (unsigned char)(a[2] >> 16);
        dst[3 * BLUR_PITCH] = (unsigned char)(a[3] >> 16);
    }
}

// Scharr energy of two horizontally adjacent pixels whose 3x3 neighbourhoods are bytes 0..2 (first)
// and 1..3 (second) of the three row words.
__device__ __forceinline__ void scharr2(unsigned int r0, unsigned int r1, unsigned int r2,
                                        unsigned int& s_first, unsigned int& s_second) {
    constexpr int X3 = 0x000300FD, X10 = 0x000A00F6, YP = 0x00030A03, YN = 0x00FDF6FD;
    {
        const int gx = dp4a_us(r2, X3, dp4a_us(r1, X10, dp4a_us(r0, X3, 0)));
        const int gy = dp4a_us(r2, YP, dp4a_us(r0, YN, 0));
        s_first = max(s_first, (unsigned int)(gx * gx + gy * gy));
    }
    {
        const int gx = dp4a_us(r2, X3 << 8, dp4a_us(r1, X10 << 8, dp4a_us(r0, X3 << 8, 0)));
        const int gy = dp4a_us(r2, YP << 8, dp4a_us(r0, (int)((unsigned int)YN << 8), 0));
        s_second = max(s_second, (unsigned int)(gx * gx + gy * gy));
    }
}

// kTma: the BGR patch of the block is fetched by ONE cp.async.bulk.tensor.2d (TMA) issued by one thread - no per-thread
// address arithmetic, alignment shifts or bounds tests - and converted to gray from shared memory; requires a 16-byte
// aligned map base and row pitch (3 W a multiple of 16).  Rows and columns outside the TILE arrive as whatever the map
// holds there (or zeros outside the map) and are overwritten by the REFLECT_101 mirror passes of the border blocks.
template <bool kTma>
__global__ void __launch_bounds__(THREADS, GM_GRAD_MINB)
k_grad_fast(const uint8_t* __restrict__ map, int W, long long map_bytes, const gm_tile* __restrict__ tiles,
            const __grid_constant__ Coef coef, unsigned int* __restrict__ S_out, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) unsigned int sm[SMEM_WORDS];
    __shared__ __align__(8) unsigned long long tma_bar;
    // Grid = (column blocks, row blocks, tiles): the hardware hands out x fastest, so the blocks of ONE tile are resident
    // together and their overlapping halo patches - and the 100-px overlap with the neighbouring tiles - are served by L2;
    // with the tile index fastest every patch of a sweep came from DRAM (0.64 GB read per 8192^2 map against 0.20 GB of map).
    // The host passes `tiles` already offset when a plan has more than 65,535 tiles (several launches).
    const gm_tile t = tiles[blockIdx.z];
    // block coordinates straight from the 3-D grid (a runtime division per thread costs as much as a pixel of the stage)
    const int bx = (int)blockIdx.x * BW;
    const int by = (int)blockIdx.y * BH;
    if (bx >= t.w || by >= t.h) return;
    const int tid = threadIdx.x;
    unsigned int* gray = sm + OFF_GRAY;

    // ---- gray patch: byte b of row p <-> tile pixel (by - 8 + p, bx - 8 + b), REFLECT_101 at the tile border.
    // Rows are reflected by index when loading.  Columns: 4-pixel groups that lie inside the tile are
    // converted from the map (all of a thread's loads are issued before the first conversion); the bytes
    // left and right of the tile are then mirrored inside shared memory (border CTAs only), so the
    // division-heavy reflect and the byte loads never run in the common case.
    const int rows_valid = min(BH, t.h - by);                   // output rows of this block inside the tile
    const int p_need = rows_valid + 2 * HALO;                   // patch rows any of them needs
    if constexpr (kTma) {
        unsigned int* bgr = sm + OFF_BGR;
        if (tid == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        const int xb = (t.x0 + bx - HALO) * 3;                     // first byte of the patch in its map row (may be negative)
        const int xb16 = xb & ~15;                                  // ... rounded down to the 16-byte boundary the box starts on
        if (tid == 0) {
            mbar_expect_tx(&tma_bar, BGR_BYTES);
            tma_load_2d(bgr, &tmap, xb16, t.y0 + by - HALO, &tma_bar);
        }
        const int delta = xb - xb16;
        const unsigned int* bgr_w = bgr + (delta >> 2);
        const unsigned int shb = (unsigned int)(delta & 3) * 8u;
        mbar_wait(&tma_bar, 0);
        // 4 pixels = 12 bytes at a block-uniform offset of the patch row: four words, three funnel shifts by the same
        // amount -> one gray word (the 13th word of a gray row only ever meets zero taps)
        for (int task = tid; task < p_need * 12; task += THREADS) {
            const int p = task / 12;
            const int g = task - p * 12;
            const unsigned int* src = bgr_w + p * BGR_ROW_WORDS + 3 * g;
            const unsigned int w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3];
            gray[p * PWW + g] = gray4(__funnelshift_r(w0, w1, shb), __funnelshift_r(w1, w2, shb), __funnelshift_r(w2, w3, shb));
        }
        if (tid < p_need) gray[tid * PWW + 12] = 0u;
        if (by < HALO || by + rows_valid + HALO > t.h) {
            // rows above / below the tile: REFLECT_101 copies of rows inside it (always inside the patch)
            __syncthreads();
            for (int i = tid; i < p_need * PWW; i += THREADS) {
                const int p = i / PWW;
                const int ty = by - HALO + p;
                if ((unsigned)ty >= (unsigned)t.h) {
                    const int sp = gm_reflect101(ty, t.h) - by + HALO;
                    gray[i] = gray[sp * PWW + (i - p * PWW)];
                }
            }
        }
    } else {
    constexpr int GRAY_TASKS = PH * PWW;
    constexpr int GRAY_ITERS = (GRAY_TASKS + THREADS - 1) / THREADS;
    {
        unsigned int w4[GRAY_ITERS][4];
        unsigned int shv[GRAY_ITERS];
        int kind[GRAY_ITERS];                   // 0 nothing, 1 aligned-word path, 2 per-pixel path
#pragma unroll
        for (int it = 0; it < GRAY_ITERS; ++it) {
            const int task = tid + it * THREADS;
            kind[it] = 0;
            if (task < p_need * PWW) {
                const int p = task / PWW;
                const int g = task - p * PWW;
                int ty = by - HALO + p;
                if ((unsigned)ty >= (unsigned)t.h) {
                    ty = ty < 0 ? -ty : 2 * t.h - 2 - ty;                          // one bounce covers every tile taller than the halo
                    if ((unsigned)ty >= (unsigned)t.h) ty = gm_reflect101(by - HALO + p, t.h);
                }
                const int xf = bx - HALO + 4 * g;
                const long long a_off = ((long long)(t.y0 + ty) * W + t.x0 + xf) * 3LL;
                if (xf >= 0 && xf + 3 < t.w && a_off + 16 <= map_bytes) {
                    const unsigned long long sa = reinterpret_cast<unsigned long long>(map + a_off);
                    const unsigned int* sw = reinterpret_cast<const unsigned int*>(sa & ~3ULL);
                    shv[it] = (unsigned int)(sa & 3ULL) * 8u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) w4[it][k] = __ldg(sw + k);
                    kind[it] = 1;
                } else if (xf + 3 >= 0 && xf < t.w) {
                    kind[it] = 2;
                }
            }
        }
#pragma unroll
        for (int it = 0; it < GRAY_ITERS; ++it) {
            const int task = tid + it * THREADS;
            if (kind[it] == 1) {
                gray[task] = gray4(__funnelshift_r(w4[it][0], w4[it][1], shv[it]), __funnelshift_r(w4[it][1], w4[it][2], shv[it]),
                                   __funnelshift_r(w4[it][2], w4[it][3], shv[it]));
            } else if (kind[it] == 2) {
                // group cut by the tile's edge or by the end of the map buffer: per pixel, in-tile pixels only
                const int p = task / PWW;
                const int g = task - p * PWW;
                const int ty = gm_reflect101(by - HALO + p, t.h);
                const int xf = bx - HALO + 4 * g;
                unsigned int packed = 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int tx = xf + e;
                    if (tx >= 0 && tx < t.w) {
                        const uint8_t* px = map + ((long long)(t.y0 + ty) * W + t.x0 + tx) * 3LL;
                        const unsigned int b = __ldg(px), gg = __ldg(px + 1), r = __ldg(px + 2);
                        packed |= ((3735u * b + 19235u * gg + 9798u * r + 16384u) >> 15) << (8 * e);
                    }
                }
                gray[task] = packed;
            }
        }
    }
    }
    __syncthreads();
    if (bx < HALO || bx + BW + HALO > t.w) {
        // mirror the columns outside the tile (patch bytes whose x is < 0 or >= w) from inside it
        unsigned char* g8 = reinterpret_cast<unsigned char*>(gray);
        const int b_lo = max(0, HALO - bx);                      // bytes [0, b_lo) are left of the tile
        const int b_hi = min(PWW * 4, t.w - bx + HALO);          // bytes [b_hi, 52) are right of it
        const int n_fill = b_lo + (PWW * 4 - b_hi);
        for (int i = tid; i < p_need * n_fill; i += THREADS) {
            const int p = i / n_fill;
            const int k = i - p * n_fill;
            const int b = k < b_lo ? k : b_hi + (k - b_lo);
            const int x = bx - HALO + b;
            int sx = x < 0 ? -x : 2 * t.w - 2 - x;
            if ((unsigned)sx >= (unsigned)t.w) sx = gm_reflect101(x, t.w);
            const int sb = min(max(sx - bx + HALO, 0), PWW * 4 - 1);     // columns no output needs may fall outside the patch
            g8[p * (PWW * 4) + b] = g8[p * (PWW * 4) + sb];
        }
        __syncthreads();
    }

    // ---- pass 1: horizontal taps of the three blurred scales, stored column-major (u16)
    constexpr int HP_ITERS = (9 * PH + THREADS - 1) / THREADS;
#pragma unroll
    for (int it = 0; it < HP_ITERS; ++it) {
        const int task = tid + it * THREADS;
        const int q = task / PH;
        const int p = task - q * PH;
        if (task >= 9 * PH || p >= p_need) continue;
        unsigned int w[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] = (q + k < PWW) ? gray[p * PWW + q + k] : 0u;
        hpass_scale<0>(w, coef, reinterpret_cast<unsigned short*>(sm + OFF_H0), p, q);
        hpass_scale<1>(w, coef, reinterpret_cast<unsigned short*>(sm + OFF_H1), p, q);
        hpass_scale<2>(w, coef, reinterpret_cast<unsigned short*>(sm + OFF_H2), p, q);
    }
    __syncthreads();

    // ---- pass 2: vertical taps + rounding -> blurred bytes, x = -1..32 at byte x+1, y = -1..32 at row y+1
    {
        // thread = (row group j0 + k * VP_STRIDE, column o): the division happens once per thread, not once per task
        constexpr int VP_STRIDE = THREADS / NCOL;                   // row groups covered per sweep (7 of 17)
        unsigned char* blur = reinterpret_cast<unsigned char*>(sm + OFF_B0);
        const int groups_need = (rows_valid + 2 + 3) / 4;           // row groups that hold a needed blurred row
        const int j0 = tid / NCOL;
        const int o = tid - j0 * NCOL;
        if (j0 < VP_STRIDE) {
            for (int j = j0; j < groups_need; j += VP_STRIDE) vpass_scale<2>(sm + OFF_H2, coef, blur + 2 * BLUR_WORDS * 4, j, o);
            for (int j = j0; j < groups_need; j += VP_STRIDE) vpass_scale<1>(sm + OFF_H1, coef, blur + BLUR_WORDS * 4, j, o);
            for (int j = j0; j < groups_need; j += VP_STRIDE) vpass_scale<0>(sm + OFF_H0, coef, blur, j, o);
        }
    }
    __syncthreads();

    // ---- Scharr on the four scales: thread = (row oy, 4 columns 4q..4q+3)
    const int q = tid & 7;
#pragma unroll 1
    for (int oy = tid >> 3; oy < BH; oy += THREADS / 8) {
    if (by + oy >= t.h) break;
    unsigned int s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;
    {
        // unblurred scale: gray byte of x is x + 8; columns 4q+e need bytes 4q+e+7 .. 4q+e+9
        unsigned int A[3], B[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const unsigned int* row = gray + (oy + HALO - 1 + d) * PWW + q + 1;
            const unsigned int w0 = row[0], w1 = row[1], w2 = row[2];
            A[d] = __funnelshift_r(w0, w1, 24);     // bytes 4q+7 .. 4q+10
            B[d] = __funnelshift_r(w1, w2, 8);      // bytes 4q+9 .. 4q+12
        }
        scharr2(A[0], A[1], A[2], s0, s1);
        scharr2(B[0], B[1], B[2], s2, s3);
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        // blurred scales: byte of x is x + 1; columns 4q+e need bytes 4q+e .. 4q+e+2
        const unsigned int* blur = sm + OFF_B0 + s * BLUR_WORDS;
        unsigned int A[3], B[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const unsigned int* row = blur + (oy + d) * (BLUR_PITCH / 4) + q;
            const unsigned int w0 = row[0], w1 = row[1];
            A[d] = w0;
            B[d] = __funnelshift_r(w0, w1, 16);
        }
        scharr2(A[0], A[1], A[2], s0, s1);
        scharr2(B[0], B[1], B[2], s2, s3);
    }
    const int y = by + oy;
    const int x = bx + 4 * q;
    if (y < t.h && x < t.w) {
        unsigned int* dst = S_out + t.px_off + (long long)y * t.w + x;
        if (x + 3 < t.w && ((reinterpret_cast<unsigned long long>(dst) & 15ULL) == 0ULL)) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(s0, s1, s2, s3);
        } else {
            dst[0] = s0;
            if (x + 1 < t.w) dst[1] = s1;
            if (x + 2 < t.w) dst[2] = s2;
            if (x + 3 < t.w) dst[3] = s3;
        }
    }
    }
}

}  // namespace gradfast
