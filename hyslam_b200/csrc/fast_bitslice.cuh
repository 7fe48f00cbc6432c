// fast_bitslice.cuh -- bit-sliced FAST-9/16 corner test: 32 horizontally adjacent pixels per 32-bit word.
//
// The segment test of cv::FAST (called per cell by ORBFinder::detect, src/features/low_level/ORBFinder.cpp:66-68) compares
// 16 ring pixels with centre +- 20 and asks for 9 contiguous "brighter" or 9 contiguous "darker" ones.  On packed bytes
// (4 pixels per register) that is ~54 integer instructions per pixel, almost all on the half-rate ALU pipe.  Here the
// pixels of a tile are first transposed into 8 bit planes (bit j of plane word b = bit b of pixel j), after which
//   * an unsigned 8-bit compare is an 8-step ripple of ONE 3-input logic op per bit for 32 pixels at once,
//   * a ring offset dx is a funnel shift of the 8 plane words,
//   * the 9-of-16 arc logic works on 32 pixels per op instead of 4,
// about 19 instructions per pixel including the transposition.  Everything is exact integer logic, so the result is
// bit-identical to the byte formulation (checked against a scalar restatement on the CPU while developing, and by the
// GPU parity tests against the oracle).  Host+device so that the primitives can be exercised without a GPU.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define BS_HD __host__ __device__ __forceinline__
#else
#define BS_HD static inline
#endif

// 32 pixels (8 words, byte m of word j = pixel 4j+m) -> 8 bit planes (bit i of P[b] = bit b of pixel i)
BS_HD void bs_transpose(const uint32_t (&w)[8], uint32_t (&P)[8])
{
    uint32_t lo[4], hi[4];      // per 8-pixel group g: lo = planes 0..3 (one byte each), hi = planes 4..7
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t a = w[2 * g], b = w[2 * g + 1], t;
        // 8x8 bit-matrix transpose of the 8 bytes a:b (Hacker's Delight 7-3, split into 32-bit halves)
        t = (a ^ (a >> 7)) & 0x00AA00AAu; a ^= t ^ (t << 7);
        t = (b ^ (b >> 7)) & 0x00AA00AAu; b ^= t ^ (t << 7);
        t = (a ^ (a >> 14)) & 0x0000CCCCu; a ^= t ^ (t << 14);
        t = (b ^ (b >> 14)) & 0x0000CCCCu; b ^= t ^ (t << 14);
        t = ((a >> 4) ^ b) & 0x0F0F0F0Fu; b ^= t; a ^= t << 4;
        lo[g] = a; hi[g] = b;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
#ifdef __CUDA_ARCH__
        const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);      // byte k of each source
        P[k] = __byte_perm(__byte_perm(lo[0], lo[1], sel), __byte_perm(lo[2], lo[3], sel), 0x5410);
        P[4 + k] = __byte_perm(__byte_perm(hi[0], hi[1], sel), __byte_perm(hi[2], hi[3], sel), 0x5410);
#else
        P[k] = ((lo[0] >> (8 * k)) & 0xFFu) | (((lo[1] >> (8 * k)) & 0xFFu) << 8) | (((lo[2] >> (8 * k)) & 0xFFu) << 16) | (((lo[3] >> (8 * k)) & 0xFFu) << 24);
        P[4 + k] = ((hi[0] >> (8 * k)) & 0xFFu) | (((hi[1] >> (8 * k)) & 0xFFu) << 8) | (((hi[2] >> (8 * k)) & 0xFFu) << 16) | (((hi[3] >> (8 * k)) & 0xFFu) << 24);
#endif
    }
}

BS_HD uint32_t BS_FSR(uint32_t lo, uint32_t hi, int n)      // bits n .. n+31 of hi:lo
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, n);
#else
    return n ? (lo >> n) | (hi << (32 - n)) : lo;
#endif
}
BS_HD uint32_t BS_FSL(uint32_t lo, uint32_t hi, int n)      // bits 32-n .. 63-n of hi:lo
{
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, n);
#else
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
#endif
}

// saturating centre +- T (T = 20) on bit planes: hi = min(c + 20, 255), lo = max(c - 20, 0)
BS_HD void bs_thresholds(const uint32_t (&c)[8], uint32_t (&h)[8], uint32_t (&l)[8])
{
    // + 0b00010100
    uint32_t k;
    h[0] = c[0]; h[1] = c[1];
    h[2] = ~c[2]; k = c[2];
    h[3] = c[3] ^ k; k = c[3] & k;
    h[4] = ~(c[4] ^ k); k = c[4] | k;
    h[5] = c[5] ^ k; k = c[5] & k;
    h[6] = c[6] ^ k; k = c[6] & k;
    h[7] = c[7] ^ k; k = c[7] & k;           // k = overflow
#pragma unroll
    for (int b = 0; b < 8; b++) h[b] |= k;
    // - 0b00010100
    uint32_t q;
    l[0] = c[0]; l[1] = c[1];
    l[2] = ~c[2]; q = ~c[2];
    l[3] = c[3] ^ q; q = ~c[3] & q;
    l[4] = ~(c[4] ^ q); q = ~c[4] | q;
    l[5] = c[5] ^ q; q = ~c[5] & q;
    l[6] = c[6] ^ q; q = ~c[6] & q;
    l[7] = c[7] ^ q; q = ~c[7] & q;          // q = underflow
#pragma unroll
    for (int b = 0; b < 8; b++) l[b] &= ~q;
}

// ring value planes r vs thresholds: bright = r > h, dark = r < l (unsigned, LSB-first ripple: one 3-input op per bit)
BS_HD void bs_compare(const uint32_t (&r)[8], const uint32_t (&h)[8], const uint32_t (&l)[8], uint32_t &bright, uint32_t &dark)
{
    uint32_t g = 0, d = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        g = (r[b] & ~h[b]) | (~(r[b] ^ h[b]) & g);
        d = (~r[b] & l[b]) | (~(r[b] ^ l[b]) & d);
    }
    bright = g; dark = d;
}

// 9 contiguous of 16
BS_HD uint32_t bs_arc9(const uint32_t (&m)[16])
{
    uint32_t t3[16], t9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t3[k] = m[k] & m[(k + 1) & 15] & m[(k + 2) & 15];
#pragma unroll
    for (int k = 0; k < 16; k++) t9[k] = t3[k] & t3[(k + 3) & 15] & t3[(k + 6) & 15];
    uint32_t o = (t9[0] | t9[1] | t9[2]) | (t9[3] | t9[4] | t9[5]) | (t9[6] | t9[7] | t9[8]);
    o |= (t9[9] | t9[10] | t9[11]) | (t9[12] | t9[13] | t9[14]) | t9[15];
    return o;
}

// Corner word of one item: the 32 pixels whose plane words start at `rowp` (8 consecutive words = planes 0..7; the
// previous / next segment's words sit 8 words before / after; the rows above and below PLP words apart).  The caller's
// layout keeps an all-zero segment on either side of a row, so no access is conditional.
template <int PLP>
BS_HD uint32_t bs_corners(const uint32_t *rowp)
{
    uint32_t c[8], h[8], l[8];
#pragma unroll
    for (int b = 0; b < 8; b++) c[b] = rowp[b];
    bs_thresholds(c, h, l);
    uint32_t B[16], D[16];
    // ring position k at (dx, dy): bit j of the shifted planes = pixel j + dx of the row dy below
#define BS_RING(k, dx, dy) { \
        const uint32_t *p = rowp + (dy) * PLP; \
        uint32_t rv[8]; \
        if ((dx) > 0) { _Pragma("unroll") for (int b = 0; b < 8; b++) rv[b] = BS_FSR(p[b], p[8 + b], (dx)); } \
        else if ((dx) < 0) { _Pragma("unroll") for (int b = 0; b < 8; b++) rv[b] = BS_FSL(p[b - 8], p[b], -(dx)); } \
        else { _Pragma("unroll") for (int b = 0; b < 8; b++) rv[b] = p[b]; } \
        bs_compare(rv, h, l, B[k], D[k]); }
    BS_RING(0, 0, 3) BS_RING(1, 1, 3) BS_RING(2, 2, 2) BS_RING(3, 3, 1) BS_RING(4, 3, 0) BS_RING(5, 3, -1) BS_RING(6, 2, -2) BS_RING(7, 1, -3)
    BS_RING(8, 0, -3) BS_RING(9, -1, -3) BS_RING(10, -2, -2) BS_RING(11, -3, -1) BS_RING(12, -3, 0) BS_RING(13, -3, 1) BS_RING(14, -2, 2) BS_RING(15, -1, 3)
#undef BS_RING
    return bs_arc9(B) | bs_arc9(D);
}
