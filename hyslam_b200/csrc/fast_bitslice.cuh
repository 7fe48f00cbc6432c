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

// ---------------------------------------------------------------------------------------------------------------------------
// Paired formulation: every ordered comparison of the segment test is shared by TWO pixels.  With o_k the ring offset of
// position k (k = 0..7: the half of the ring with dx >= 0) define at every pixel p
//     Gp_k(p) = [I(p + o_k) > I(p) + 20]  = bright_k(p)          Gm_k(p) = [I(p + o_k) < I(p) - 20]  = dark_k(p)
// then, because a > b + 20 <=> b < a - 20 on integers (the saturations at 255 / 0 leave both sides false together),
//     bright_{k+8}(p) = [I(p - o_k) > I(p) + 20] = Gm_k(p - o_k)   dark_{k+8}(p) = [I(p - o_k) < I(p) - 20] = Gp_k(p - o_k)
// so 16 ripple compares per 32 pixels instead of 32 and 7 plane shifts instead of 14; the other half of the ring is read back from
// the pair words of the pixel rows r - dy_k, shifted by dx_k (exchange through a buffer every item writes its 16 words to).
// bs_pairs: the 16 pair words of one item; bs_corners_paired: corner word from the item's own pair words and the buffer.
// Buffer layout: word (2k + t) of item (row, seg) at g[row * GP + seg * 16 + 2k + t], t = 0: Gp_k, t = 1: Gm_k; the 16 words before an
// item belong to the segment on its left (for segment 0: whatever precedes -- only bits of columns < 3 come from there, never valid).
template <int PLP>
BS_HD void bs_pairs(const uint32_t *rowp, uint32_t (&Gp)[8], uint32_t (&Gm)[8])
{
    uint32_t c[8], h[8], l[8];
#pragma unroll
    for (int b = 0; b < 8; b++) c[b] = rowp[b];
    bs_thresholds(c, h, l);
#define BS_PAIR(k, dx, dy) { \
        const uint32_t *p = rowp + (dy) * PLP; \
        uint32_t rv[8]; \
        if ((dx) > 0) { _Pragma("unroll") for (int b = 0; b < 8; b++) rv[b] = BS_FSR(p[b], p[8 + b], (dx)); } \
        else { _Pragma("unroll") for (int b = 0; b < 8; b++) rv[b] = p[b]; } \
        bs_compare(rv, h, l, Gp[k], Gm[k]); }
    BS_PAIR(0, 0, 3) BS_PAIR(1, 1, 3) BS_PAIR(2, 2, 2) BS_PAIR(3, 3, 1) BS_PAIR(4, 3, 0) BS_PAIR(5, 3, -1) BS_PAIR(6, 2, -2) BS_PAIR(7, 1, -3)
#undef BS_PAIR
}

template <int GP>
BS_HD uint32_t bs_corners_paired(const uint32_t (&Gp)[8], const uint32_t (&Gm)[8], const uint32_t *g)
{
    uint32_t B[16], D[16];
#pragma unroll
    for (int k = 0; k < 8; k++) { B[k] = Gp[k]; D[k] = Gm[k]; }
    // position k + 8 of pixel p = position k of pixel p - o_k: row - dy_k, column - dx_k (bit j of the result = bit j - dx of the row's words)
#define BS_BACK(k, dx, dy) { \
        const uint32_t *q = g - (dy) * GP + 2 * (k); \
        if ((dx) > 0) { B[(k) + 8] = BS_FSL(q[-15], q[1], (dx)); D[(k) + 8] = BS_FSL(q[-16], q[0], (dx)); } \
        else { B[(k) + 8] = q[1]; D[(k) + 8] = q[0]; } }
    BS_BACK(0, 0, 3) BS_BACK(1, 1, 3) BS_BACK(2, 2, 2) BS_BACK(3, 3, 1) BS_BACK(4, 3, 0) BS_BACK(5, 3, -1) BS_BACK(6, 2, -2) BS_BACK(7, 1, -3)
#undef BS_BACK
    return bs_arc9(B) | bs_arc9(D);
}

