// pyramid.cu -- K1: scale pyramid.  Replaces ORBExtractor::ComputePyramid (src/features/ORBExtractor.cpp:564-589):
// level l = cv::resize(level l-1, INTER_LINEAR) with OpenCV's 8-bit fixed-point bilinear kernel (Q11 coefficients,
// (b*(h>>4))>>16 vertical combine, (+2)>>2).  The 19-px reflected border the reference adds is never read
// downstream and is not materialised.
//
// HBM-bound streaming kernel; one launch per level covers the whole batch.  A thread owns 4 adjacent destination
// columns and walks down RS_ROWS destination rows.  For each source row it needs, it loads three aligned 32-bit
// words (the <= 9 source bytes under its 4 columns), funnel-shifts them to the first tap and forms each horizontal
// interpolation with one PRMT (pick the two taps) + one DP2A (taps x Q11 coefficient pair); a source row's four
// horizontal values are reused by the next destination row when the vertical taps overlap (most rows at 1.2x).
#include "common.cuh"

namespace hyorb {

constexpr int RS_BX = 64, RS_BY = 4, RS_ROWS = 8;

struct HRow { int h[4]; };

__device__ __forceinline__ HRow hrow(const uint8_t *__restrict__ row, int s0, int roww, bool wordsafe, const uint32_t (&sel)[4], const uint32_t (&c01)[4])
{
    // bytes s0 .. s0+5 of the row, as two registers A (s0..s0+3) and B (s0+4..s0+7)
    uint32_t A, B;
    if (wordsafe && s0 + 12 <= roww) {
        const unsigned mis = (unsigned)((uintptr_t)(row + s0) & 3);
        const uint32_t *p = (const uint32_t *)(row + s0 - mis);
        const unsigned sh = mis * 8;
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        A = __funnelshift_r(w0, w1, sh);
        B = __funnelshift_r(w1, w2, sh);
    } else {
        A = B = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (s0 + j < roww) A |= (uint32_t)row[s0 + j] << (8 * j);
            if (s0 + 4 + j < roww) B |= (uint32_t)row[s0 + 4 + j] << (8 * j);
        }
    }
    HRow r;
#pragma unroll
    for (int j = 0; j < 4; j++) r.h[j] = (int)__dp2a_lo(c01[j], __byte_perm(A, B, sel[j]), 0u);   // src[s0]*c0 + src[s0+1]*c1
    return r;
}

__global__ void __launch_bounds__(RS_BX * RS_BY)
k_resize(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride, int sw, int sh,
         uint8_t *__restrict__ dst, int dpitch, unsigned long long dstride, int dw, int dh,
         const ResizeTab *__restrict__ tx, const ResizeTab *__restrict__ ty, int area2x)
{
    const int x4 = (blockIdx.x * RS_BX + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * RS_BY + threadIdx.y) * RS_ROWS;
    if (x4 >= dw || y0 >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sstride;
    uint8_t *d = dst + (size_t)blockIdx.z * dstride + x4;
    const int y1 = min(y0 + RS_ROWS, dh);
    if (area2x) {
        // cv::resize silently takes the INTER_AREA 2x2 mean when the ratio is exactly 2
        for (int y = y0; y < y1; y++) {
            const uint8_t *r0 = s + (size_t)(2 * y) * spitch, *r1 = r0 + spitch;
            for (int j = 0; j < 4 && x4 + j < dw; j++) {
                const int x = x4 + j;
                d[(size_t)y * dpitch + j] = (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
            }
        }
        return;
    }
    // per-thread horizontal constants: first tap of column 0, tap selectors relative to it, Q11 coefficient pairs
    int s0 = 0;
    uint32_t sel[4], c01[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const ResizeTab t = tx[min(x4 + j, dw - 1)];
        if (j == 0) s0 = t.ofs;
        const uint32_t dlt = (uint32_t)(t.ofs - s0);          // 0..4 for a downscale by <= 1.33; checked on the host
        sel[j] = dlt | ((dlt + 1) << 4);                      // PRMT: byte0 = tap0, byte1 = tap1 (byte 2,3 = A[0], unused)
        c01[j] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
    }
    // aligned-down word loads start at most 3 bytes before a row's first tap: inside the previous row / image, except
    // for the very first bytes of an unaligned batch base
    const bool base_aligned = ((uintptr_t)src & 3) == 0;
    int haveRow = -1;          // source row whose horizontal values sit in hb
    HRow ha, hb;
    for (int y = y0; y < y1; y++) {
        const ResizeTab vy = ty[y];
        const int sy0 = min(max(vy.ofs, 0), sh - 1), sy1 = min(max(vy.ofs + 1, 0), sh - 1);
        const bool wordsafe = base_aligned || s0 >= 3 || sy0 > 0 || blockIdx.z > 0;
        if (sy0 == haveRow) ha = hb;
        else ha = hrow(s + (size_t)sy0 * spitch, s0, sw, wordsafe, sel, c01);
        if (sy1 == sy0) hb = ha;
        else hb = hrow(s + (size_t)sy1 * spitch, s0, sw, wordsafe, sel, c01);
        haveRow = sy1;
        const int b0 = vy.c0, b1 = vy.c1;
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int v = (((b0 * (ha.h[j] >> 4)) >> 16) + ((b1 * (hb.h[j] >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            out |= (uint32_t)v << (8 * j);
        }
        uint8_t *o = d + (size_t)y * dpitch;
        if (x4 + 3 < dw) *(uint32_t *)o = out;     // pitch and level offsets are multiples of 16
        else for (int j = 0; x4 + j < dw; j++) o[j] = (uint8_t)(out >> (8 * j));
    }
}

int launch_pyramid(const PlanDev &hp, const PlanDev *, Level0 l0, uint8_t *pyr, const ResizeTab *tabs, int B, cudaStream_t st, long *launches)
{
    for (int l = 1; l < hp.nlevels; l++) {
        const LevelDev &S = hp.lv[l - 1], &D = hp.lv[l];
        const uint8_t *src = (l == 1) ? l0.base : pyr + S.off;
        const int spitch = (l == 1) ? l0.pitch : S.pitch;
        const unsigned long long sstride = (l == 1) ? l0.stride : hp.pyrStride;
        dim3 blk(RS_BX, RS_BY), grd((D.w + 4 * RS_BX - 1) / (4 * RS_BX), (D.h + RS_BY * RS_ROWS - 1) / (RS_BY * RS_ROWS), B);
        k_resize<<<grd, blk, 0, st>>>(src, spitch, sstride, S.w, S.h, pyr + D.off, D.pitch, hp.pyrStride, D.w, D.h,
                                      tabs + D.rsX, tabs + D.rsY, D.area2x);
        ++*launches;
    }
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
