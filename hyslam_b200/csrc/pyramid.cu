// pyramid.cu -- K1: scale pyramid.  Replaces ORBExtractor::ComputePyramid (src/features/ORBExtractor.cpp:564-589):
// level l = cv::resize(level l-1, INTER_LINEAR) with OpenCV's 8-bit fixed-point bilinear kernel (Q11 coefficients,
// (b*(h>>4))>>16 vertical combine, (+2)>>2).  The 19-px reflected border the reference adds is never read
// downstream and is not materialised.
//
// Streaming kernel; one launch per level covers the whole batch.  A thread owns 4 adjacent destination columns and
// walks down RS_ROWS destination rows.  Source rows are 16-byte aligned (pyramid layout; level 0 is TMA-addressable by
// construction, api.cu): for each source row it needs, a thread loads three aligned 32-bit words (the <= 9 source bytes
// under its 4 columns), funnel-shifts them to the first tap and forms each horizontal interpolation with one PRMT (pick
// the two taps) + one DP2A (taps x Q11 coefficient pair); a source row's four horizontal values are reused by the next
// destination row when the vertical taps overlap (most rows at 1.2x).  The vertical combine (b*(h>>4))>>16 is one
// IMAD.HI per tap against the coefficient pre-shifted by 16.
#include "common.cuh"

namespace hyorb {

#ifndef HYORB_RS_ROWS
#define HYORB_RS_ROWS 8
#endif
constexpr int RS_THREADS = 256, RS_ROWS = HYORB_RS_ROWS;

struct HRow { uint32_t h[4]; };      // horizontal values, already >> 4
struct RowW { uint32_t A, B; };      // bytes s0 .. s0+7 of a source row

__device__ __forceinline__ RowW load_roww(const uint8_t *__restrict__ row, int s0, int roww, int rowpitch, unsigned sh)
{
    RowW r;
    // words may run past the image width into the row padding: those bytes only ever meet a zero coefficient (the table
    // clamps the last taps to src[w-1] with c1 = 0)
    if ((s0 & ~3) + 12 <= rowpitch) {
        const uint32_t *p = (const uint32_t *)(row + (s0 & ~3));
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        r.A = __funnelshift_r(w0, w1, sh);
        r.B = __funnelshift_r(w1, w2, sh);
    } else {
        r.A = r.B = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (s0 + j < roww) r.A |= (uint32_t)row[s0 + j] << (8 * j);
            if (s0 + 4 + j < roww) r.B |= (uint32_t)row[s0 + 4 + j] << (8 * j);
        }
    }
    return r;
}
__device__ __forceinline__ HRow hrow(const RowW &w, const uint32_t (&sel)[4], const uint32_t (&c01)[4])
{
    HRow r;
#pragma unroll
    for (int j = 0; j < 4; j++) r.h[j] = __dp2a_lo(c01[j], __byte_perm(w.A, w.B, sel[j]), 0u) >> 4;   // (src[s0]*c0 + src[s0+1]*c1) >> 4, non-negative
    return r;
}

#ifndef HYORB_RS_MINB
#define HYORB_RS_MINB 8      // 32 registers: full occupancy, which this latency-bound kernel needs (measured 0.48 -> 0.44 ms per 256 images)
#endif
__global__ void __launch_bounds__(RS_THREADS, HYORB_RS_MINB)
k_resize(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride, int sw, int sh,
         uint8_t *__restrict__ dst, int dpitch, unsigned long long dstride, int dw, int dh,
         const ResizeTab *__restrict__ tx, const ResizeTab *__restrict__ ty, int area2x, int nx, int ny, int total)
{
    // flat strip index -> (image, strip row, strip column), columns fastest: neighbouring threads read neighbouring words, and no
    // thread of the grid is idle whatever the level's shape (the narrow upper levels wasted a third of a 2-D grid's threads)
    grid_dependency_wait();      // launch_dependent (common.cuh): nothing of the source level is read before the previous launch is complete
    const int item = blockIdx.x * RS_THREADS + threadIdx.x;
    if (item >= total) return;
    const int per = nx * ny;
    const int b = item / per, rem = item - b * per;
    const int syi = rem / nx, sxi = rem - syi * nx;
    const int x4 = sxi * 4, y0 = syi * RS_ROWS;
    const uint8_t *s = src + (size_t)b * sstride;
    uint8_t *d = dst + (size_t)b * dstride + x4;
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
        const uint32_t dlt = (uint32_t)(t.ofs - s0);          // 0..6 for a downscale by <= 2; checked on the host
        sel[j] = dlt | ((dlt + 1) << 4);                      // PRMT: byte0 = tap0, byte1 = tap1 (byte 2,3 = A[0], unused)
        c01[j] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
    }
    const unsigned shf = (unsigned)(s0 & 3) * 8;
    const bool full = x4 + 3 < dw;
    int haveRow = -1;          // source row whose horizontal values sit in hb
    HRow ha, hb;
    hb.h[0] = hb.h[1] = hb.h[2] = hb.h[3] = 0;
    for (int y = y0; y < y1; y++) {
        const ResizeTab vy = ty[y];
        const int sy0 = min(max(vy.ofs, 0), sh - 1), sy1 = min(max(vy.ofs + 1, 0), sh - 1);
        // both source rows are requested before either is consumed
        RowW wa, wb;
        wa.A = wa.B = wb.A = wb.B = 0;
        if (sy0 != haveRow) wa = load_roww(s + (size_t)sy0 * spitch, s0, sw, spitch, shf);
        if (sy1 != sy0) wb = load_roww(s + (size_t)sy1 * spitch, s0, sw, spitch, shf);
        ha = sy0 != haveRow ? hrow(wa, sel, c01) : hb;
        hb = sy1 != sy0 ? hrow(wb, sel, c01) : ha;
        haveRow = sy1;
        // ((b0 * h0) >> 16) + ((b1 * h1) >> 16): high halves of products with the coefficients pre-shifted by 16 (0 <= b <= 2048)
        const uint32_t b0 = (uint32_t)vy.c0 << 16, b1 = (uint32_t)vy.c1 << 16;
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t v = (__umulhi(b0, ha.h[j]) + __umulhi(b1, hb.h[j]) + 2u) >> 2;     // <= 255: the coefficients sum to 2048 (+-1)
            out |= v << (8 * j);
        }
        uint8_t *o = d + (size_t)y * dpitch;
        if (full) *(uint32_t *)o = out;     // pitch and level offsets are multiples of 16
        else for (int j = 0; x4 + j < dw; j++) o[j] = (uint8_t)(out >> (8 * j));
    }
}

// one level (l >= 1) from level l-1
int launch_resize_level(const PlanDev &hp, int l, Level0 l0, uint8_t *pyr, const ResizeTab *tabs, int B, cudaStream_t st, long *launches)
{
    const LevelDev &S = hp.lv[l - 1], &D = hp.lv[l];
    const uint8_t *src = (l == 1) ? l0.base : pyr + S.off;
    const int spitch = (l == 1) ? l0.pitch : S.pitch;
    const unsigned long long sstride = (l == 1) ? l0.stride : hp.pyrStride;
    const int nx = (D.w + 3) / 4, ny = (D.h + RS_ROWS - 1) / RS_ROWS;
    const long long total = (long long)nx * ny * B;
    if (total > 0x7fffffffLL) { set_error("pyramid level too large for one launch"); return HYORB_EUNSUPPORTED; }
    // levels 2.. depend only on the launch before them (level 1 follows the caller's copies)
    const dim3 grd((unsigned)((total + RS_THREADS - 1) / RS_THREADS));
    if (l > 1)
        HY_CUDA(launch_dependent(k_resize, grd, dim3(RS_THREADS), 0, st, src, spitch, sstride, S.w, S.h, pyr + D.off, D.pitch, hp.pyrStride, D.w, D.h,
                                 tabs + D.rsX, tabs + D.rsY, D.area2x, nx, ny, (int)total));
    else
        k_resize<<<grd, RS_THREADS, 0, st>>>(src, spitch, sstride, S.w, S.h, pyr + D.off, D.pitch, hp.pyrStride, D.w, D.h, tabs + D.rsX, tabs + D.rsY, D.area2x,
                                             nx, ny, (int)total);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

int launch_pyramid(const PlanDev &hp, const PlanDev *, Level0 l0, uint8_t *pyr, const ResizeTab *tabs, int B, cudaStream_t st, long *launches)
{
    for (int l = 1; l < hp.nlevels; l++) HY_TRY(launch_resize_level(hp, l, l0, pyr, tabs, B, st, launches));
    return HYORB_OK;
}

}  // namespace hyorb
