// fast.cu -- K2: FAST-9/16 detection with the reference's *cell-local* non-max suppression.
//
// Replaces the per-cell loop of ORBExtractor::ComputeKeyPointsOctTree (src/features/ORBExtractor.cpp:430-470), which
// calls ORBFinder::detect -> cv::FAST(roi, kps, 20, true) (src/features/low_level/ORBFinder.cpp:66-68) once per ~30 px
// cell.  Restated per pixel (DESIGN.md "FAST"):
//   * a pixel (x,y) is tested iff 19 <= x < w-19 and 19 <= y < h-19 (the cells' detect regions tile that range);
//   * score = max(20, max over the 16 contiguous 9-arcs of the arc's min (centre-ring) / min (ring-centre)) - 1,
//     corner iff that max exceeds 20 (OpenCV FAST_t<16> + cornerScore<16>);
//   * NMS: strictly greater than the 8 neighbours' scores, where neighbours that belong to ANOTHER cell count as 0
//     (each cell was a separate cv::FAST call).  Cell of a pixel: ((x-19)/wCell, (y-19)/hCell).
// Output: unordered candidate list per (image, level), packed lattice x | y<<12 | response<<24; the reference's
// generation order is recovered downstream from cand_order_key().
//
// Mapping: one CTA per 62x30 output tile; pixels staged once in shared memory (coalesced 32-bit loads); the
// corner test runs 4 pixels per thread on packed bytes (SWAR compares, LOP3 arc logic); scores are then computed
// only for the compacted corner list with 3-input integer min/max (VIMNMX3).
#include "common.cuh"

namespace hyorb {

__device__ __forceinline__ uint32_t gt_msb(uint32_t a, uint32_t b)   // per byte: bit 7 = (a > b), unsigned
{
    const uint32_t s = (a & 0x7f7f7f7fu) + (~b & 0x7f7f7f7fu);
    return (a & ~b) | (~(a ^ b) & s);
}

template <int O>
__device__ __forceinline__ uint32_t pick(uint32_t w0, uint32_t w1, uint32_t w2)   // 4 bytes starting at byte O of w0:w1:w2
{
    if (O == 0) return w0;
    if (O < 4) return __funnelshift_r(w0, w1, 8 * O);
    if (O == 4) return w1;
    return __funnelshift_r(w1, w2, 8 * (O - 4));
}

// Bresenham circle of radius 3, OpenCV's order (features2d/fast_score.cpp makeOffsets); byte offset inside the staged tile
__device__ __forceinline__ constexpr int ring_off(int k)
{
    constexpr int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * 192 + dx[k];
}

constexpr int EMIT_CAP = 512 * FT_ITEMS;     // NMS survivors of one tile: < (FT_OW/2+1)*(FT_OH/2+1), doubled for cell edges
constexpr int RW = 48;            // shared-memory row stride in words: == 16 (mod 32), so the two rows a warp touches per
                                  // load (16 groups x 2 rows) fall into disjoint banks
constexpr int LW = 18;            // words actually loaded per row (70 pixels + 2)
constexpr int PITCHB = RW * 4;    // row stride in bytes

__global__ void __launch_bounds__(FT_THREADS)
k_fast(const PlanDev *__restrict__ plan, Level0 l0, const uint8_t *__restrict__ pyr,
       uint32_t *__restrict__ cand, int *__restrict__ candCount, int *__restrict__ status)
{
    __shared__ uint32_t s_pix[FT_PH * RW];
    __shared__ __align__(4) uint8_t s_score[FT_SH * FT_SW];
    __shared__ uint16_t s_list[FT_SH * FT_SW];
    __shared__ uint32_t s_emit[EMIT_CAP];
    __shared__ uint8_t s_cf[FT_SW], s_rf[FT_SH];
    __shared__ int s_n, s_ne, s_base;

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    // which level does this tile belong to
    int l = 0;
    const int nl = plan->nlevels;
    while (l + 1 < nl && (int)blockIdx.x >= plan->lv[l + 1].tileBase) l++;
    const LevelDev &L = plan->lv[l];
    const int t = blockIdx.x - L.tileBase;
    const int tX = t % L.tilesX, tY = t / L.tilesX;
    const int tx0 = DET_MIN + tX * FT_OW, ty0 = DET_MIN + tY * FT_OH;   // first emitted pixel
    const int sx0 = tx0 - 1, sy0 = ty0 - 1;                             // score region origin
    const int gx0 = sx0 - 3, gy0 = sy0 - 3;                             // pixel region origin
    const int w = L.w, h = L.h;
    const uint8_t *img; int pitch;
    if (l == 0) { img = l0.base + (size_t)b * l0.stride; pitch = l0.pitch; }
    else { img = pyr + (size_t)b * plan->pyrStride + L.off; pitch = L.pitch; }

    if (tid == 0) { s_n = 0; s_ne = 0; }
    // ---- stage pixels: FT_PH rows x 18 words; warp `wy` takes rows wy, wy+8, ...; lanes 0..17 one word each: aligned
    // 32-bit global loads re-aligned with a funnel shift, all loads of a thread issued before the first store
    {
        const int lane = tid & 31, wy = tid >> 5;
        constexpr int NR = (FT_PH + 7) / 8;
        uint32_t v[NR];
        const int x = gx0 + 4 * lane;
#pragma unroll
        for (int k = 0; k < NR; k++) {
            const int rr = wy + 8 * k, y = gy0 + rr;
            v[k] = 0;
            if (lane < LW && rr < FT_PH && y < h && x < w) {
                const uint8_t *p = img + (size_t)y * pitch;
                const unsigned mis = (unsigned)((uintptr_t)(p + x) & 3);
                const int xa = x - (int)mis;                           // aligned-down start; gx0 >= 15 keeps xa >= 0
                if (xa + 7 < w) {
                    const uint32_t *q = (const uint32_t *)(p + xa);
                    v[k] = __funnelshift_r(__ldg(q), __ldg(q + 1), mis * 8);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (x + j < w) v[k] |= (uint32_t)p[x + j] << (8 * j);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NR; k++) {
            const int rr = wy + 8 * k;
            if (lane < LW && rr < FT_PH) s_pix[rr * RW + lane] = v[k];
        }
    }
    for (int i = tid; i < FT_SH * FT_SW / 4; i += FT_THREADS) ((uint32_t *)s_score)[i] = 0;
    if (tid < FT_SW) {
        const int m = (sx0 + tid - DET_MIN) % L.wCell;      // sx0+tid >= 18; the halo column left of x=19 is never valid
        s_cf[tid] = (uint8_t)((m == 0 ? 1 : 0) | (m == L.wCell - 1 ? 2 : 0));
    } else if (tid >= 64 && tid < 64 + FT_SH) {      // FT_SH <= 128
        const int r = tid - 64;
        const int m = (sy0 + r - DET_MIN) % L.hCell;
        s_rf[r] = (uint8_t)((m == 0 ? 1 : 0) | (m == L.hCell - 1 ? 2 : 0));
    }
    __syncthreads();

    const int xEnd = L.maxBX - 3, yEnd = L.maxBY - 3;     // detect range [19, xEnd) x [19, yEnd)
    // ---- corner test, 4 pixels per item
    uint32_t nflag[FT_ITEMS];
#pragma unroll
    for (int it = 0; it < FT_ITEMS; it++) nflag[it] = 0u;
#pragma unroll 1
    for (int it = 0; it < FT_ITEMS; it++) {
        const int id = tid + it * FT_THREADS;
        const int g = id & 15, r = id >> 4;
        const int sy = sy0 + r;
        uint32_t valid = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int sx = sx0 + 4 * g + j;
            if (sx >= DET_MIN && sx < xEnd) valid |= 0x80u << (8 * j);
        }
        if (sy < DET_MIN || sy >= yEnd) valid = 0;
        if (valid == 0) continue;       // nflag[it] stays 0
        const uint32_t *row = s_pix + (r + 3) * RW + g;
        uint32_t a0, a1, a2;
        a0 = row[0]; a1 = row[1]; a2 = row[2];
        const uint32_t c = pick<3>(a0, a1, a2);
        const uint32_t hi = __vaddus4(c, 0x01010101u * FAST_T), lo = __vsubus4(c, 0x01010101u * FAST_T);
        uint32_t B[16], D[16];
        // per byte, bit 7 of ((a & ~b) | (~(a ^ b) & s)) is (a > b) when s = (a & 0x7f) + (~b & 0x7f); the two sums that involve
        // the ring pixel share its low 7 bits: bright s = r7 + Kb, dark s = Kd - r7 (no carry or borrow crosses a byte)
        const uint32_t Kb = ~hi & 0x7f7f7f7fu, Kd = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
#define RING(k, O) { const uint32_t rv = pick<O>(a0, a1, a2); const uint32_t r7 = rv & 0x7f7f7f7fu; const uint32_t sb = r7 + Kb, sd = Kd - r7; \
                     B[k] = (rv & ~hi) | (~(rv ^ hi) & sb); D[k] = (lo & ~rv) | (~(lo ^ rv) & sd); }
        RING(12, 0) RING(4, 6)                                           // dy = 0 : dx = -3, +3
        a0 = row[RW]; a1 = row[RW + 1]; a2 = row[RW + 2];                // dy = +1
        RING(13, 0) RING(3, 6)
        a0 = row[2 * RW]; a1 = row[2 * RW + 1]; a2 = row[2 * RW + 2];    // dy = +2
        RING(14, 1) RING(2, 5)
        a0 = row[3 * RW]; a1 = row[3 * RW + 1];                          // dy = +3 : dx = -1, 0, +1
        RING(15, 2) RING(0, 3) RING(1, 4)
        a0 = row[-RW]; a1 = row[-RW + 1]; a2 = row[-RW + 2];             // dy = -1
        RING(11, 0) RING(5, 6)
        a0 = row[-2 * RW]; a1 = row[-2 * RW + 1]; a2 = row[-2 * RW + 2]; // dy = -2
        RING(10, 1) RING(6, 5)
        a0 = row[-3 * RW]; a1 = row[-3 * RW + 1];                        // dy = -3
        RING(9, 2) RING(8, 3) RING(7, 4)
#undef RING
        // 9 contiguous: a3[k] = m[k]&m[k+1]&m[k+2]; a9[k] = a3[k]&a3[k+3]&a3[k+6]
        uint32_t any;
        {
            uint32_t t3[16], t9[16];
#pragma unroll
            for (int k = 0; k < 16; k++) t3[k] = B[k] & B[(k + 1) & 15] & B[(k + 2) & 15];
#pragma unroll
            for (int k = 0; k < 16; k++) t9[k] = t3[k] & t3[(k + 3) & 15] & t3[(k + 6) & 15];
            uint32_t ob = (t9[0] | t9[1] | t9[2]) | (t9[3] | t9[4] | t9[5]) | (t9[6] | t9[7] | t9[8]);
            ob |= (t9[9] | t9[10] | t9[11]) | (t9[12] | t9[13] | t9[14]) | t9[15];
#pragma unroll
            for (int k = 0; k < 16; k++) t3[k] = D[k] & D[(k + 1) & 15] & D[(k + 2) & 15];
#pragma unroll
            for (int k = 0; k < 16; k++) t9[k] = t3[k] & t3[(k + 3) & 15] & t3[(k + 6) & 15];
            uint32_t od = (t9[0] | t9[1] | t9[2]) | (t9[3] | t9[4] | t9[5]) | (t9[6] | t9[7] | t9[8]);
            od |= (t9[9] | t9[10] | t9[11]) | (t9[12] | t9[13] | t9[14]) | t9[15];
            any = ob | od;
        }
        any &= valid;
        nflag[it] = any;
    }
    // ---- compact the corner flags into the CTA list: one ballot per (item, pixel) slot gives every lane its offset inside
    // the warp's contiguous chunk; one shared-memory atomic per warp reserves the chunk
    {
        const int lane = tid & 31;
        const unsigned lt = (1u << lane) - 1u;
        unsigned bal[FT_ITEMS * 4];
        int total = 0;
#pragma unroll
        for (int it = 0; it < FT_ITEMS; it++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                bal[it * 4 + j] = __ballot_sync(0xffffffffu, (nflag[it] >> (8 * j + 7)) & 1u);
                total += __popc(bal[it * 4 + j]);
            }
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(&s_n, total);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int it = 0; it < FT_ITEMS; it++) {
            const int id = tid + it * FT_THREADS;
            const int e0 = (id >> 4) * FT_SW + 4 * (id & 15);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned bj = bal[it * 4 + j];
                if (bj & (1u << lane)) s_list[base + __popc(bj & lt)] = (uint16_t)(e0 + j);
                base += __popc(bj);
            }
        }
    }
    __syncthreads();

    // ---- scores of the compacted corners
    const int ncorner = s_n;
    const uint8_t *pix8 = (const uint8_t *)s_pix;
    for (int i = tid; i < ncorner; i += FT_THREADS) {
        const int e = s_list[i];
        const int r = e / FT_SW, cidx = e - r * FT_SW;
        const uint8_t *p = pix8 + (r + 3) * PITCHB + (cidx + 3);
        const int v = p[0];
        // 16-bit lanes: low = centre - ring (dark arcs), high = ring - centre (bright arcs).  With R' = (ring+1)*65535 =
        // (ring << 16 | -(ring+1)) and A' = (-centre << 16 | centre+1), the lane-wise sum A' + R' is exactly that pair.
        const uint32_t A = ((uint32_t)(v + 1) & 0xFFFFu) | ((uint32_t)(-v) << 16);
        uint32_t wv[16];
#pragma unroll
        for (int k = 0; k < 16; k++) wv[k] = __vadd2(A, ((uint32_t)p[ring_off(k)] + 1u) * 65535u);
        uint32_t m3[16];
#pragma unroll
        for (int k = 0; k < 16; k++) m3[k] = __vimin3_s16x2(wv[k], wv[(k + 1) & 15], wv[(k + 2) & 15]);
        uint32_t m9[16];
#pragma unroll
        for (int k = 0; k < 16; k++) m9[k] = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);   // min over the 9-arc starting at k
        uint32_t mx = __vimax3_s16x2(m9[0], m9[1], m9[2]);
        mx = __vimax3_s16x2(mx, m9[3], m9[4]); mx = __vimax3_s16x2(mx, m9[5], m9[6]); mx = __vimax3_s16x2(mx, m9[7], m9[8]);
        mx = __vimax3_s16x2(mx, m9[9], m9[10]); mx = __vimax3_s16x2(mx, m9[11], m9[12]); mx = __vimax3_s16x2(mx, m9[13], m9[14]);
        mx = __vimax3_s16x2(mx, m9[15], m9[15]);
        const int sd = (int)(short)(mx & 0xFFFFu), sb = (int)mx >> 16;
        const int sc = __vimax3_s32(sd, sb, FAST_T) - 1;
        s_score[e] = (uint8_t)sc;
    }
    __syncthreads();

    // ---- cell-local 3x3 NMS over the interior, stage survivors
    for (int i0 = 0; i0 < ncorner; i0 += FT_THREADS) {      // warp-uniform trip count: the ballot below needs all lanes
        const int i = i0 + tid;
        bool kept = false;
        uint32_t packed = 0;
        do {
        if (i >= ncorner) break;
        const int e = s_list[i];
        const int r = e / FT_SW, cidx = e - r * FT_SW;
        if (r < 1 || r > FT_OH || cidx < 1 || cidx > FT_OW) break;      // halo: belongs to the neighbouring tile
        const int s = s_score[e];
        const int cf = s_cf[cidx], rf = s_rf[r];
        const bool L_ok = !(cf & 1), R_ok = !(cf & 2), U_ok = !(rf & 1), D_ok = !(rf & 2);
        const uint8_t *q = s_score + e;
        bool keep = true;
        if (L_ok) keep = keep && s > q[-1];
        if (R_ok) keep = keep && s > q[1];
        if (U_ok) {
            keep = keep && s > q[-FT_SW];
            if (L_ok) keep = keep && s > q[-FT_SW - 1];
            if (R_ok) keep = keep && s > q[-FT_SW + 1];
        }
        if (D_ok) {
            keep = keep && s > q[FT_SW];
            if (L_ok) keep = keep && s > q[FT_SW - 1];
            if (R_ok) keep = keep && s > q[FT_SW + 1];
        }
        kept = keep;
        packed = pack_cand(sx0 + cidx - LATTICE_MIN, sy0 + r - LATTICE_MIN, s);
        } while (0);
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (bal) {
            const int lane = tid & 31;
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_ne, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int pos = base + __popc(bal & ((1u << lane) - 1));
            if (kept && pos < EMIT_CAP) s_emit[pos] = packed;
        }
    }
    __syncthreads();
    int ne = s_ne;
    if (ne == 0) return;
    int *cnt = candCount + b * HYORB_MAX_LEVELS + l;
    if (tid == 0) {
        if (ne > EMIT_CAP) { atomicOr(status, ST_CAND_OVERFLOW); }
        s_base = atomicAdd(cnt, min(ne, EMIT_CAP));
    }
    __syncthreads();
    ne = min(ne, EMIT_CAP);
    const int base = s_base;
    uint32_t *out = cand + (size_t)b * plan->candStride + L.candOff;
    for (int i = tid; i < ne; i += FT_THREADS) {
        if (base + i < L.candCap) out[base + i] = s_emit[i];
        else atomicOr(status, ST_CAND_OVERFLOW);
    }
}

int launch_fast(const PlanDev &hp, const PlanDev *dp, Level0 l0, const uint8_t *pyr, uint32_t *cand, int *candCount, int *status,
                int B, cudaStream_t st, long *launches)
{
    if (hp.tilesPerImage == 0) return HYORB_OK;
    dim3 grd(hp.tilesPerImage, B);
    k_fast<<<grd, FT_THREADS, 0, st>>>(dp, l0, pyr, cand, candCount, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
