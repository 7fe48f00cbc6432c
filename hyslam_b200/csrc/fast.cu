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
// Mapping: persistent CTAs (a few per SM) walk the list of 62x62 output tiles of the whole batch.  The 70x70 pixel
// region of a tile arrives in shared memory as ONE TMA box (cp.async.bulk.tensor; the box starts at the region's left
// edge rounded down to 16 bytes, as the TMA requires, and the hardware zero-fills past the image edge) into a double
// buffer: the box of the CTA's next tile is in flight while the current one is processed, so no thread spends
// instructions or scoreboard stalls on staging.  The corner test runs 4 pixels per thread on packed bytes (SWAR
// compares, LOP3 arc logic), instantiated for the four byte alignments the region can have inside the box; scores are
// then computed only for the compacted corner list with 3-input integer min/max (VIMNMX3); cell-local NMS and a
// CTA-aggregated emit follow.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "tma.cuh"

namespace hyorb {

// 4 bytes starting at byte O (0..11) of w0:w1:w2:w3
template <int O>
__device__ __forceinline__ uint32_t pick(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    if (O == 0) return w0;
    if (O < 4) return __funnelshift_r(w0, w1, 8 * O);
    if (O == 4) return w1;
    if (O < 8) return __funnelshift_r(w1, w2, 8 * (O & 3));
    if (O == 8) return w2;
    return __funnelshift_r(w2, w3, 8 * (O & 3));
}

// Bresenham circle of radius 3, OpenCV's order (features2d/fast_score.cpp makeOffsets); byte offset inside the staged tile
__device__ __forceinline__ constexpr int ring_off(int k)
{
    constexpr int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * FT_BOXW + dx[k];
}

constexpr int EMIT_CAP = 512 * FT_ITEMS;     // NMS survivors of one tile: < (FT_OW/2+1)*(FT_OH/2+1), doubled for cell edges
constexpr int RW = FT_BOXW / 4;   // shared-memory row stride in words = the TMA box width (dense box, no padding)
constexpr int PITCHB = FT_BOXW;   // row stride in bytes
constexpr int FT_BUF_BYTES = (FT_BOXW * FT_PH + 16 + 127) & ~127;   // one pixel buffer (+ one word of read slack), 128-byte aligned for TMA

struct FastTile { int b, l, sx0, sy0; };      // image, level, score-region origin

// tile id -> image, level, score-region origin
__device__ __forceinline__ FastTile fast_tile(const PlanDev *__restrict__ plan, int T)
{
    FastTile t;
    t.b = T / plan->tilesPerImage;
    const int ti = T - t.b * plan->tilesPerImage;
    int l = 0;
    const int nl = plan->nlevels;
    while (l + 1 < nl && ti >= plan->lv[l + 1].tileBase) l++;
    const LevelDev &L = plan->lv[l];
    const int k = ti - L.tileBase;
    const int tY = k / L.tilesX, tX = k - tY * L.tilesX;
    t.l = l;
    t.sx0 = DET_MIN + tX * FT_OW - 1;     // score region origin = first emitted pixel - 1
    t.sy0 = DET_MIN + tY * FT_OH - 1;
    return t;
}

// FAST-9/16 corner flags (bit 7 of each byte) of the 4 pixels of one work item.  `row` points at the word that holds the
// item's leftmost ring column (centre - 3) on the centre row; S = that column's byte position inside the word.
template <int S>
__device__ __forceinline__ uint32_t corner_flags(const uint32_t *__restrict__ row)
{
    uint32_t a0, a1, a2, a3 = 0;
#define LOADROW(dy) { a0 = row[(dy) * RW]; a1 = row[(dy) * RW + 1]; a2 = row[(dy) * RW + 2]; if (S + 6 > 8) a3 = row[(dy) * RW + 3]; }
    LOADROW(0)
    const uint32_t c = pick<S + 3>(a0, a1, a2, a3);
    const uint32_t hi = __vaddus4(c, 0x01010101u * FAST_T), lo = __vsubus4(c, 0x01010101u * FAST_T);
    uint32_t B[16], D[16];
    // per byte, bit 7 of ((a & ~b) | (~(a ^ b) & s)) is (a > b) when s = (a & 0x7f) + (~b & 0x7f); the two sums that involve
    // the ring pixel share its low 7 bits: bright s = r7 + Kb, dark s = Kd - r7 (no carry or borrow crosses a byte)
    const uint32_t Kb = ~hi & 0x7f7f7f7fu, Kd = (lo & 0x7f7f7f7fu) + 0x7f7f7f7fu;
#define RING(k, O) { const uint32_t rv = pick<S + O>(a0, a1, a2, a3); const uint32_t r7 = rv & 0x7f7f7f7fu; const uint32_t sb = r7 + Kb, sd = Kd - r7; \
                     B[k] = (rv & ~hi) | (~(rv ^ hi) & sb); D[k] = (lo & ~rv) | (~(lo ^ rv) & sd); }
    RING(12, 0) RING(4, 6)              // dy = 0 : dx = -3, +3
    LOADROW(1) RING(13, 0) RING(3, 6)
    LOADROW(2) RING(14, 1) RING(2, 5)
    LOADROW(3) RING(15, 2) RING(0, 3) RING(1, 4)       // dy = +3 : dx = -1, 0, +1
    LOADROW(-1) RING(11, 0) RING(5, 6)
    LOADROW(-2) RING(10, 1) RING(6, 5)
    LOADROW(-3) RING(9, 2) RING(8, 3) RING(7, 4)
#undef RING
#undef LOADROW
    // 9 contiguous: a3[k] = m[k]&m[k+1]&m[k+2]; a9[k] = a3[k]&a3[k+3]&a3[k+6]
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
    return ob | od;
}

// corner test of a thread's FT_ITEMS work items; pix = word that holds pixel-region column 0 of row 0
template <int S>
__device__ __forceinline__ void corner_test(const uint32_t *__restrict__ pix, int tid, int sx0, int sy0, int xEnd, int yEnd, uint32_t (&nflag)[FT_ITEMS])
{
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
        uint32_t f = 0;
        if (valid) f = corner_flags<S>(pix + (r + 3) * RW + g) & valid;
        nflag[it] = f;
    }
}

#ifndef HYORB_FAST_STOP
#define HYORB_FAST_STOP 0      // timing experiments only: 1..3 leave the tile loop after the test / compaction / scoring phase
#endif
#ifndef HYORB_FT_MINB
#define HYORB_FT_MINB 4
#endif
__global__ void __launch_bounds__(FT_THREADS, HYORB_FT_MINB)
k_fast(const PlanDev *__restrict__ plan, const __grid_constant__ CUtensorMap tm0, const CUtensorMap *__restrict__ tmaps, int img0, int nTiles,
       uint32_t *__restrict__ cand, int *__restrict__ candCount, int *__restrict__ status)
{
    __shared__ __align__(128) uint8_t s_pixbuf[2][FT_BUF_BYTES];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ __align__(4) uint8_t s_score[FT_SH * FT_SW];
    __shared__ uint16_t s_list[FT_SH * FT_SW];
    __shared__ uint32_t s_emit[EMIT_CAP];
    __shared__ uint8_t s_cf[FT_SW], s_rf[FT_SH];
    __shared__ int s_n, s_ne, s_base;

    const int tid = threadIdx.x;
    // pixel box of tile T -> buffer `buf`; issued by one thread, completion lands on s_bar[buf]
    auto issue = [&](int T, int buf) {
        const FastTile t = fast_tile(plan, T);
        mbar_arrive_expect_tx(&s_bar[buf], FT_BOXW * FT_PH);
        tma_load_3d(s_pixbuf[buf], t.l == 0 ? &tm0 : &tmaps[t.l], &s_bar[buf], (t.sx0 - 3) & ~15, t.sy0 - 3, img0 + t.b);
    };
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        fence_proxy_async();
        if ((int)blockIdx.x < nTiles) issue(blockIdx.x, 0);
    }
    __syncthreads();

    int it = 0;
    for (int T = blockIdx.x; T < nTiles; T += gridDim.x, it++) {
    const int buf = it & 1;
    // the other buffer was last read before the barrier that ended the previous iteration: refill it now
    if (tid == 0 && T + (int)gridDim.x < nTiles) { fence_proxy_async(); issue(T + gridDim.x, buf ^ 1); }
    const FastTile tile = fast_tile(plan, T);
    const int l = tile.l, b = tile.b;
    const LevelDev &L = plan->lv[l];
    const int sx0 = tile.sx0, sy0 = tile.sy0;
    const int off = (sx0 - 3) & 15;                         // byte position of pixel-region column 0 inside the box
    const uint8_t *pix8 = s_pixbuf[buf] + off;              // pixel region, row pitch PITCHB

    if (tid == 0) { s_n = 0; s_ne = 0; }
    for (int i = tid; i < FT_SH * FT_SW / 4; i += FT_THREADS) ((uint32_t *)s_score)[i] = 0;
    if (tid < FT_SW) {
        const int m = (sx0 + tid - DET_MIN) % L.wCell;      // sx0+tid >= 18; the halo column left of x=19 is never valid
        s_cf[tid] = (uint8_t)((m == 0 ? 1 : 0) | (m == L.wCell - 1 ? 2 : 0));
    } else if (tid >= 64 && tid < 64 + FT_SH) {      // FT_SH <= 128
        const int r = tid - 64;
        const int m = (sy0 + r - DET_MIN) % L.hCell;
        s_rf[r] = (uint8_t)((m == 0 ? 1 : 0) | (m == L.hCell - 1 ? 2 : 0));
    }
    mbar_wait(&s_bar[buf], (it >> 1) & 1);      // the pixel box has landed
    __syncthreads();

    const int xEnd = L.maxBX - 3, yEnd = L.maxBY - 3;     // detect range [19, xEnd) x [19, yEnd)
    // ---- corner test, 4 pixels per item; the byte alignment of the region inside the box is CTA-uniform
    uint32_t nflag[FT_ITEMS];
    {
        const uint32_t *pixw = (const uint32_t *)s_pixbuf[buf] + (off >> 2);
        switch (off & 3) {
        case 0: corner_test<0>(pixw, tid, sx0, sy0, xEnd, yEnd, nflag); break;
        case 1: corner_test<1>(pixw, tid, sx0, sy0, xEnd, yEnd, nflag); break;
        case 2: corner_test<2>(pixw, tid, sx0, sy0, xEnd, yEnd, nflag); break;
        default: corner_test<3>(pixw, tid, sx0, sy0, xEnd, yEnd, nflag); break;
        }
    }
#if HYORB_FAST_STOP == 1
    { uint32_t acc = 0;
#pragma unroll
      for (int k = 0; k < FT_ITEMS; k++) acc |= nflag[k];
      if (acc == 0x12345678u) atomicOr(status, 128); __syncthreads(); continue; }
#endif
    // ---- compact the corner flags into the CTA list.  A lane's FT_ITEMS corner counts (<= 4 each) ride in one register,
    // one byte per item: a single packed warp scan (no byte overflows: a warp holds <= 128 corners per item) gives every
    // lane its offsets for all items at once; one shared-memory atomic per warp reserves the chunk.
    {
        const int lane = tid & 31;
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < FT_ITEMS; k++) cnt |= (uint32_t)__popc(nflag[k] & 0x80808080u) << (8 * k);
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);      // per-item warp totals, one byte each
        const uint32_t excl = inc - cnt;
        int total = 0;
        int ibase[FT_ITEMS];
#pragma unroll
        for (int k = 0; k < FT_ITEMS; k++) { ibase[k] = total + (int)((excl >> (8 * k)) & 0xFFu); total += (int)((tot >> (8 * k)) & 0xFFu); }
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(&s_n, total);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int k = 0; k < FT_ITEMS; k++) {
            const uint32_t f = nflag[k];
            if (f) {
                const int id = tid + k * FT_THREADS;
                const int e0 = (id >> 4) * FT_SW + 4 * (id & 15);
                int pos = base + ibase[k];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((f >> (8 * j + 7)) & 1u) s_list[pos++] = (uint16_t)(e0 + j);
            }
        }
    }
    __syncthreads();

#if HYORB_FAST_STOP == 2
    __syncthreads(); continue;
#endif
    // ---- scores of the compacted corners
    const int ncorner = s_n;
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

#if HYORB_FAST_STOP == 3
    __syncthreads(); continue;
#endif
    // ---- cell-local 3x3 NMS over the interior, stage survivors
    for (int i0 = 0; i0 < ncorner; i0 += FT_THREADS) {      // warp-uniform trip count: the ballot below needs all lanes
        const int i = i0 + tid;
        bool kept = false;
        uint32_t packed = 0;
        // branch-free: the eight neighbour scores are fetched together (independent loads, every address stays inside
        // s_score for an interior pixel) and neighbours that belong to another cell are replaced by 0
        const int e = i < ncorner ? (int)s_list[i] : (FT_SW + 1);
        const int r = e / FT_SW, cidx = e - r * FT_SW;
        const bool inner = i < ncorner && r >= 1 && r <= FT_OH && cidx >= 1 && cidx <= FT_OW;      // else: halo, belongs to the neighbouring tile
        const uint8_t *q = s_score + (inner ? e : FT_SW + 1);
        const int s = q[0];
        const int nl = q[-1], nr = q[1], nu = q[-FT_SW], nul = q[-FT_SW - 1], nur = q[-FT_SW + 1], nd = q[FT_SW], ndl = q[FT_SW - 1], ndr = q[FT_SW + 1];
        const int cf = s_cf[cidx], rf = s_rf[r];
        const bool L_ok = !(cf & 1), R_ok = !(cf & 2), U_ok = !(rf & 1), D_ok = !(rf & 2);
        const int m0 = __vimax3_s32(L_ok ? nl : 0, R_ok ? nr : 0, U_ok ? nu : 0);
        const int m1 = __vimax3_s32(U_ok && L_ok ? nul : 0, U_ok && R_ok ? nur : 0, D_ok ? nd : 0);
        const int m2 = __vimax3_s32(D_ok && L_ok ? ndl : 0, D_ok && R_ok ? ndr : 0, m0);
        kept = inner && s > max(m1, m2);
        packed = pack_cand(sx0 + cidx - LATTICE_MIN, sy0 + r - LATTICE_MIN, s);
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
#if HYORB_FAST_STOP == 4
    if (s_ne == 0x7fffffff) atomicOr(status, 128);
    continue;
#endif
    int ne = s_ne;
    if (ne > 0) {       // CTA-uniform
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
    __syncthreads();      // every shared array (and the pixel buffer) is free for the next tile
    }
}

int launch_fast(const PlanDev &hp, const PlanDev *dp, const CUtensorMap &tm0, const CUtensorMap *tmaps, int img0, uint32_t *cand, int *candCount,
                int *status, int B, int sm_count, cudaStream_t st, long *launches)
{
    if (hp.tilesPerImage == 0) return HYORB_OK;
    static std::atomic<int> ctas_per_sm{0};          // a property of the compiled kernel, identical on every sm_100a device
    int per = ctas_per_sm.load(std::memory_order_relaxed);
    if (!per) {
        HY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_fast, FT_THREADS, 0));
        if (per < 1) per = 1;
        ctas_per_sm.store(per, std::memory_order_relaxed);
    }
    const int sms = sm_count > 0 ? sm_count : 148;
    const long long nTiles = (long long)hp.tilesPerImage * B;
    if (nTiles > 0x7fffffffLL) { set_error("too many FAST tiles in one batch"); return HYORB_EUNSUPPORTED; }
    const int grid = (int)std::min<long long>(nTiles, (long long)sms * per);
    k_fast<<<grid, FT_THREADS, 0, st>>>(dp, tm0, tmaps, img0, (int)nTiles, cand, candCount, status);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
