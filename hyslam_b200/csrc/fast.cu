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
// Mapping: persistent CTAs (a few per SM) walk the list of 88x56 output tiles of the whole batch.  The 96x64 pixel
// region of a tile arrives in shared memory as ONE TMA box (cp.async.bulk.tensor; the box starts at the region's left
// edge rounded down to 16 bytes, as the TMA requires, and the hardware zero-fills past the image edge) into a double
// buffer: the box of the CTA's next tile is in flight while the current one is processed, so no thread spends
// instructions or scoreboard stalls on staging.  Per tile:
//   T. every thread transposes 32 pixels into 8 bit-plane words (fast_bitslice.cuh);
//   A. every thread runs the bit-sliced segment test on 32 pixels in the PAIRED formulation: one 3-input logic op per bit for the
//      half ring with dx >= 0 (funnel shifts for the x offsets), the 16 pair words published to shared memory, the other half ring
//      read back from the items three rows above / below; 9-of-16 arc logic on 32 pixels per op;
//   B. corner bits are compacted into the CTA's list (warp scan + one shared-memory atomic per warp; rows interleaved two apart
//      so that the byte gathers of C and D spread over the banks);
//   C. scores are computed only for the listed corners with 3-input integer min/max (VIMNMX3);
//   D. cell-local 3x3 NMS (branch-free); every warp stages its survivors and appends them to the level's list with one atomic.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "tma.cuh"
#include "fast_bitslice.cuh"

namespace hyorb {

// Bresenham circle of radius 3, OpenCV's order (features2d/fast_score.cpp makeOffsets); byte offset inside the staged tile
__device__ __forceinline__ constexpr int ring_off(int k)
{
    constexpr int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    return dy[k] * FT_BOXW + dx[k];
}

constexpr int EMIT_W = 512;                    // staging entries per warp: a tile has at most (FT_OW/2+1)*(FT_OH/2+1) = 1305 NMS survivors plus adjacent pairs across
                                               // cell edges, spread over the six warps by list position; a warp that still exceeded its segment is reported
constexpr int PITCHB = FT_BOXW;                // box row stride in bytes
constexpr int FT_BUF_BYTES = (FT_BOXW * FT_PH + 127) & ~127;   // one pixel buffer, 128-byte aligned for TMA
constexpr int FT_PLP = 8 * (FT_NSEG + 2) + 4;  // plane row pitch in words: an all-zero segment either side, +4 so that the 128-bit
                                               // loads of 8 consecutive rows hit disjoint banks
static_assert(EMIT_W * (FT_THREADS / 32) <= (FT_PH + 6) * FT_PLP, "the emit staging aliases the plane buffer (slack rows included)");
constexpr int FT_SP = FT_BOXW;                 // score-map row pitch = pixel-box row pitch (28 words): a list entry indexes both, and the rows r, r+2, r+4, r+6 a
                                               // warp's 32 list entries come from (see phase B) start 8 banks apart -- their byte gathers do not collide
constexpr int FT_GP = 16 * FT_NSEG + 2;        // pair-word buffer: row pitch in words (+2: the 64-bit accesses of 16 consecutive rows hit disjoint banks)

struct FastTile { int b, l, x0, sy0; };      // image, level, image column of region column 0, image row of score row 0

// tile id -> image, level, origin
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
    t.x0 = DET_MIN - FT_C0 + tX * FT_OW;      // first emitted column (region column FT_C0) of tile 0 is x = 19
    t.sy0 = DET_MIN - 1 + tY * FT_OH;         // first emitted row (score row 1) of tile 0 is y = 19
    return t;
}

// bits [a, b) of a 32-bit word (any integers a, b)
__device__ __forceinline__ uint32_t bit_range(int a, int b)
{
    const uint32_t hi = b >= 32 ? 0xFFFFFFFFu : (b <= 0 ? 0u : (1u << b) - 1u);
    const uint32_t lo = a >= 32 ? 0xFFFFFFFFu : (a <= 0 ? 0u : (1u << a) - 1u);
    return hi & ~lo;
}

#ifndef HYORB_FT_MINB
#define HYORB_FT_MINB 5      // 64 registers: five 192-thread CTAs per SM (measured faster than 4 x 80 or 3 x 86 registers)
#endif
__global__ void __launch_bounds__(FT_THREADS, HYORB_FT_MINB)
k_fast(const PlanDev *__restrict__ plan, const __grid_constant__ CUtensorMap tm0, const CUtensorMap *__restrict__ tmaps, int img0, int nTiles,
       uint32_t *__restrict__ cand, int *__restrict__ candCount, int *__restrict__ status)
{
    __shared__ __align__(128) uint8_t s_pixbuf[2][FT_BUF_BYTES];
    __shared__ __align__(16) uint32_t s_planes_all[(FT_PH + 6) * FT_PLP];   // three rows of slack above and below: the pair words of the first / last
                                                                           // three pixel rows look outside the box (their results are never used)
    // score map + corner list; during the corner test the same bytes hold the pair words of every (pixel row, segment) item (fast_bitslice.cuh)
    __shared__ __align__(8) uint8_t s_sl[FT_SH * FT_SP + 2 * FT_SH * FT_SW];
    static_assert(sizeof(uint32_t) * (16 + FT_PH * FT_GP) <= FT_SH * FT_SP + 2 * FT_SH * FT_SW, "the pair words alias score map + list");
    static_assert((FT_SH * FT_SP) % 4 == 0 && FT_SH * FT_SP < 65536, "list alignment / 16-bit list entries");
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ __align__(16) FastTile s_tile[2];
    __shared__ uint8_t s_msk[2 * FT_SW + 128];                          // 1 = the left / right / upper / lower neighbour is in the same cell
    uint8_t *const s_cl = s_msk, *const s_cr = s_msk + FT_SW, *const s_ru = s_msk + 2 * FT_SW, *const s_rd = s_msk + 2 * FT_SW + 64;
    __shared__ int s_n;
    uint32_t *const s_planes = s_planes_all + 3 * FT_PLP;
    uint32_t *s_emit = s_planes_all;        // the planes are dead once the corner test is done
    uint8_t *const s_score = s_sl;
    uint16_t *const s_list = (uint16_t *)(s_sl + FT_SH * FT_SP);
    uint32_t *const s_g = (uint32_t *)s_sl + 16;

    grid_dependency_wait();      // launch_dependent (common.cuh): follows the last pyramid level
    const int tid = threadIdx.x;
    // pixel box of tile T -> buffer `buf`; issued by one thread, completion lands on s_bar[buf]
    auto issue = [&](int T, int buf) {
        const FastTile t = fast_tile(plan, T);
        s_tile[buf] = t;         // decoded once by the issuing thread; everybody reads it after the barrier that ends the previous tile
        mbar_arrive_expect_tx(&s_bar[buf], FT_BOXW * FT_PH);
        tma_load_3d(s_pixbuf[buf], t.l == 0 ? &tm0 : &tmaps[t.l], &s_bar[buf], t.x0 & ~15, t.sy0 - 3, img0 + t.b);
    };
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        fence_proxy_async();
#ifndef HYORB_NO_TMAP_FENCE
        for (int l = 1; l < plan->nlevels; l++) tensormap_acquire(&tmaps[l]);
#endif
        if ((int)blockIdx.x < nTiles) issue(blockIdx.x, 0);
    }
    __syncthreads();

    int it = 0;
    for (int T = blockIdx.x; T < nTiles; T += gridDim.x, it++) {
    const int buf = it & 1;
    // the other buffer was last read before the barrier that ended the previous iteration: refill it now
    if (tid == 0 && T + (int)gridDim.x < nTiles) { fence_proxy_async(); issue(T + gridDim.x, buf ^ 1); }
    const FastTile tile = s_tile[buf];
    const int l = tile.l, b = tile.b;
    const LevelDev &L = plan->lv[l];
    const int x0 = tile.x0, sy0 = tile.sy0;
    const int off = x0 & 15;                                // byte position of region column 0 inside the box
    const uint8_t *pix8 = s_pixbuf[buf] + off;              // region, row pitch PITCHB

    if (tid == 0) s_n = 0;
    if (tid < FT_SW) {
        const int x = x0 + tid;                              // columns left of x = 19 are never emitted
        const int m = x >= DET_MIN ? (x - DET_MIN) % L.wCell : 1;
        s_cl[tid] = (uint8_t)(m != 0);
        s_cr[tid] = (uint8_t)(m != L.wCell - 1);
    } else if (tid >= 128 && tid < 128 + FT_SH) {
        const int r = tid - 128;
        const int y = sy0 + r;
        const int m = y >= DET_MIN ? (y - DET_MIN) % L.hCell : 1;
        s_ru[r] = (uint8_t)(m != 0);
        s_rd[r] = (uint8_t)(m != L.hCell - 1);
    }
    mbar_wait(&s_bar[buf], (it >> 1) & 1);      // the pixel box has landed

    const int xEnd = L.maxBX - 3, yEnd = L.maxBY - 3;     // detect range [19, xEnd) x [19, yEnd)
    // Partial tiles (bottom / right edge of a level): score rows [rLo, rHi) and region columns [cl, ch) hold detectable pixels.  Pair
    // words are needed for pixel rows rLo .. rHi+5 (score row r = pixel row r+3 reads the items of pixel rows r .. r+6) and columns
    // cl-3 .. ch+2, bit planes three rows further: everything else is skipped (whole warps on the short tiles of a level's last row).
    const int rLo = max(0, DET_MIN - sy0), rHi = min(FT_SH, yEnd - sy0);
    const int cl = max(3, DET_MIN - x0), ch = min(FT_SW - 3, xEnd - x0);
    // ---- T. bit-plane transposition: thread -> (region row, segment)
    if (const int row = tid / FT_NSEG; row >= rLo - 3 && row <= rHi + 8) {
        const int seg = tid - row * FT_NSEG;
        const int o = off + 32 * seg;
        const uint32_t *src = (const uint32_t *)(s_pixbuf[buf] + row * PITCHB) + (o >> 2);
        const unsigned sh = (unsigned)(o & 3) * 8;
        uint32_t x[9], w[8], P[8];
#pragma unroll
        for (int j = 0; j < 9; j++) x[j] = src[j];
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = __funnelshift_r(x[j], x[j + 1], sh);
        bs_transpose(w, P);
        uint4 *dst = (uint4 *)(s_planes + row * FT_PLP + (seg + 1) * 8);
        dst[0] = make_uint4(P[0], P[1], P[2], P[3]);
        dst[1] = make_uint4(P[4], P[5], P[6], P[7]);
        // the zero segments either side of the row (the emit staging of the previous tile may have overwritten them)
        if (seg == 0) { dst[-2] = make_uint4(0, 0, 0, 0); dst[-1] = make_uint4(0, 0, 0, 0); }
        if (seg == FT_NSEG - 1) { dst[2] = make_uint4(0, 0, 0, 0); dst[3] = make_uint4(0, 0, 0, 0); }
    }
    __syncthreads();

    // ---- A. corner test in the paired formulation (fast_bitslice.cuh): warp -> (segment, block of 32 PIXEL rows), lane -> pixel row.
    // A1: every item compares the half ring with dx >= 0 against centre +- 20 (16 ripple compares) and publishes its 16 pair words;
    // A2: the other half of the ring is read back from the items three rows above / below, shifted by the ring's x offsets.
    uint32_t flags = 0;
    const int a_seg = (tid >> 5) % FT_NSEG, a_pr = ((tid >> 5) / FT_NSEG) * 32 + (tid & 31);
    const int a_r = a_pr - 3;                              // score row of this item (valid: 0 .. FT_SH-1)
    uint32_t Gp[8], Gm[8];
    if (a_pr >= rLo && a_pr <= rHi + 5 && 32 * a_seg + 31 >= cl - 3 && 32 * a_seg < ch + 3) {
        bs_pairs<FT_PLP>(s_planes + a_pr * FT_PLP + (a_seg + 1) * 8, Gp, Gm);
        uint2 *gdst = (uint2 *)(s_g + a_pr * FT_GP + a_seg * 16);
#pragma unroll
        for (int k = 0; k < 8; k++) gdst[k] = make_uint2(Gp[k], Gm[k]);
    }
    __syncthreads();
    if (a_r >= 0 && a_r < FT_SH) {
        const int sy = sy0 + a_r;
        // region columns with a full ring [3, 93) that lie in the detect range, restricted to this segment
        uint32_t valid = bit_range(cl - 32 * a_seg, ch - 32 * a_seg);
        if (sy < DET_MIN || sy >= yEnd) valid = 0;
        if (valid) flags = bs_corners_paired<FT_GP>(Gp, Gm, s_g + a_pr * FT_GP + a_seg * 16) & valid;
    }
    __syncthreads();      // the pair words are dead: their bytes become the score map (zeroed here) and the corner list
    for (int i = tid; i < FT_SH * FT_SP / 4; i += FT_THREADS) ((uint32_t *)s_score)[i] = 0;
    // ---- B. compact the corner bits into the CTA list: warp scan of the per-lane counts, one shared-memory atomic per warp
    {
        // the lanes take the warp's 32 rows in the order 0, 2, 4, .., 30, 1, 3, .., 31: consecutive list entries (= the lanes of a warp in the
        // phases below) then come from rows two apart, whose bytes lie 8 banks apart in the pixel box and the score map
        const int lane = tid & 31, srcl = ((lane & 15) << 1) | (lane >> 4);
        flags = __shfl_sync(0xffffffffu, flags, srcl);
        const int l_r = a_r - lane + srcl;
        const int cnt = __popc(flags);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        int base = 0;
        if (lane == 0 && total) base = atomicAdd(&s_n, total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int pos = base + inc - cnt;
        const int e0 = l_r * FT_SP + 32 * a_seg;
        // branch-free extraction: every lane walks all 32 bit positions with predicated stores (a data-dependent loop runs at the
        // pace of the warp's busiest lane and costs several times more instructions)
        uint16_t *dstp = s_list + pos;
#pragma unroll
        for (int j = 0; j < 32; j++)
            if ((flags >> j) & 1u) *dstp++ = (uint16_t)(e0 + j);
    }
    __syncthreads();

    // ---- C. scores of the compacted corners
    const int ncorner = s_n;
    for (int i = tid; i < ncorner; i += FT_THREADS) {
        const int e = s_list[i];
        static_assert(FT_SP == PITCHB, "a list entry indexes the pixel box and the score map alike");
        const uint8_t *p = pix8 + 3 * PITCHB + e;
        const int v = p[0];
        // 16-bit lanes: low = centre - ring + 0x4000 (dark arcs, biased), high = ring - centre (bright arcs), as ONE multiply-add
        // per ring pixel (FMA pipe; the ALU pipe is this kernel's bottleneck): (ring - centre) * 65535 + 0x4000 =
        // (ring - centre) << 16 | (0x4000 + centre - ring) -- the bias keeps the low lane positive, so nothing carries or borrows
        // across the lanes; min / max are order-preserving under it and it comes off at the end.
        uint32_t K65535;                          // kept opaque: the compiler otherwise splits x * 65535 + A into a shift-add and a subtract
        asm("mov.u32 %0, 65535;" : "=r"(K65535));
        uint32_t A = 0x4000u - (uint32_t)v * K65535;
        asm("" : "+r"(A));                        // ... or refactors it into (ring - centre) * 65535 + 0x4000, again two instructions per ring pixel
        uint32_t wv[16];
#pragma unroll
        for (int k = 0; k < 16; k++) wv[k] = (uint32_t)p[ring_off(k)] * K65535 + A;
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
        const int sd = (int)(mx & 0xFFFFu) - 0x4000, sb = (int)mx >> 16;
        const int sc = __vimax3_s32(sd, sb, FAST_T) - 1;
        s_score[e] = (uint8_t)sc;
    }
    __syncthreads();

    // ---- D. cell-local 3x3 NMS over the interior, stage survivors (s_emit aliases the plane buffer)
    const int wbase = (tid >> 5) * EMIT_W;
    int wn = 0;
    for (int i0 = 0; i0 < ncorner; i0 += FT_THREADS) {      // warp-uniform trip count: the ballot below needs all lanes
        const int i = i0 + tid;
        // branch-free: the eight neighbour scores are fetched together (independent loads, every address stays inside
        // s_score for an interior pixel) and neighbours that belong to another cell are replaced by 0
        const int e = i < ncorner ? (int)s_list[i] : (FT_SP + FT_C0);
        const int r = e / FT_SP, cidx = e - r * FT_SP;
        const bool inner = i < ncorner && r >= 1 && r <= FT_OH && cidx >= FT_C0 && cidx < FT_C0 + FT_OW;      // else: halo, belongs to the neighbouring tile
        const uint8_t *q = s_score + (inner ? e : FT_SP + FT_C0);
        const int s = q[0];
        const int nl = q[-1], nr = q[1], nu = q[-FT_SP], nul = q[-FT_SP - 1], nur = q[-FT_SP + 1], nd = q[FT_SP], ndl = q[FT_SP - 1], ndr = q[FT_SP + 1];
        // masks applied as multiplies by 0 / 1 (FMA pipe) instead of selects (ALU pipe, the bottleneck)
        const int Lk = s_cl[cidx], Rk = s_cr[cidx], Uk = s_ru[r], Dk = s_rd[r];
        const int m0 = __vimax3_s32(nl * Lk, nr * Rk, nu * Uk);
        const int m1 = __vimax3_s32(nul * (Uk * Lk), nur * (Uk * Rk), nd * Dk);
        const int m2 = __vimax3_s32(ndl * (Dk * Lk), ndr * (Dk * Rk), m0);
        const bool kept = inner && s > max(m1, m2);
        const uint32_t packed = pack_cand(x0 + cidx - LATTICE_MIN, sy0 + r - LATTICE_MIN, s);
        // survivors go to the WARP's own staging segment (count in a warp-uniform register: no shared-memory atomic, no CTA barrier)
        const unsigned bal = __ballot_sync(0xffffffffu, kept);
        if (bal) {
            const int pos = wn + __popc(bal & ((1u << (tid & 31)) - 1));
            if (kept && pos < EMIT_W) s_emit[wbase + pos] = packed;
            wn += __popc(bal);
        }
    }
    if (wn > 0) {       // warp-uniform: the warp reserves its range of the level's candidate list and copies its segment out
        const int lane = tid & 31;
        if (wn > EMIT_W) { if (lane == 0) atomicOr(status, ST_CAND_OVERFLOW); wn = EMIT_W; }
        int base = 0;
        if (lane == 0) base = atomicAdd(candCount + b * HYORB_MAX_LEVELS + l, wn);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();            // the staged entries of the other lanes
        uint32_t *out = cand + (size_t)b * plan->candStride + L.candOff;
        for (int i = lane; i < wn; i += 32) {
            if (base + i < L.candCap) out[base + i] = s_emit[wbase + i];
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
        if (const char *v = getenv("HYORB_FT_CTAS")) { const int want = atoi(v); if (want >= 1 && want < per) per = want; }   // leave room for other lanes' kernels
        ctas_per_sm.store(per, std::memory_order_relaxed);
    }
    const int sms = sm_count > 0 ? sm_count : 148;
    const long long nTiles = (long long)hp.tilesPerImage * B;
    if (nTiles > 0x7fffffffLL) { set_error("too many FAST tiles in one batch"); return HYORB_EUNSUPPORTED; }
    const int grid = (int)std::min<long long>(nTiles, (long long)sms * per);
    HY_CUDA(launch_dependent(k_fast, dim3(grid), dim3(FT_THREADS), 0, st, dp, tm0, tmaps, img0, (int)nTiles, cand, candCount, status));
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
