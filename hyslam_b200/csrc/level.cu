// level.cu -- K1+K4 fused: one pass over pyramid level l produces BOTH its blurred copy and level l+1.
//
// Replaces, per level, `cv::resize(level l, level l+1, INTER_LINEAR)` of ORBExtractor::ComputePyramid (src/features/ORBExtractor.cpp:564-589)
// and `GaussianBlur(level l clone, 7x7, 2, 2, BORDER_REFLECT_101)` of ORBExtractor::operator() (:536-537).  Arithmetic is the one of
// pyramid.cu / blur.cu (OpenCV's 8-bit fixed-point kernels) -- only the data movement changes:
//   * every WARP is its own pipeline: it walks 128 x 20 pixel tiles of the SOURCE level; a tile and its halo (16 columns either side -- TMA
//     boxes start at multiples of 16 bytes -- and 3 rows above / below) arrive in the warp's own shared-memory ring as ONE
//     cp.async.bulk.tensor box per tile, the next boxes in flight while the current one is processed: no thread waits on a global load
//     (the stand-alone kernels spent half of their stall samples on the long scoreboard) and there is no CTA-wide barrier at all;
//   * the level is read from HBM once instead of twice (B_pyr + B_desc share the read of every source level);
//   * BORDER_REFLECT_101 is written into the box once per edge tile, so the two passes below carry no edge logic;
//   * blur: lane = 4 columns walking down the tile, horizontal 7-tap as 2 DP4A per pixel, vertical 7-tap as 3 DP2A + 1 multiply-add per
//     pixel on PAIRS of 16-bit row sums kept in registers; then the same warp produces the pixels of level l+1 whose first bilinear tap
//     lies in the tile (PRMT + DP2A horizontal, IMAD.HI vertical), straight from the box.
// A destination 4-pixel group that straddles two tiles is computed by both (all its taps lie inside either halo) and stored twice with
// identical bytes, which keeps every store an aligned 32-bit word.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "tma.cuh"

namespace hyorb {

constexpr int LV_BOX = LV_BW * LV_BH;                  // bytes of one box
constexpr int LV_BUF = (LV_BOX + 127) & ~127;          // buffer stride: TMA destinations are 128-byte aligned
#ifndef HYORB_LV_NBUF
#define HYORB_LV_NBUF 2
#endif
constexpr int LV_NBUF = HYORB_LV_NBUF;         // boxes per warp: one being processed, the others in flight
constexpr size_t LV_SMEM = (size_t)LV_WARPS * LV_NBUF * LV_BUF + sizeof(uint64_t) * LV_WARPS * LV_NBUF;

// 7-tap row sums of the 4 pixels x0 .. x0+3 from the 12 bytes x0-4 .. x0+7 (q = word holding pixel x0-4): 6 funnel shifts + 8 DP4A
__device__ __forceinline__ void lv_hsum4(const uint32_t *q, uint32_t (&h)[4])
{
    const uint32_t G0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps -3..0
    const uint32_t G1 = 48u | (34u << 8) | (18u << 16);                   // taps +1..+3
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    h[0] = __dp4a(__funnelshift_r(w0, w1, 8), G0, __dp4a(__funnelshift_r(w1, w2, 8), G1, 0u));
    h[1] = __dp4a(__funnelshift_r(w0, w1, 16), G0, __dp4a(__funnelshift_r(w1, w2, 16), G1, 0u));
    h[2] = __dp4a(__funnelshift_r(w0, w1, 24), G0, __dp4a(__funnelshift_r(w1, w2, 24), G1, 0u));
    h[3] = __dp4a(w1, G0, __dp4a(w2, G1, 0u));
}

// ---- resize pieces (see pyramid.cu) ----
struct LvH { uint32_t h[4]; };
__device__ __forceinline__ LvH lv_hrow(const uint8_t *boxrow, int so, unsigned sh, const uint32_t (&sel)[4], const uint32_t (&c01)[4])
{
    // bytes s0 .. s0+7 of the source row: three aligned words around box offset `so`, funnel-shifted to the first tap
    const uint32_t *p = (const uint32_t *)(boxrow + (so & ~3));
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    const uint32_t A = __funnelshift_r(w0, w1, sh), B = __funnelshift_r(w1, w2, sh);
    LvH r;
#pragma unroll
    for (int j = 0; j < 4; j++) r.h[j] = __dp2a_lo(c01[j], __byte_perm(A, B, sel[j]), 0u) >> 4;
    return r;
}

// BORDER_REFLECT_101 written INTO the box (the TMA zero-fills outside the image), so that the passes below need no edge logic:
// 3 columns left of x = 0 / right of x = w-1, then 3 whole rows above y = 0 / below y = h-1 (corners come out reflected in both axes)
__device__ __forceinline__ void lv_reflect_box(uint8_t *box, int lane, int X0, int Y0, int w, int h)
{
    const bool eL = X0 == 0, eR = X0 + LV_TW + LV_HY > w, eT = Y0 == 0, eB = Y0 + LV_TH + LV_HY > h;
    if (!(eL | eR | eT | eB)) return;
    const int byLo = eT ? LV_HY : 0, byHi = min(LV_BH, h - (Y0 - LV_HY));        // box rows that hold image rows
    if (eL)
        for (int by = byLo + lane; by < byHi; by += 32) {
            uint32_t *r = (uint32_t *)(box + by * LV_BW);
            r[LV_HX / 4 - 1] = __byte_perm(r[LV_HX / 4], r[LV_HX / 4 + 1], 0x1234);     // pixels -4..-1 = pixels 4, 3, 2, 1
        }
    if (eR) {
        const int bxw = w - (X0 - LV_HX);                                            // box column of x = w
        for (int by = byLo + lane; by < byHi; by += 32) {
            uint8_t *r = box + by * LV_BW;
            const uint8_t a = r[bxw - 2], b = r[bxw - 3], c = r[bxw - 4];
            r[bxw] = a; r[bxw + 1] = b; r[bxw + 2] = c;
        }
    }
    __syncwarp();
    if (eT)
        for (int i = lane; i < LV_HY * (LV_BW / 4); i += 32) {
            const int k = i / (LV_BW / 4), c = i - k * (LV_BW / 4);
            ((uint32_t *)(box + (LV_HY - 1 - k) * LV_BW))[c] = ((const uint32_t *)(box + (LV_HY + 1 + k) * LV_BW))[c];
        }
    if (eB) {
        const int byh = h - (Y0 - LV_HY);                                            // box row of y = h
        for (int i = lane; i < LV_HY * (LV_BW / 4); i += 32) {
            const int k = i / (LV_BW / 4), c = i - k * (LV_BW / 4);
            if (byh + k < LV_BH && byh - 2 - k >= 0) ((uint32_t *)(box + (byh + k) * LV_BW))[c] = ((const uint32_t *)(box + (byh - 2 - k) * LV_BW))[c];
        }
    }
    __syncwarp();
}

#ifndef HYORB_LV_MINB
#define HYORB_LV_MINB 6
#endif
// Every WARP is its own pipeline: it walks 128 x 20 pixel tiles of the source level (tile = global warp index, + number of warps, ...),
// owns LV_NBUF box buffers and mbarriers, and never meets a CTA-wide barrier.
__global__ void __launch_bounds__(LV_THREADS, HYORB_LV_MINB)
k_level(const PlanDev *__restrict__ plan, int l, const __grid_constant__ CUtensorMap tm0, const CUtensorMap *__restrict__ tmaps, int img0,
        uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, const ResizeTab *__restrict__ tabs, const int *__restrict__ lvtab, int nTiles, int do_resize)
{
    extern __shared__ __align__(128) uint8_t lv_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *const boxes = lv_smem + (size_t)warp * LV_NBUF * LV_BUF;
    uint64_t *const bars = (uint64_t *)(lv_smem + (size_t)LV_WARPS * LV_NBUF * LV_BUF) + warp * LV_NBUF;
    grid_dependency_wait();      // launch_dependent: follows the launch that produced this level
    const LevelDev &L = plan->lv[l];
    const int tilesX = L.lvTilesX, perImage = L.lvTilesX * L.lvTilesY;
    const CUtensorMap *tm = l == 0 ? &tm0 : &tmaps[l];
    const long long gw = (long long)blockIdx.x * LV_WARPS + warp, nw = (long long)gridDim.x * LV_WARPS;
    // Tile coordinates advance incrementally (a tile index -> (image, tile row, tile column) decode costs two integer divisions, ~80
    // instructions, and a warp visits a tile every ~3000): `cur` is the tile being processed, `pre` the one being prefetched.
    struct TileId { int b, ty, tx; };
    const int tilesY = L.lvTilesY;
    auto decode = [&](long long T) { TileId t; t.b = (int)(T / perImage); const int r = (int)(T - (long long)t.b * perImage); t.ty = r / tilesX; t.tx = r - t.ty * tilesX; return t; };
    const TileId step = decode(nw);
    auto advance = [&](TileId &t) {
        t.tx += step.tx; if (t.tx >= tilesX) { t.tx -= tilesX; t.ty++; }
        t.ty += step.ty; if (t.ty >= tilesY) { t.ty -= tilesY; t.b++; }
        t.b += step.b;
    };
    auto issue = [&](const TileId &t, int buf) {
        mbar_arrive_expect_tx(&bars[buf], LV_BOX);
        tma_load_3d(boxes + buf * LV_BUF, tm, &bars[buf], t.tx * LV_TW - LV_HX, t.ty * LV_TH - LV_HY, img0 + t.b);
    };
    TileId cur = decode(gw), pre = cur;
    if (lane == 0) {
        for (int k = 0; k < LV_NBUF; k++) mbar_init(&bars[k], 1);
        mbar_fence_init();
        fence_proxy_async();
        if (l > 0) tensormap_acquire(tm);
    }
    for (int k = 0; k < LV_NBUF - 1; k++) {            // `pre` ends LV_NBUF - 1 tiles ahead of `cur` on every lane
        if (lane == 0 && gw + k * nw < nTiles) issue(pre, k);
        advance(pre);
    }
    __syncwarp();

    const int w = L.w, h = L.h;
    int it = 0;
    for (long long TT = gw; TT < nTiles; TT += nw, it++, advance(cur), advance(pre)) {
        const int buf = it % LV_NBUF;
        // buffer (it - 1) % LV_NBUF was released by the __syncwarp that ended the previous iteration: refill it with the tile LV_NBUF - 1 ahead
        if (lane == 0 && TT + (LV_NBUF - 1) * nw < nTiles) { fence_proxy_async(); issue(pre, (it + LV_NBUF - 1) % LV_NBUF); }
        const int b = cur.b, ty = cur.ty, tx = cur.tx;
        const int X0 = tx * LV_TW, Y0 = ty * LV_TH;
        uint8_t *box = boxes + buf * LV_BUF;            // box(0,0) = image (X0 - 16, Y0 - 3)
        mbar_wait(&bars[buf], (it / LV_NBUF) & 1);
        lv_reflect_box(box, lane, X0, Y0, w, h);

        // ---------------- blur: lane = 4 columns, walking down the tile with the row sums of the last 6 rows in registers ----------------
        const int x0 = X0 + 4 * lane;
        if (x0 < w) {
            const int rows = min(LV_TH, h - Y0);
            const bool full = x0 + 3 < w;
            uint8_t *out = blur + (size_t)b * plan->pyrStride + L.off + (size_t)Y0 * L.pitch + x0;
            const int opitch = L.pitch;
            const uint32_t *q = (const uint32_t *)(box + (LV_HX - 4) + 4 * lane);          // word holding pixel x0-4 in box row 0
            // vertical pass on PAIRS of row sums (a row sum is <= 255 * 256, 16 bits): out(r) = dp2a(Q[r], g0 g1) + dp2a(Q[r+2], g2 g3) +
            // dp2a(Q[r+4], g4 g5) + g6 * hsum(r+6) + 2^15, Q[t] = hsum(t) | hsum(t+1) << 16 -- 5 multiply-adds per pixel instead of 7 + 3 adds
            const uint32_t G01 = 18u | (34u << 8), G23 = 48u | (56u << 8), G45 = 48u | (34u << 8);
            uint32_t Q[5][4], hl[4];
            {
                uint32_t hp[4];
                lv_hsum4(q, hp);
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    lv_hsum4(q + (k + 1) * (LV_BW / 4), hl);
#pragma unroll
                    for (int j = 0; j < 4; j++) { Q[k][j] = __byte_perm(hp[j], hl[j], 0x5410); hp[j] = hl[j]; }
                }
            }
            uint8_t *d = out;
            for (int r0 = 0; r0 < rows; r0 += 5) {
#pragma unroll
                for (int k = 0; k < 5; k++, d += opitch) {
                    const int r = r0 + k;
                    uint32_t hn[4], v[4];
                    lv_hsum4(q + (r + 6) * (LV_BW / 4), hn);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        v[j] = __dp2a_lo(Q[k][j], G01, __dp2a_lo(Q[(k + 2) % 5][j], G23, __dp2a_lo(Q[(k + 4) % 5][j], G45, hn[j] * 18u + 32768u)));
                        Q[k][j] = __byte_perm(hl[j], hn[j], 0x5410);        // Q[r + 5] takes the slot of Q[r]
                        hl[j] = hn[j];
                    }
                    const uint32_t o = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);     // (v >> 16) of the four
                    if (r < rows) {
                        if (full) *(uint32_t *)d = o;
                        else for (int j = 0; x0 + j < w; j++) d[j] = (uint8_t)(o >> (8 * j));
                    }
                }
            }
        }
        if (do_resize) {
            // ---------------- level l+1: destination rows whose first tap row lies in this tile, destination columns likewise ----------------
            // The warp walks the SOURCE rows Y0 .. Y0+20 once (horizontal pass of each: PRMT + DP2A on the lane's 4 destination columns) and
            // emits a destination row whenever its second tap row has just been computed (tinv: source row -> destination row + weights).
            const LevelDev &D = plan->lv[l + 1];
            const int dw = D.w;
            const int dxa = lvtab[L.lvDx + tx], dxb = lvtab[L.lvDx + tx + 1];
            const int dya = lvtab[L.lvDy + ty];             // first destination row of this tile; the rows it emits are dya, dya + 1, ...
            const int sEnd = min(Y0 + LV_TH, h - 1);
            uint8_t *dimg = pyr + (size_t)b * plan->pyrStride + D.off;
            const int dpitch = D.pitch;
            for (int g = (dxa >> 2) + lane; 4 * g < dxb; g += 32) {
                const int x4 = 4 * g;
                // column constants of this 4-pixel destination group, precomputed on the host (level_tiles): first source column, the
                // byte selectors of the four tap pairs relative to it, the four Q11 coefficient pairs
                const uint4 *ce = (const uint4 *)(lvtab + L.lvCol) + 3 * (size_t)g;
                const uint4 e0 = ce[0], e1 = ce[1], e2 = ce[2];
                const int s0 = (int)e0.x;
                const uint32_t sel[4] = {e0.y, e0.z, e0.w, e1.x}, c01[4] = {e1.y, e1.z, e1.w, e2.x};
                const int so = s0 - (X0 - LV_HX);                               // box column of the first tap
                const unsigned shf = (unsigned)(so & 3) * 8;
                const bool full = x4 + 3 < dw;
                const uint8_t *brow = box + LV_HY * LV_BW;                       // box row of source row Y0
                const ResizeTab *ti = tabs + D.rsInv + Y0 + 1;                  // entry of source row s: the destination row whose taps are (s-1, s)
                uint8_t *d = dimg + (size_t)dya * dpitch + x4;
                auto emit = [&](const ResizeTab e, const LvH &h0, const LvH &h1) {
                    if (e.ofs < 0) return;                                      // warp-uniform: no destination row ends on this source row
                    const uint32_t b0 = (uint32_t)(uint16_t)e.c0 << 16, b1 = (uint32_t)(uint16_t)e.c1 << 16;
                    uint32_t v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = __umulhi(b0, h0.h[j]) + __umulhi(b1, h1.h[j]) + 2u;      // <= 1023; the result is v >> 2
                    const uint32_t p01 = (v[0] | (v[1] << 16)) >> 2, p23 = (v[2] | (v[3] << 16)) >> 2;
                    const uint32_t o = __byte_perm(p01, p23, 0x6420);
                    if (full) *(uint32_t *)d = o;
                    else for (int j = 0; x4 + j < dw; j++) d[j] = (uint8_t)(o >> (8 * j));
                    d += dpitch;
                };
                LvH hA = lv_hrow(brow, so, shf, sel, c01), hB;
                ResizeTab e = ti[0];
                for (int sr = Y0 + 1; sr <= sEnd; sr += 2) {
                    brow += LV_BW;
                    hB = lv_hrow(brow, so, shf, sel, c01);
                    const ResizeTab e1 = ti[1];
                    emit(e, hA, hB);
                    if (sr + 1 > sEnd) break;
                    brow += LV_BW;
                    hA = lv_hrow(brow, so, shf, sel, c01);
                    e = ti[2];
                    ti += 2;
                    emit(e1, hB, hA);
                }
            }
        }
        fence_proxy_async();  // edge tiles wrote into the box (lv_reflect_box): order those generic-proxy stores before the TMA write that reuses it
        __syncwarp();         // the box is free for the load after next
    }
}

int launch_levels(const PlanDev &hp, const PlanDev *dp, const CUtensorMap &tmL0, const CUtensorMap *tmaps, int img0, Level0 l0, uint8_t *pyr, uint8_t *blur,
                  const ResizeTab *tabs, const int *lvtab, int B, int sm_count, cudaStream_t st, long *launches)
{
    static std::atomic<int> ctas_per_sm{0};
    int per = ctas_per_sm.load(std::memory_order_relaxed);
    if (!per) {
        HY_CUDA(cudaFuncSetAttribute(k_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LV_SMEM));
        HY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_level, LV_THREADS, LV_SMEM));
        if (per < 1) per = 1;
        ctas_per_sm.store(per, std::memory_order_relaxed);
    }
    const int sms = sm_count > 0 ? sm_count : 148;
    for (int l = 0; l < hp.nlevels; l++) {
        const LevelDev &L = hp.lv[l];
        const long long nTiles = (long long)L.lvTilesX * L.lvTilesY * B;
        if (nTiles > 0x7fffffffLL) { set_error("too many level tiles in one batch"); return HYORB_EUNSUPPORTED; }
        const bool has_next = l + 1 < hp.nlevels;
        const int do_resize = has_next && hp.lv[l + 1].lvFused;
        const int grid = (int)std::min<long long>((nTiles + LV_WARPS - 1) / LV_WARPS, (long long)sms * per);
        if (l > 0)
            HY_CUDA(launch_dependent(k_level, dim3(grid), dim3(LV_THREADS), LV_SMEM, st, dp, l, tmL0, tmaps, img0, pyr, blur, tabs, lvtab, (int)nTiles, do_resize));
        else
            k_level<<<grid, LV_THREADS, LV_SMEM, st>>>(dp, l, tmL0, tmaps, img0, pyr, blur, tabs, lvtab, (int)nTiles, do_resize);
        ++*launches;
        if (has_next && !do_resize) HY_TRY(launch_resize_level(hp, l + 1, l0, pyr, tabs, B, st, launches));     // cv::resize's exact-2x INTER_AREA shortcut
    }
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

// host: per source level, the first destination column / row owned by every tile column / row (+ one-past-the-end entries)
void level_tiles(HostPlan *plan)
{
    PlanDev &P = plan->dev;
    plan->lvtab.clear();
    for (int l = 0; l < P.nlevels; l++) {
        LevelDev &L = P.lv[l];
        L.lvTilesX = (L.w + LV_TW - 1) / LV_TW;
        L.lvTilesY = (L.h + LV_TH - 1) / LV_TH;
        L.lvDx = L.lvDy = 0;
        if (l + 1 >= P.nlevels) continue;
        LevelDev &D = P.lv[l + 1];
        // source row -> the destination row whose two tap rows are (s-1, s), or -1.  The fused kernel needs every destination row to have taps
        // (ofs, ofs + 1) inside the image and no two destination rows on the same source pair (true for every down-scale); otherwise the level
        // is produced by k_resize (lvFused = 0).
        D.rsInv = (int)plan->resize.size();
        plan->resize.resize(plan->resize.size() + L.h + 4, ResizeTab{-1, 0, 0});
        D.lvFused = !D.area2x;
        {
            const ResizeTab *tyy = plan->resize.data() + D.rsY;
            ResizeTab *inv = plan->resize.data() + D.rsInv;
            for (int y = 0; y < D.h && D.lvFused; y++) {
                const int o = tyy[y].ofs;
                if (o < 0 || o + 1 > L.h - 1 || inv[o + 1].ofs >= 0) { D.lvFused = 0; break; }
                inv[o + 1] = ResizeTab{y, tyy[y].c0, tyy[y].c1};
            }
        }
        const ResizeTab *tx = plan->resize.data() + D.rsX, *ty = plan->resize.data() + D.rsY;
        L.lvDx = (int)plan->lvtab.size();
        for (int t = 0, d = 0; t <= L.lvTilesX; t++) {
            while (d < D.w && tx[d].ofs < t * LV_TW) d++;
            plan->lvtab.push_back(t == L.lvTilesX ? D.w : d);
        }
        // per destination 4-pixel group: {s0, sel[4], c01[4]} padded to 12 words, 16-byte aligned
        while (plan->lvtab.size() % 4) plan->lvtab.push_back(0);
        L.lvCol = (int)plan->lvtab.size();
        for (int x4 = 0; x4 < D.w; x4 += 4) {
            int ent[12] = {0};
            const int s0 = tx[x4].ofs;
            ent[0] = s0;
            for (int j = 0; j < 4; j++) {
                const ResizeTab t = tx[std::min(x4 + j, D.w - 1)];
                const unsigned dlt = (unsigned)(t.ofs - s0);
                ent[1 + j] = (int)(dlt | ((dlt + 1) << 4));
                ent[5 + j] = (int)((uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16));
            }
            for (int k = 0; k < 12; k++) plan->lvtab.push_back(ent[k]);
        }
        L.lvDy = (int)plan->lvtab.size();
        for (int t = 0, d = 0; t <= L.lvTilesY; t++) {
            auto sy0 = [&](int i) { const int o = ty[i].ofs; return o < 0 ? 0 : (o > L.h - 1 ? L.h - 1 : o); };
            while (d < D.h && sy0(d) < t * LV_TH) d++;
            plan->lvtab.push_back(t == L.lvTilesY ? D.h : d);
        }
    }
    if (plan->lvtab.empty()) plan->lvtab.push_back(0);
}

}  // namespace hyorb
