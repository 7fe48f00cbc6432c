// level.cu -- K1+K4 fused: one pass over pyramid level l produces BOTH its blurred copy and level l+1.
//
// Replaces, per level, `cv::resize(level l, level l+1, INTER_LINEAR)` of ORBExtractor::ComputePyramid (src/features/ORBExtractor.cpp:564-589)
// and `GaussianBlur(level l clone, 7x7, 2, 2, BORDER_REFLECT_101)` of ORBExtractor::operator() (:536-537).  Arithmetic is the one of
// pyramid.cu / blur.cu (OpenCV's 8-bit fixed-point kernels) -- only the data movement changes:
//   * persistent CTAs walk 128 x 56 pixel tiles of the SOURCE level; a tile and its halo (16 columns either side -- TMA boxes start at
//     multiples of 16 bytes -- and 3 rows above / below) arrive in shared memory as ONE cp.async.bulk.tensor box, double-buffered, so the
//     next tile is in flight while the current one is processed and no warp waits on a global load (the stand-alone kernels spent half of
//     their stall samples on the long scoreboard);
//   * the level is read from HBM once instead of twice (B_pyr + B_desc share the read of every source level);
//   * warps 0-3 blur the tile's interior (4 columns x 14 rows per thread, 7-row register window, DP4A), warps 4-5 produce the pixels of
//     level l+1 whose first bilinear tap lies in the interior (PRMT + DP2A horizontal, IMAD.HI vertical), both straight from the box.
// A destination 4-pixel group that straddles two tiles is computed by both (all its taps lie inside either halo) and stored twice with
// identical bytes, which keeps every store an aligned 32-bit word.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "tma.cuh"

namespace hyorb {

constexpr int LV_BUF = (LV_BW * LV_BH + 127) & ~127;
#ifndef HYORB_LV_NBUF
#define HYORB_LV_NBUF 4
#endif
constexpr int LV_NBUF = HYORB_LV_NBUF;

// ---- blur pieces (see blur.cu for the derivation of the edge permutes) ----
__device__ __forceinline__ int lv_reflect101(int i, int n)
{
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}
struct LvEdge { uint32_t sel1, sel2; int edge, pair2; };
__device__ __forceinline__ LvEdge lv_edge(int x0, int w)
{
    LvEdge f;
    f.sel1 = 0x7654u; f.sel2 = 0x7654u; f.pair2 = 1; f.edge = 0;
    if (x0 == 0) f.edge = 1;
    const int e = w - (x0 - 4);
    if (e < 12) {
        f.edge = 2;
        f.pair2 = e >= 8;
        uint32_t s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i1 = 4 + k, i2 = 8 + k;
            int a = i1 < e ? i1 : 2 * (e - 1) - i1;
            a = a < 0 ? 0 : a;
            int c = i2 < e ? i2 : 2 * (e - 1) - i2;
            c -= f.pair2 ? 4 : 0;
            c = c < 0 ? 0 : (c > 7 ? 7 : c);
            s1 |= (uint32_t)a << (4 * k);
            s2 |= (uint32_t)c << (4 * k);
        }
        f.sel1 = s1; f.sel2 = s2;
    }
    return f;
}
struct LvRow3 { uint32_t w0, w1, w2; };
// the 12 bytes x0-4 .. x0+7 of one box row (q = address of the word that holds pixel x0-4)
__device__ __forceinline__ LvRow3 lv_load_row(const uint32_t *q, const LvEdge &f)
{
    LvRow3 r;
    r.w0 = q[0]; r.w1 = q[1]; r.w2 = q[2];
    if (f.edge) {
        if (f.edge == 1) r.w0 = __byte_perm(r.w1, r.w2, 0x1234);                // REFLECT_101: pixels -4..-1 = pixels 4,3,2,1
        else {
            const uint32_t n1 = __byte_perm(r.w0, r.w1, f.sel1);
            const uint32_t n2 = f.pair2 ? __byte_perm(r.w1, r.w2, f.sel2) : __byte_perm(r.w0, r.w1, f.sel2);
            r.w1 = n1; r.w2 = n2;
        }
    }
    return r;
}
__device__ __forceinline__ void lv_hsum4(const LvRow3 &r, int (&h)[4])
{
    const uint32_t G0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps -3..0
    const uint32_t G1 = 48u | (34u << 8) | (18u << 16);                   // taps +1..+3
    h[0] = (int)__dp4a(__funnelshift_r(r.w0, r.w1, 8), G0, __dp4a(__funnelshift_r(r.w1, r.w2, 8), G1, 0u));
    h[1] = (int)__dp4a(__funnelshift_r(r.w0, r.w1, 16), G0, __dp4a(__funnelshift_r(r.w1, r.w2, 16), G1, 0u));
    h[2] = (int)__dp4a(__funnelshift_r(r.w0, r.w1, 24), G0, __dp4a(__funnelshift_r(r.w1, r.w2, 24), G1, 0u));
    h[3] = (int)__dp4a(r.w1, G0, __dp4a(r.w2, G1, 0u));
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

#ifndef HYORB_LV_MINB
#define HYORB_LV_MINB 4
#endif
__global__ void __launch_bounds__(LV_THREADS, HYORB_LV_MINB)
k_level(const PlanDev *__restrict__ plan, int l, const __grid_constant__ CUtensorMap tm0, const CUtensorMap *__restrict__ tmaps, int img0,
        uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, const ResizeTab *__restrict__ tabs, const int *__restrict__ lvtab, int nTiles, int do_resize)
{
    __shared__ __align__(128) uint8_t s_box[LV_NBUF][LV_BUF];
    __shared__ __align__(8) uint64_t s_bar[LV_NBUF];
    grid_dependency_wait();      // launch_dependent: follows the launch that produced this level
    const int tid = threadIdx.x;
    const LevelDev &L = plan->lv[l];
    const int tilesX = L.lvTilesX, perImage = L.lvTilesX * L.lvTilesY;
    const CUtensorMap *tm = l == 0 ? &tm0 : &tmaps[l];
    auto issue = [&](int T, int buf) {
        const int b = T / perImage, r = T - b * perImage;
        const int ty = r / tilesX, tx = r - ty * tilesX;
        mbar_arrive_expect_tx(&s_bar[buf], LV_BW * LV_BH);
        tma_load_3d(s_box[buf], tm, &s_bar[buf], tx * LV_TW - LV_HX, ty * LV_TH - LV_HY, img0 + b);
    };
    if (tid == 0) {
        for (int k = 0; k < LV_NBUF; k++) mbar_init(&s_bar[k], 1);
        mbar_fence_init();
        fence_proxy_async();
        if (l > 0) tensormap_acquire(tm);
        // a box is 62 rows of 160 bytes -- 62 separate memory segments -- and takes microseconds to arrive: keep LV_NBUF - 1 boxes in flight
        for (int k = 0; k < LV_NBUF - 1; k++)
            if ((long long)blockIdx.x + (long long)k * gridDim.x < nTiles) issue(blockIdx.x + k * gridDim.x, k);
    }
    __syncthreads();

    const int w = L.w, h = L.h;
    int it = 0;
    for (int T = blockIdx.x; T < nTiles; T += gridDim.x, it++) {
        const int buf = it % LV_NBUF;
        // buffer (it - 1) % LV_NBUF was released by the barrier that ended the previous iteration: refill it with the tile LV_NBUF - 1 ahead
        if (tid == 0 && (long long)T + (long long)(LV_NBUF - 1) * gridDim.x < nTiles) { fence_proxy_async(); issue(T + (LV_NBUF - 1) * gridDim.x, (it + LV_NBUF - 1) % LV_NBUF); }
        const int b = T / perImage, rem = T - b * perImage;
        const int ty = rem / tilesX, tx = rem - ty * tilesX;
        const int X0 = tx * LV_TW, Y0 = ty * LV_TH;
        const uint8_t *box = s_box[buf];            // box(0,0) = image (X0 - 16, Y0 - 3)
        mbar_wait(&s_bar[buf], (it / LV_NBUF) & 1);

        if (tid < 128) {
            // ---------------- blur: 4 columns x 14 rows per thread ----------------
            const int cg = tid & 31, strip = tid >> 5;
            const int x0 = X0 + 4 * cg, y0 = Y0 + 14 * strip;
            if (x0 < w && y0 < h) {
                const LvEdge fx = lv_edge(x0, w);
                const int yEnd = min(y0 + 14, h);
                const bool full = x0 + 3 < w;
                uint8_t *out = blur + (size_t)b * plan->pyrStride + L.off + x0;
                const int opitch = L.pitch;
                const uint32_t *col = (const uint32_t *)(box + 4 * cg + (LV_HX - 4));          // word holding pixel x0-4 in box row 0
                auto rowp = [&](int y) { return col + (lv_reflect101(y, h) - (Y0 - LV_HY)) * (LV_BW / 4); };
                int win[7][4];
                {
                    LvRow3 pr[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) pr[k] = lv_load_row(rowp(y0 - 3 + k), fx);
#pragma unroll
                    for (int k = 0; k < 6; k++) lv_hsum4(pr[k], win[k]);
                }
#pragma unroll
                for (int half = 0; half < 2; half++) {
#pragma unroll
                    for (int k = 0; k < 7; k++) {
                        const int y = y0 + 7 * half + k;
                        // rows past the image bottom are computed from reflected (in-box) rows and never stored
                        const LvRow3 cur = lv_load_row(rowp(min(y + 3, h + 2)), fx);
                        lv_hsum4(cur, win[(6 + k) % 7]);
                        uint32_t o = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t v = 18u * (uint32_t)(win[k % 7][j] + win[(k + 6) % 7][j]) + 34u * (uint32_t)(win[(k + 1) % 7][j] + win[(k + 5) % 7][j]) +
                                               48u * (uint32_t)(win[(k + 2) % 7][j] + win[(k + 4) % 7][j]) + 56u * (uint32_t)win[(k + 3) % 7][j];
                            o |= ((v + 32768u) >> 16) << (8 * j);
                        }
                        if (y < yEnd) {
                            uint8_t *d = out + (size_t)y * opitch;
                            if (full) *(uint32_t *)d = o;
                            else for (int j = 0; x0 + j < w; j++) d[j] = (uint8_t)(o >> (8 * j));
                        }
                    }
                }
            }
        } else if (do_resize) {
            // ---------------- level l+1: destination pixels whose first taps (ofs_x, clamped ofs_y) lie in this tile's interior ----------------
            const LevelDev &D = plan->lv[l + 1];
            const ResizeTab *tabx = tabs + D.rsX, *taby = tabs + D.rsY;
            const int dw = D.w;
            const int dxa = lvtab[L.lvDx + tx], dxb = lvtab[L.lvDx + tx + 1], dya = lvtab[L.lvDy + ty], dyb = lvtab[L.lvDy + ty + 1];
            const int lane = tid & 31, wsel = (tid >> 5) - 4;                  // two warps split the rows
            const int nrows = dyb - dya, half = (nrows + 1) >> 1;
            const int ya = dya + wsel * half, yb = min(ya + half, dyb);
            uint8_t *dimg = pyr + (size_t)b * plan->pyrStride + D.off;
            for (int g = (dxa >> 2) + lane; 4 * g < dxb; g += 32) {
                const int x4 = 4 * g;
                int s0 = 0;
                uint32_t sel[4], c01[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const ResizeTab t = tabx[min(x4 + j, dw - 1)];
                    if (j == 0) s0 = t.ofs;
                    const uint32_t dlt = (uint32_t)(t.ofs - s0);
                    sel[j] = dlt | ((dlt + 1) << 4);
                    c01[j] = (uint32_t)(uint16_t)t.c0 | ((uint32_t)(uint16_t)t.c1 << 16);
                }
                const int so = s0 - (X0 - LV_HX);                               // box column of the first tap
                const unsigned shf = (unsigned)(so & 3) * 8;
                const bool full = x4 + 3 < dw;
                int haveRow = -0x7fffffff;
                LvH ha, hb;
                hb.h[0] = hb.h[1] = hb.h[2] = hb.h[3] = 0;
                for (int y = ya; y < yb; y++) {
                    const ResizeTab vy = taby[y];
                    const int sy0 = min(max(vy.ofs, 0), h - 1), sy1 = min(max(vy.ofs + 1, 0), h - 1);
                    ha = sy0 != haveRow ? lv_hrow(box + (sy0 - (Y0 - LV_HY)) * LV_BW, so, shf, sel, c01) : hb;
                    hb = sy1 != sy0 ? lv_hrow(box + (sy1 - (Y0 - LV_HY)) * LV_BW, so, shf, sel, c01) : ha;
                    haveRow = sy1;
                    const uint32_t b0 = (uint32_t)vy.c0 << 16, b1 = (uint32_t)vy.c1 << 16;
                    uint32_t o = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t v = (__umulhi(b0, ha.h[j]) + __umulhi(b1, hb.h[j]) + 2u) >> 2;
                        o |= v << (8 * j);
                    }
                    uint8_t *d = dimg + (size_t)y * D.pitch + x4;
                    if (full) *(uint32_t *)d = o;
                    else for (int j = 0; x4 + j < dw; j++) d[j] = (uint8_t)(o >> (8 * j));
                }
            }
        }
        __syncthreads();         // the box is free for the load after next
    }
}

int launch_levels(const PlanDev &hp, const PlanDev *dp, const CUtensorMap &tmL0, const CUtensorMap *tmaps, int img0, Level0 l0, uint8_t *pyr, uint8_t *blur,
                  const ResizeTab *tabs, const int *lvtab, int B, int sm_count, cudaStream_t st, long *launches)
{
    static std::atomic<int> ctas_per_sm{0};
    int per = ctas_per_sm.load(std::memory_order_relaxed);
    if (!per) {
        HY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_level, LV_THREADS, 0));
        if (per < 1) per = 1;
        ctas_per_sm.store(per, std::memory_order_relaxed);
    }
    const int sms = sm_count > 0 ? sm_count : 148;
    for (int l = 0; l < hp.nlevels; l++) {
        const LevelDev &L = hp.lv[l];
        const long long nTiles = (long long)L.lvTilesX * L.lvTilesY * B;
        if (nTiles > 0x7fffffffLL) { set_error("too many level tiles in one batch"); return HYORB_EUNSUPPORTED; }
        const bool has_next = l + 1 < hp.nlevels;
        const int do_resize = has_next && !hp.lv[l + 1].area2x;
        const int grid = (int)std::min<long long>(nTiles, (long long)sms * per);
        if (l > 0)
            HY_CUDA(launch_dependent(k_level, dim3(grid), dim3(LV_THREADS), 0, st, dp, l, tmL0, tmaps, img0, pyr, blur, tabs, lvtab, (int)nTiles, do_resize));
        else
            k_level<<<grid, LV_THREADS, 0, st>>>(dp, l, tmL0, tmaps, img0, pyr, blur, tabs, lvtab, (int)nTiles, do_resize);
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
        const LevelDev &D = P.lv[l + 1];
        const ResizeTab *tx = plan->resize.data() + D.rsX, *ty = plan->resize.data() + D.rsY;
        L.lvDx = (int)plan->lvtab.size();
        for (int t = 0, d = 0; t <= L.lvTilesX; t++) {
            while (d < D.w && tx[d].ofs < t * LV_TW) d++;
            plan->lvtab.push_back(t == L.lvTilesX ? D.w : d);
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
