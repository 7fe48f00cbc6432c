// blur.cu -- K4: per-level 7x7 Gaussian.  Replaces `cv::GaussianBlur(working_mat, working_mat, Size(7,7), 2, 2,
// BORDER_REFLECT_101)` on a borderless clone of each level (src/features/ORBExtractor.cpp:536-537).
// OpenCV's 8-bit path is fixed point: separable integer kernel [18,34,48,56,48,34,18] (sum 256), horizontal pass
// exact (Q8, fits 16 bits), vertical pass Q16, out = (v + 32768) >> 16 (pinned against cv2 by
// tests/test_oracle_vs_cv2.py through the oracle).
//
// Streaming kernel, no shared memory: a thread owns 4 adjacent columns and walks down a band of rows.  Every plane it
// reads has 16-byte aligned rows (pyramid layout; level 0 is TMA-addressable by construction, api.cu), so a row's 12
// source bytes are three aligned 32-bit loads.  The four 7-tap horizontal sums are byte-funnel shifts + two DP4A each;
// the last 7 rows of horizontal sums stay in registers (rotating window, resolved at compile time by unrolling 7
// rows); the vertical 7-tap sum of each column is 3 adds + 4 IMADs and four results leave as one 32-bit store.
// Reflection (BORDER_REFLECT_101) is branch-free index arithmetic on the row number and, for the few threads whose
// 12-byte window crosses the image's left/right edge, a byte permute / per-byte gather.
#include "common.cuh"

namespace hyorb {

constexpr int BL_THREADS = 128;          // 512 columns per CTA
constexpr int BL_ROWS = 56;              // output rows per CTA (multiple of 7: the register window rotates with period 7)

// BORDER_REFLECT_101 index for |overshoot| < n (always true here: <= 10 rows against levels of >= 62 rows), without a loop
__device__ __forceinline__ int reflect101_near(int i, int n)
{
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

// The 12 source bytes x0-4 .. x0+7 of one row (window bytes 0..11) are three aligned 32-bit loads.  Threads whose window
// crosses the image's left or right edge then rebuild the out-of-image bytes from their BORDER_REFLECT_101 sources with
// byte permutes whose selectors are per-thread constants (the sources always lie inside the same 12-byte window):
//   left  (x0 == 0): window bytes 0..3 (pixels -4..-1) = pixels 4,3,2,1 = window bytes 8,7,6,5;
//   right (e = w - (x0-4) < 12 is the first out-of-image window byte): byte i >= e takes byte 2(e-1) - i.  Only bytes up
//   to e+2 can reach a stored output; their sources lie in bytes 0..7 for word 1 and, for word 2, in bytes 4..11 when
//   e >= 8 and in bytes 0..7 otherwise.
struct EdgeFix {
    uint32_t sel1, sel2;      // PRMT selectors of window words 1 and 2
    int edge;                 // 0 interior, 1 left, 2 right
    int pair2;                // word 2 is permuted out of (w1,w2) [1] or (w0,w1) [0]
    int ld2;                  // window word 2 lies inside the row pitch
};
__device__ __forceinline__ EdgeFix edge_fix(int x0, int w, int pitch)
{
    EdgeFix f;
    f.sel1 = 0x7654u; f.sel2 = 0x7654u; f.pair2 = 1; f.edge = 0;
    f.ld2 = x0 + 8 <= pitch;
    if (x0 == 0) f.edge = 1;
    const int e = w - (x0 - 4);
    if (e < 12) {
        f.edge = 2;
        f.pair2 = e >= 8;
        uint32_t s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i1 = 4 + k, i2 = 8 + k;
            int a = i1 < e ? i1 : 2 * (e - 1) - i1;          // source byte inside bytes 0..7
            a = a < 0 ? 0 : a;                                 // bytes past e+2 are never used: just keep the selector legal
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

struct Row3 { uint32_t w0, w1, w2; };
__device__ __forceinline__ Row3 load_row(const uint8_t *__restrict__ row, int x0, const EdgeFix &f)
{
    Row3 r;
    const uint32_t *q = (const uint32_t *)(row + x0) - 1;      // the row's aligned word that holds pixel x0-4
    r.w0 = x0 > 0 ? __ldg(q) : 0u;
    r.w1 = __ldg(q + 1);
    r.w2 = f.ld2 ? __ldg(q + 2) : 0u;
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
// horizontal 7-tap sums of the 4 pixels x0..x0+3
__device__ __forceinline__ void hsum4(const Row3 &r, int (&h)[4])
{
    const uint32_t G0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps -3..0
    const uint32_t G1 = 48u | (34u << 8) | (18u << 16);                   // taps +1..+3
    // output j needs bytes (x0+j-3 .. x0+j+3) = offsets j+1 .. j+7 of w0:w1:w2
    h[0] = (int)__dp4a(__funnelshift_r(r.w0, r.w1, 8), G0, __dp4a(__funnelshift_r(r.w1, r.w2, 8), G1, 0u));
    h[1] = (int)__dp4a(__funnelshift_r(r.w0, r.w1, 16), G0, __dp4a(__funnelshift_r(r.w1, r.w2, 16), G1, 0u));
    h[2] = (int)__dp4a(__funnelshift_r(r.w0, r.w1, 24), G0, __dp4a(__funnelshift_r(r.w1, r.w2, 24), G1, 0u));
    h[3] = (int)__dp4a(r.w1, G0, __dp4a(r.w2, G1, 0u));
}

__global__ void __launch_bounds__(BL_THREADS)
k_blur(const PlanDev *__restrict__ plan, Level0 l0, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur)
{
    const int b = blockIdx.y;
    int l = 0;
    const int nl = plan->nlevels;
    while (l + 1 < nl && (int)blockIdx.x >= plan->blurTileBase[l + 1]) l++;
    const LevelDev &L = plan->lv[l];
    const int t = blockIdx.x - plan->blurTileBase[l];
    const int w = L.w, h = L.h;
    const int bandsX = (w + 4 * BL_THREADS - 1) / (4 * BL_THREADS);
    const int bx = t % bandsX;
    const int x0 = (bx * BL_THREADS + threadIdx.x) * 4;
    const int y0 = (t / bandsX) * BL_ROWS;
    if (x0 >= w) return;
    const uint8_t *img; int pitch;
    if (l == 0) { img = l0.base + (size_t)b * l0.stride; pitch = l0.pitch; }
    else { img = pyr + (size_t)b * plan->pyrStride + L.off; pitch = L.pitch; }
    const EdgeFix fx = edge_fix(x0, w, pitch);
    uint8_t *out = blur + (size_t)b * plan->pyrStride + L.off + x0;
    const int opitch = L.pitch;
    const int yEnd = min(y0 + BL_ROWS, h);
    const bool full = x0 + 3 < w;

    int win[7][4];       // win[k] = horizontal sums of source row (y - 3 + k) relative to the current output row
    // prime rows y0-3 .. y0+2 into slots 0..5 (six independent loads in flight)
    {
        Row3 pr[6];
#pragma unroll
        for (int k = 0; k < 6; k++) pr[k] = load_row(img + (size_t)reflect101_near(y0 - 3 + k, h) * pitch, x0, fx);
#pragma unroll
        for (int k = 0; k < 6; k++) hsum4(pr[k], win[k]);
    }
    // software pipeline: the words of source row y+4 are requested before row y+3 is consumed
    Row3 cur = load_row(img + (size_t)reflect101_near(y0 + 3, h) * pitch, x0, fx);
    for (int yb = y0; yb < yEnd; yb += 7) {
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int y = yb + k;
            // the newest source row (y+3) goes into slot (6+k)%7; output row y then reads slots (k .. k+6)%7.  Rows past
            // yEnd of the last band are computed from in-range (reflected) rows and never stored.
            const Row3 nxt = load_row(img + (size_t)reflect101_near(y + 4, h) * pitch, x0, fx);
            hsum4(cur, win[(6 + k) % 7]);
            cur = nxt;
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t v = 18u * (uint32_t)(win[k % 7][j] + win[(k + 6) % 7][j]) + 34u * (uint32_t)(win[(k + 1) % 7][j] + win[(k + 5) % 7][j]) +
                                   48u * (uint32_t)(win[(k + 2) % 7][j] + win[(k + 4) % 7][j]) + 56u * (uint32_t)win[(k + 3) % 7][j];
                o |= ((v + 32768u) >> 16) << (8 * j);
            }
            if (y < yEnd) {
                uint8_t *d = out + (size_t)y * opitch;
                if (full) *(uint32_t *)d = o;       // level offsets and pitches are multiples of 16
                else for (int j = 0; x0 + j < w; j++) d[j] = (uint8_t)(o >> (8 * j));
            }
        }
    }
}

int launch_blur(const PlanDev &hp, const PlanDev *dp, Level0 l0, const uint8_t *pyr, uint8_t *blur, int B, cudaStream_t st, long *launches)
{
    dim3 grd(hp.blurTileBase[hp.nlevels], B);
    k_blur<<<grd, BL_THREADS, 0, st>>>(dp, l0, pyr, blur);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

void blur_tiles(PlanDev *hp)
{
    int tb = 0;
    for (int l = 0; l < hp->nlevels; l++) {
        hp->blurTileBase[l] = tb;
        tb += ((hp->lv[l].w + 4 * BL_THREADS - 1) / (4 * BL_THREADS)) * ((hp->lv[l].h + BL_ROWS - 1) / BL_ROWS);
    }
    hp->blurTileBase[hp->nlevels] = tb;
}

}  // namespace hyorb
