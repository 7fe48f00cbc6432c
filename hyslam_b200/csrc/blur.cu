// blur.cu -- K4: per-level 7x7 Gaussian.  Replaces `cv::GaussianBlur(working_mat, working_mat, Size(7,7), 2, 2,
// BORDER_REFLECT_101)` on a borderless clone of each level (src/features/ORBExtractor.cpp:536-537).
// OpenCV's 8-bit path is fixed point: separable integer kernel [18,34,48,56,48,34,18] (sum 256), horizontal pass
// exact (Q8, fits 16 bits), vertical pass Q16, out = (v + 32768) >> 16 (pinned against cv2 by
// tests/test_oracle_vs_cv2.py through the oracle).
//
// HBM-bound streaming kernel, no shared memory: a thread owns 4 adjacent columns and walks down a band of rows.
// Per row it loads three aligned 32-bit words (12 source bytes), forms the four 7-tap horizontal sums with
// byte-funnel shifts + two DP4A each, and keeps the last 7 rows of horizontal sums in registers (rotating window,
// resolved at compile time by unrolling 7 rows); the vertical 7-tap sum of each column is then 3 adds + 4 IMADs and
// four results leave as one 32-bit store.  Reflection (BORDER_REFLECT_101) is index arithmetic on the row number and,
// for the few threads whose 12-byte window crosses the image's left/right edge, a per-byte gather.
#include "common.cuh"

namespace hyorb {

constexpr int BL_THREADS = 128;          // 512 columns per CTA
constexpr int BL_ROWS = 56;              // output rows per CTA (multiple of 7: the register window rotates with period 7)

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// horizontal 7-tap sums of the 4 pixels x0..x0+3 of one source row
__device__ __forceinline__ void hsum4(const uint8_t *__restrict__ row, int x0, int w, bool fast, int (&h)[4])
{
    uint32_t w0, w1, w2;
    if (fast) {           // 4-aligned row, window x0-4 .. x0+7 inside the image
        const uint32_t *p = (const uint32_t *)(row + x0);
        w0 = p[-1]; w1 = p[0]; w2 = p[1];
    } else {
        w0 = w1 = w2 = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            w0 |= (uint32_t)row[reflect101(x0 - 4 + j, w)] << (8 * j);
            w1 |= (uint32_t)row[reflect101(x0 + j, w)] << (8 * j);
            w2 |= (uint32_t)row[reflect101(x0 + 4 + j, w)] << (8 * j);
        }
    }
    const uint32_t G0 = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps -3..0
    const uint32_t G1 = 48u | (34u << 8) | (18u << 16);                   // taps +1..+3
    // output j needs bytes (x0+j-3 .. x0+j+3) = offsets j+1 .. j+7 of w0:w1:w2
    h[0] = (int)__dp4a(__funnelshift_r(w0, w1, 8), G0, __dp4a(__funnelshift_r(w1, w2, 8), G1, 0u));
    h[1] = (int)__dp4a(__funnelshift_r(w0, w1, 16), G0, __dp4a(__funnelshift_r(w1, w2, 16), G1, 0u));
    h[2] = (int)__dp4a(__funnelshift_r(w0, w1, 24), G0, __dp4a(__funnelshift_r(w1, w2, 24), G1, 0u));
    h[3] = (int)__dp4a(w1, G0, __dp4a(w2, G1, 0u));
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
    const int x0 = ((t % bandsX) * BL_THREADS + threadIdx.x) * 4;
    const int y0 = (t / bandsX) * BL_ROWS;
    if (x0 >= w) return;
    const uint8_t *img; int pitch;
    if (l == 0) { img = l0.base + (size_t)b * l0.stride; pitch = l0.pitch; }
    else { img = pyr + (size_t)b * plan->pyrStride + L.off; pitch = L.pitch; }
    const bool fast = (((uintptr_t)img | (uintptr_t)pitch) & 3) == 0 && x0 >= 4 && x0 + 8 <= w;
    uint8_t *out = blur + (size_t)b * plan->pyrStride + L.off + x0;
    const int opitch = L.pitch;
    const int yEnd = min(y0 + BL_ROWS, h);

    int win[7][4];       // win[k] = horizontal sums of source row (y - 3 + k) relative to the current output row
    // prime rows y0-3 .. y0+2 into slots 0..5
#pragma unroll
    for (int k = 0; k < 6; k++) hsum4(img + (size_t)reflect101(y0 - 3 + k, h) * pitch, x0, w, fast, win[k]);
    for (int yb = y0; yb < yEnd; yb += 7) {
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int y = yb + k;
            if (y < yEnd) {
                // the newest source row (y+3) goes into slot (6+k)%7; output row y then reads slots (k .. k+6)%7
                hsum4(img + (size_t)reflect101(y + 3, h) * pitch, x0, w, fast, win[(6 + k) % 7]);
                uint32_t o = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t v = 18u * (uint32_t)(win[k % 7][j] + win[(k + 6) % 7][j]) + 34u * (uint32_t)(win[(k + 1) % 7][j] + win[(k + 5) % 7][j]) +
                                       48u * (uint32_t)(win[(k + 2) % 7][j] + win[(k + 4) % 7][j]) + 56u * (uint32_t)win[(k + 3) % 7][j];
                    o |= ((v + 32768u) >> 16) << (8 * j);
                }
                uint8_t *d = out + (size_t)y * opitch;
                if (x0 + 3 < w) *(uint32_t *)d = o;       // level offsets and pitches are multiples of 16
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
