// blur.cu -- K4: per-level 7x7 Gaussian.  Replaces `cv::GaussianBlur(working_mat, working_mat, Size(7,7), 2, 2,
// BORDER_REFLECT_101)` on a borderless clone of each level (src/features/ORBExtractor.cpp:536-537).
// OpenCV's 8-bit path is fixed point: separable integer kernel [18,34,48,56,48,34,18] (sum 256), horizontal pass
// exact (Q8, fits 16 bits), vertical pass Q16, out = (v + 32768) >> 16 (pinned against cv2 by
// tests/test_oracle_vs_cv2.py through the oracle).
//
// HBM-bound streaming kernel, no shared memory: a thread owns 4 adjacent columns and walks down a band of rows.
// Per row it loads four aligned 32-bit words (the 12 source bytes at any row alignment), forms the four 7-tap horizontal sums with
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

// how a thread fetches the 12 source bytes x0-4 .. x0+7 of a row
enum { HS_WORDS = 0,      // interior: aligned-down 32-bit loads + funnel shift (any row alignment)
       HS_LEFT = 1,       // x0 == 0: bytes -4..-1 are the reflection of bytes 4..1
       HS_BYTES = 2 };    // right edge: per-byte gather through precomputed reflected column indices

// horizontal 7-tap sums of the 4 pixels x0..x0+3 of one source row
__device__ __forceinline__ void hsum4(const uint8_t *__restrict__ row, int x0, int mode, const int (&xi)[12], int (&h)[4])
{
    uint32_t w0, w1, w2;
    if (mode == HS_WORDS) {
        const uint8_t *a = row + x0 - 4;
        const unsigned mis = (unsigned)((uintptr_t)a & 3);
        const uint32_t *q = (const uint32_t *)(a - mis);
        const uint32_t q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
        w0 = __funnelshift_r(q0, q1, mis * 8); w1 = __funnelshift_r(q1, q2, mis * 8); w2 = __funnelshift_r(q2, q3, mis * 8);
    } else if (mode == HS_LEFT) {
        const unsigned mis = (unsigned)((uintptr_t)row & 3);
        const uint32_t *q = (const uint32_t *)(row - mis);     // mis > 0 only for rows > 0 or an unaligned base inside an allocation
        const uint32_t q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        w1 = __funnelshift_r(q0, q1, mis * 8); w2 = __funnelshift_r(q1, q2, mis * 8);
        w0 = __byte_perm(w1, w2, 0x1234);                      // REFLECT_101: pixels -4,-3,-2,-1 = pixels 4,3,2,1
    } else {
        w0 = w1 = w2 = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            w0 |= (uint32_t)row[xi[j]] << (8 * j);
            w1 |= (uint32_t)row[xi[4 + j]] << (8 * j);
            w2 |= (uint32_t)row[xi[8 + j]] << (8 * j);
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
    // aligned-down word loads read at most 3 bytes before x0-4 (still inside the row because x0 >= 4) and up to byte x0+11
    int mode = HS_BYTES;
    if (x0 >= 4 && x0 + 12 <= w && (x0 >= 8 || ((uintptr_t)img & 3) == 0 || b > 0)) mode = HS_WORDS;
    else if (x0 == 0 && (((uintptr_t)img & 3) == 0 || b > 0)) mode = HS_LEFT;     // w >= 62 always holds (tables.cu)
    int xi[12];
#pragma unroll
    for (int j = 0; j < 12; j++) xi[j] = reflect101(x0 - 4 + j, w);
    uint8_t *out = blur + (size_t)b * plan->pyrStride + L.off + x0;
    const int opitch = L.pitch;
    const int yEnd = min(y0 + BL_ROWS, h);

    int win[7][4];       // win[k] = horizontal sums of source row (y - 3 + k) relative to the current output row
    // prime rows y0-3 .. y0+2 into slots 0..5
#pragma unroll
    for (int k = 0; k < 6; k++) hsum4(img + (size_t)reflect101(y0 - 3 + k, h) * pitch, x0, mode, xi, win[k]);
    for (int yb = y0; yb < yEnd; yb += 7) {
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int y = yb + k;
            if (y < yEnd) {
                // the newest source row (y+3) goes into slot (6+k)%7; output row y then reads slots (k .. k+6)%7
                hsum4(img + (size_t)reflect101(y + 3, h) * pitch, x0, mode, xi, win[(6 + k) % 7]);
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
