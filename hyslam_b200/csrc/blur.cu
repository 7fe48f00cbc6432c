// blur.cu -- K4: per-level 7x7 Gaussian.  Replaces `cv::GaussianBlur(working_mat, working_mat, Size(7,7), 2, 2,
// BORDER_REFLECT_101)` on a borderless clone of each level (src/features/ORBExtractor.cpp:536-537).
// OpenCV's 8-bit path is fixed point: separable integer kernel [18,34,48,56,48,34,18] (sum 256), horizontal pass
// exact (Q8, fits 16 bits), vertical pass Q16, out = (v + 32768) >> 16 (pinned against cv2 by
// tests/test_oracle_vs_cv2.py through the oracle).
// Mapping: one CTA per 64x16 output tile of one level of one image; the 70x22 source window is staged in shared
// memory with the reflection applied on load, the horizontal pass writes a 64x22 u16 tile, the vertical pass
// produces 4 adjacent outputs per thread and stores them as one 32-bit word.
#include "common.cuh"

namespace hyorb {

constexpr int BL_W = 64, BL_H = 16, BL_THREADS = 256;
constexpr int BL_SW = BL_W + 6, BL_SH = BL_H + 6;

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(BL_THREADS)
k_blur(const PlanDev *__restrict__ plan, Level0 l0, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur)
{
    __shared__ uint8_t s_src[BL_SH][BL_SW + 2];
    __shared__ uint16_t s_h[BL_SH][BL_W];

    const int tid = threadIdx.x, b = blockIdx.y;
    int l = 0;
    const int nl = plan->nlevels;
    while (l + 1 < nl && (int)blockIdx.x >= plan->blurTileBase[l + 1]) l++;
    const LevelDev &L = plan->lv[l];
    const int t = blockIdx.x - plan->blurTileBase[l];
    const int tilesX = (L.w + BL_W - 1) / BL_W;
    const int x0 = (t % tilesX) * BL_W, y0 = (t / tilesX) * BL_H;
    const uint8_t *img; int pitch;
    if (l == 0) { img = l0.base + (size_t)b * l0.stride; pitch = l0.pitch; }
    else { img = pyr + (size_t)b * plan->pyrStride + L.off; pitch = L.pitch; }
    const int w = L.w, h = L.h;

    for (int i = tid; i < BL_SH * BL_SW; i += BL_THREADS) {
        const int r = i / BL_SW, c = i - r * BL_SW;
        const int y = reflect101(y0 + r - 3, h), x = reflect101(x0 + c - 3, w);
        s_src[r][c] = img[(size_t)y * pitch + x];
    }
    __syncthreads();
    for (int i = tid; i < BL_SH * BL_W; i += BL_THREADS) {
        const int r = i / BL_W, c = i - r * BL_W;
        const uint8_t *p = &s_src[r][c];
        const int v = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
        s_h[r][c] = (uint16_t)v;
    }
    __syncthreads();
    {
        const int r = tid / (BL_W / 4), c4 = (tid - r * (BL_W / 4)) * 4;     // 16 rows x 16 groups
        const int y = y0 + r, x = x0 + c4;
        if (y < h && x < w) {
            uint32_t out = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = c4 + j;
                const uint32_t v = 18u * ((uint32_t)s_h[r][c] + s_h[r + 6][c]) + 34u * ((uint32_t)s_h[r + 1][c] + s_h[r + 5][c]) +
                                   48u * ((uint32_t)s_h[r + 2][c] + s_h[r + 4][c]) + 56u * (uint32_t)s_h[r + 3][c];
                out |= ((v + 32768u) >> 16) << (8 * j);
            }
            uint8_t *d = blur + (size_t)b * plan->pyrStride + L.off + (size_t)y * L.pitch + x;
            if (x + 3 < w) *(uint32_t *)d = out;
            else for (int j = 0; x + j < w; j++) d[j] = (uint8_t)(out >> (8 * j));
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
        tb += ((hp->lv[l].w + BL_W - 1) / BL_W) * ((hp->lv[l].h + BL_H - 1) / BL_H);
    }
    hp->blurTileBase[hp->nlevels] = tb;
}

}  // namespace hyorb
