// pyramid.cu -- K1: scale pyramid.  Replaces ORBExtractor::ComputePyramid (src/features/ORBExtractor.cpp:564-589):
// level l = cv::resize(level l-1, INTER_LINEAR) with OpenCV's 8-bit fixed-point bilinear kernel (Q11 coefficients,
// (b*(h>>4))>>16 vertical combine, (+2)>>2).  The 19-px reflected border the reference adds is never read
// downstream and is not materialised.
// HBM-bound streaming kernel: one thread produces 4 adjacent destination pixels (one 32-bit store); one launch
// per level covers the whole batch.
#include "common.cuh"

namespace hyorb {

__global__ void __launch_bounds__(256)
k_resize(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride, int sw, int sh,
         uint8_t *__restrict__ dst, int dpitch, unsigned long long dstride, int dw, int dh,
         const ResizeTab *__restrict__ tx, const ResizeTab *__restrict__ ty, int area2x)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= dw || y >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sstride;
    uint8_t *d = dst + (size_t)blockIdx.z * dstride + (size_t)y * dpitch + x4;
    uint32_t out = 0;
    if (area2x) {
        // cv::resize silently takes the INTER_AREA 2x2 mean when the ratio is exactly 2
        const uint8_t *r0 = s + (size_t)(2 * y) * spitch, *r1 = r0 + spitch;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = x4 + j;
            if (x < dw) out |= (uint32_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2) << (8 * j);
        }
    } else {
        const ResizeTab vy = ty[y];
        const int sy0 = min(max(vy.ofs, 0), sh - 1), sy1 = min(max(vy.ofs + 1, 0), sh - 1);
        const uint8_t *r0 = s + (size_t)sy0 * spitch, *r1 = s + (size_t)sy1 * spitch;
        const int b0 = vy.c0, b1 = vy.c1;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = x4 + j;
            if (x < dw) {
                const ResizeTab vx = tx[x];
                const int s0 = vx.ofs, s1 = min(s0 + 1, sw - 1);
                const int h0 = (int)r0[s0] * vx.c0 + (int)r0[s1] * vx.c1;
                const int h1 = (int)r1[s0] * vx.c0 + (int)r1[s1] * vx.c1;
                int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                v = min(max(v, 0), 255);
                out |= (uint32_t)v << (8 * j);
            }
        }
    }
    if (x4 + 3 < dw) *(uint32_t *)d = out;     // pitch and level offsets are multiples of 16
    else for (int j = 0; x4 + j < dw; j++) d[j] = (uint8_t)(out >> (8 * j));
}

int launch_pyramid(const PlanDev &hp, const PlanDev *, Level0 l0, uint8_t *pyr, const ResizeTab *tabs, int B, cudaStream_t st, long *launches)
{
    for (int l = 1; l < hp.nlevels; l++) {
        const LevelDev &S = hp.lv[l - 1], &D = hp.lv[l];
        const uint8_t *src = (l == 1) ? l0.base : pyr + S.off;
        const int spitch = (l == 1) ? l0.pitch : S.pitch;
        const unsigned long long sstride = (l == 1) ? l0.stride : hp.pyrStride;
        dim3 blk(64, 4), grd((D.w + 4 * 64 - 1) / (4 * 64), (D.h + 3) / 4, B);
        k_resize<<<grd, blk, 0, st>>>(src, spitch, sstride, S.w, S.h, pyr + D.off, D.pitch, hp.pyrStride, D.w, D.h,
                                      tabs + D.rsX, tabs + D.rsY, D.area2x);
        ++*launches;
    }
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
