// preprocess.cu -- A.0: the camera-image preparation in front of the extractor.
//
// Replaces ImageProcessing::PreProcessImg (src/main/ImageProcessing.cpp:118-138): cv::resize(img, img, Size(), fscale,
// fscale) followed by cvtColor(RGB|BGR[A] -> GRAY), for the two scales the reference's camera configurations use
// (pinned against cv2 by tests/test_oracle_vs_cv2.py):
//   * scale 1.0: the resize is a copy;
//   * scale 0.5: OpenCV takes the INTER_AREA fast path, a per-channel 2x2 box mean (a+b+c+d+2)>>2 over
//     dst = cvRound(src * 0.5) pixels;
//   * gray = (R*9798 + G*19235 + B*3735 + 16384) >> 15 (15-bit fixed-point BT.601 weights); alpha is ignored.
// One thread produces 4 horizontally adjacent gray pixels (one 32-bit store when the destination allows it); the source
// bytes of those pixels are contiguous (4 / 8 pixels x CN bytes per source row) and are fetched with 32-bit loads when the
// row is 4-byte aligned.  HBM-bound by construction: every source byte is read once, every gray byte written once.
#include "common.cuh"

namespace hyorb {

// N contiguous bytes (N a multiple of 4) starting at p -> words; `aligned` is uniform per row
template <int N>
__device__ __forceinline__ void load_run(const uint8_t *__restrict__ p, bool aligned, uint32_t (&w)[N / 4])
{
    if (aligned) {
#pragma unroll
        for (int i = 0; i < N / 4; i++) w[i] = __ldg((const uint32_t *)p + i);
    } else {
#pragma unroll
        for (int i = 0; i < N / 4; i++)
            w[i] = (uint32_t)__ldg(p + 4 * i) | ((uint32_t)__ldg(p + 4 * i + 1) << 8) | ((uint32_t)__ldg(p + 4 * i + 2) << 16) | ((uint32_t)__ldg(p + 4 * i + 3) << 24);
    }
}
template <int N>
__device__ __forceinline__ int byte_of(const uint32_t (&w)[N], int i) { return (int)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu); }

__device__ __forceinline__ int to_gray(int c0, int c1, int c2, int rgb)
{
    const int r = rgb ? c0 : c2, b = rgb ? c2 : c0;
    return (r * 9798 + c1 * 19235 + b * 3735 + 16384) >> 15;
}

template <int CN, bool HALF>
__global__ void __launch_bounds__(256)
k_preprocess(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride, uint8_t *__restrict__ dst, int dpitch,
             unsigned long long dstride, int ow, int oh, int rgb)
{
    const int x = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x >= ow) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sstride;
    uint8_t *d = dst + (size_t)blockIdx.z * dstride;
    const bool dal = ((((uintptr_t)d) | (unsigned)dpitch) & 3) == 0;
    const bool full = x + 4 <= ow;
    constexpr int SPX = HALF ? 8 : 4;                // source pixels per row behind 4 gray pixels
    constexpr int NB = SPX * CN;                     // bytes per source row
    for (int y = blockIdx.y; y < oh; y += gridDim.y) {
        int g[4] = {0, 0, 0, 0};
        const uint8_t *r0 = s + (size_t)(HALF ? 2 * y : y) * spitch + (size_t)(HALF ? 2 * x : x) * CN;
        if (full) {
            uint32_t a[NB / 4];
            load_run<NB>(r0, (((uintptr_t)r0) & 3) == 0, a);
            if (HALF) {
                uint32_t b[NB / 4];
                load_run<NB>(r0 + spitch, (((uintptr_t)(r0 + spitch)) & 3) == 0, b);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int c[3] = {0, 0, 0};
#pragma unroll
                    for (int k = 0; k < (CN == 1 ? 1 : 3); k++)
                        c[k] = (byte_of(a, 2 * j * CN + k) + byte_of(a, (2 * j + 1) * CN + k) + byte_of(b, 2 * j * CN + k) + byte_of(b, (2 * j + 1) * CN + k) + 2) >> 2;
                    g[j] = CN == 1 ? c[0] : to_gray(c[0], c[1], c[2], rgb);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    g[j] = CN == 1 ? byte_of(a, j) : to_gray(byte_of(a, j * CN), byte_of(a, j * CN + 1), byte_of(a, j * CN + 2), rgb);
            }
        } else {                                      // the last, partial group of a row: byte loads, nothing past pixel ow - 1
            for (int j = 0; x + j < ow; j++) {
                int c[3] = {0, 0, 0};
                for (int k = 0; k < (CN == 1 ? 1 : 3); k++) {
                    if (HALF) {
                        const uint8_t *p = r0 + 2 * j * CN + k;
                        c[k] = ((int)__ldg(p) + (int)__ldg(p + CN) + (int)__ldg(p + spitch) + (int)__ldg(p + spitch + CN) + 2) >> 2;
                    } else c[k] = __ldg(r0 + j * CN + k);
                }
                g[j] = CN == 1 ? c[0] : to_gray(c[0], c[1], c[2], rgb);
            }
        }
        uint8_t *o = d + (size_t)y * dpitch + x;
        if (dal && x + 4 <= ow) *(uint32_t *)o = (uint32_t)g[0] | ((uint32_t)g[1] << 8) | ((uint32_t)g[2] << 16) | ((uint32_t)g[3] << 24);
        else
            for (int j = 0; j < 4 && x + j < ow; j++) o[j] = (uint8_t)g[j];
    }
}

int preprocess_size(int w, int h, int half_scale, int *ow, int *oh)
{
    // cvRound(src * 0.5): ties to even, like cv::resize's dsize computation
    *ow = half_scale ? (int)lrint((double)w * 0.5) : w;
    *oh = half_scale ? (int)lrint((double)h * 0.5) : h;
    if (w < 1 || h < 1 || *ow < 1 || *oh < 1) { set_error("preprocess: empty image"); return HYORB_EINVAL; }
    if (half_scale && (2 * *ow > w || 2 * *oh > h)) {
        set_error("preprocess: %dx%d at scale 0.5 rounds up to %dx%d; the box path needs 2*dst <= src", w, h, *ow, *oh);
        return HYORB_EUNSUPPORTED;
    }
    return HYORB_OK;
}

int launch_preprocess(const uint8_t *src, int spitch, size_t sstride, int w, int h, int channels, int rgb_order, int half_scale, uint8_t *dst, int dpitch,
                      size_t dstride, int B, cudaStream_t st, long *launches)
{
    int ow, oh;
    HY_TRY(preprocess_size(w, h, half_scale, &ow, &oh));
    if (channels != 1 && channels != 3 && channels != 4) { set_error("preprocess: %d channels (1, 3 or 4)", channels); return HYORB_EINVAL; }
    if (spitch < w * channels || dpitch < ow || B < 1 || B > 65535) { set_error("preprocess: bad pitch or batch"); return HYORB_EINVAL; }
    dim3 grd(((ow + 3) / 4 + 255) / 256, oh < 128 ? oh : 128, B);
#define HY_PP(CN, HALF) k_preprocess<CN, HALF><<<grd, 256, 0, st>>>(src, spitch, (unsigned long long)sstride, dst, dpitch, (unsigned long long)dstride, ow, oh, rgb_order)
    if (half_scale) { if (channels == 1) HY_PP(1, true); else if (channels == 3) HY_PP(3, true); else HY_PP(4, true); }
    else { if (channels == 1) HY_PP(1, false); else if (channels == 3) HY_PP(3, false); else HY_PP(4, false); }
#undef HY_PP
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
