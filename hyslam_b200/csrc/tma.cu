// tma.cu -- host side of tma.cuh: tensor-map encoding via cudaGetDriverEntryPoint (libhyorb links only the static
// CUDA runtime), plus the repack kernel that gives level 0 a TMA-compatible layout when the caller's is not.
#include <mutex>

#include "common.cuh"
#include "tma.cuh"

namespace hyorb {

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn g_encode = nullptr;
static std::once_flag g_encode_once;

int tma_encode_u8_3d(CUtensorMap *out, const void *base, int w, int h, int n, size_t pitch, size_t image_stride, int box_w, int box_h)
{
    std::call_once(g_encode_once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            g_encode = (encode_tiled_fn)fn;
    });
    if (!g_encode) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return HYORB_ECUDA; }
    if (!tma_compatible(base, pitch, image_stride) || (box_w & 15) || box_w > 256 || box_h > 256) {
        set_error("tensor map: base %p pitch %zu image stride %zu box %dx%d violates the TMA alignment rules", base, pitch, image_stride, box_w, box_h);
        return HYORB_EINVAL;
    }
    // a single image still needs a valid (multiple of 16) third stride
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(n < 1 ? 1 : n)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)image_stride};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%dx%dx%d, pitch %zu, stride %zu, box %dx%d) -> CUresult %d", w, h, n, pitch, image_stride, box_w, box_h, (int)r); return HYORB_ECUDA; }
    return HYORB_OK;
}

// rows of arbitrary alignment -> rows of a 16-byte aligned pitch; 4 bytes per thread, 32-bit stores
__global__ void __launch_bounds__(256)
k_repack(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride, uint8_t *__restrict__ dst, int dpitch,
         unsigned long long dstride, int w, int h)
{
    const int x = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x >= dpitch) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sstride;
    uint8_t *d = dst + (size_t)blockIdx.z * dstride;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const uint8_t *p = s + (size_t)y * spitch + x;
        uint32_t v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) if (x + j < w) v |= (uint32_t)p[j] << (8 * j);
        *(uint32_t *)(d + (size_t)y * dpitch + x) = v;
    }
}

int launch_repack(const uint8_t *src, int spitch, size_t sstride, uint8_t *dst, int dpitch, size_t dstride, int w, int h, int B, cudaStream_t st, long *launches)
{
    dim3 grd((dpitch / 4 + 255) / 256, h < 64 ? h : 64, B);
    k_repack<<<grd, 256, 0, st>>>(src, spitch, sstride, dst, dpitch, dstride, w, h);
    ++*launches;
    HY_CUDA(cudaGetLastError());
    return HYORB_OK;
}

}  // namespace hyorb
