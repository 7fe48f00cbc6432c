// programming-guide style TMA sample (libcu++ wrappers) + 1-D bulk copy.  usage: probe2 <0|1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int SW = 64, SH = 64;
__global__ void k2d(const __grid_constant__ CUtensorMap tm, int x, int y, uint8_t *out)
{
    __shared__ alignas(128) uint8_t buf[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&buf, &tm, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) out[i] = ((uint8_t *)buf)[i];
}
__global__ void k1d(const uint8_t *src, uint8_t *out)
{
    __shared__ alignas(128) uint8_t buf[4096];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cuda::device::experimental::cp_async_bulk_global_to_shared(buf, src, sizeof(buf), bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int w = 320, h = 240, pitch = 320;
    cudaFree(0);
    uint8_t *img; cudaMalloc(&img, (size_t)pitch * h);
    std::vector<uint8_t> hi((size_t)pitch * h);
    for (size_t i = 0; i < hi.size(); i++) hi[i] = (uint8_t)(i * 7 + (i >> 8));
    cudaMemcpy(img, hi.data(), hi.size(), cudaMemcpyHostToDevice);
    uint8_t *out; cudaMalloc(&out, 8192); cudaMemset(out, 0xEE, 8192);
    std::vector<uint8_t> ho(8192);
    if (variant == 1) {
        k1d<<<1, 128>>>(img, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("1-D bulk kernel -> %s\n", cudaGetErrorString(e));
        if (e) return 1;
        cudaMemcpy(ho.data(), out, 8192, cudaMemcpyDeviceToHost);
        printf("1-D mismatches %d\n", memcmp(ho.data(), hi.data(), 4096) != 0);
        return 0;
    }
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
    alignas(64) CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, str[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    const int X = argc > 2 ? atoi(argv[2]) : 32, Y = argc > 3 ? atoi(argv[3]) : 16;
    k2d<<<1, 128>>>(tm, X, Y, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("2-D tensor kernel -> %s\n", cudaGetErrorString(e));
    if (e) return 1;
    cudaMemcpy(ho.data(), out, 8192, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < SH; yy++) for (int xx = 0; xx < SW; xx++) bad += ho[yy * SW + xx] != hi[(size_t)(Y + yy) * pitch + X + xx];
    printf("2-D mismatches %d\n", bad);
    return 0;
}
