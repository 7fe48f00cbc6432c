// TMA probe: which descriptor / address-space variants work on this box.  usage: probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <vector>
#include <dlfcn.h>
#include <string.h>
typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap *gm, int useGlobal, int x, int y, int z, int bytes, uint8_t *out)
{
    __shared__ __align__(128) uint8_t buf[8192];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *m = useGlobal ? gm : &pm;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(buf)), "l"(m), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(buf)), "l"(m), "r"(s32(&bar)), "r"(x), "r"(y) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int rank = (variant & 1) ? 2 : 3, useGlobal = (variant >> 1) & 1, bw = (variant & 4) ? 64 : 80, bh = (variant & 8) ? 64 : 70;
    const int w = 320, h = 240, n = 2, pitch = 320;
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
    uint8_t *img; cudaMalloc(&img, (size_t)pitch * h * n);
    std::vector<uint8_t> hi((size_t)pitch * h * n);
    for (size_t i = 0; i < hi.size(); i++) hi[i] = (uint8_t)(i * 7 + (i >> 8));
    cudaMemcpy(img, hi.data(), hi.size(), cudaMemcpyHostToDevice);
    if (variant & 16) { void *hd = dlopen("libcuda.so.1", RTLD_NOW); fn = hd ? dlsym(hd, "cuTensorMapEncodeTiled") : nullptr; printf("dlsym fn %p\n", fn); if (!fn) return 1; }
    alignas(64) CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n}, str[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, img, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d global %d box %dx%d encode -> %d\n", variant, rank, useGlobal, bw, bh, (int)r);
    for (int i = 0; i < 16; i++) printf("%016llx ", (unsigned long long)((uint64_t *)&tm)[i]);
    printf("\n");
    if (r) return 1;
    CUtensorMap *gm; cudaMalloc(&gm, sizeof(tm)); cudaMemcpy(gm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    uint8_t *out; cudaMalloc(&out, 8192); cudaMemset(out, 0xEE, 8192);
    const int x = argc > 2 ? atoi(argv[2]) : 77, y = argc > 3 ? atoi(argv[3]) : 15, z = rank == 3 ? 1 : 0, bytes = bw * bh;
    if (rank == 3) k<3><<<1, 128>>>(tm, gm, useGlobal, x, y, z, bytes, out); else k<2><<<1, 128>>>(tm, gm, useGlobal, x, y, z, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<uint8_t> ho(8192); cudaMemcpy(ho.data(), out, 8192, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw; xx++) {
        const int gx = x + xx, gy = y + yy;
        const uint8_t want = (gx < w && gy < h) ? hi[(size_t)z * pitch * h + (size_t)gy * pitch + gx] : 0;
        if (ho[yy * bw + xx] != want) bad++;
    }
    printf("mismatches %d of %d\n", bad, bw * bh);
    return 0;
}
