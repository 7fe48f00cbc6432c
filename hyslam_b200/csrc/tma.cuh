// tma.cuh -- Tensor Memory Accelerator plumbing (sm_100a): host-side tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda) and the device-side mbarrier / cp.async.bulk.tensor wrappers the image kernels use.
//
// Every 8-bit image plane the kernels read (level 0, pyramid levels, blurred levels) is described as a 3-D tensor
// {x: width, y: height, z: image index in the batch} with byte strides {pitch, bytes per image}.  TMA needs a 16-byte
// aligned base and strides that are multiples of 16; api.cu guarantees that for its own buffers and repacks caller
// images that are not.  Out-of-range box elements are zero-filled by the hardware, which is exactly what the staging
// code of the kernels wants at the right/bottom image edges.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hyorb {

// host: encode a u8 {w, h, n} tensor with a {box_w, box_h, 1} box; returns a HYORB status
int tma_encode_u8_3d(CUtensorMap *out, const void *base, int w, int h, int n, size_t pitch, size_t image_stride, int box_w, int box_h);
inline bool tma_compatible(const void *base, size_t pitch, size_t image_stride)
{
    return (((uintptr_t)base | pitch | image_stride) & 15) == 0;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make freshly initialised barriers visible to the async proxy (TMA unit)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's (and, after a CTA barrier, the CTA's) generic-proxy shared-memory accesses before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// a tensor map that lives in global memory and was (re)written since an earlier kernel used that address: make the
// TMA unit's descriptor fetch observe the new bytes (handles are created and destroyed, the allocator reuses addresses)
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap *map)
{
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(map) : "memory");
}
// one box of a 3-D tensor -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
#endif

}  // namespace hyorb
