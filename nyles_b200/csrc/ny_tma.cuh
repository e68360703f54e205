// TMA (cp.async.bulk.tensor) and mbarrier helpers for sm_100a, and the host-side tensor-map encoder.
//
// Level / field arrays are dense fp64 boxes F[k][j][i]; a kernel stages (rows x columns) tiles of one
// plane in shared memory with ONE bulk-tensor copy per plane issued by one thread.  Out-of-range box
// coordinates (negative, or beyond the array) are legal: the hardware fills those elements with
// zeros, which is what gives the stencil kernels their branch-free edge tiles.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

// rank-3 fp64 tensor map over a dense (nz, ny, nx) array, box = (bz, by, bx) elements, no swizzle,
// zero fill.  Needs: base 16-byte aligned, nx even (row pitch multiple of 16 bytes), bx even, each
// box extent <= 256.  At run time the COLUMN coordinate of a box must be even as well (the box has to
// start on a 16-byte boundary; an odd column raises "illegal instruction" on sm_100 -- measured with
// tools/probe/tma_probe.cu); row and plane coordinates are free, and all three may lie outside the
// array.  Returns 0 on success (error text through ny_set_error).
int ny_tma_encode_3d(CUtensorMap* map, const double* base, int nx, int ny, int nz, int bx, int by, int bz);

#ifdef __CUDACC__
namespace nytma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// order this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one (non-blocking) arrival of the calling thread
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// one plane tile: box origin (c0 = column, c1 = row, c2 = plane), completes `bytes of the box` on bar
__device__ __forceinline__ void load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
// pull the box into L2 only
__device__ __forceinline__ void prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace nytma
#endif
