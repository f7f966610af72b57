// TMA (cp.async.bulk.tensor) tile loads and the mbarrier that signals their arrival, as inline PTX for sm_100a.
// Used by the fused pyramid (pyramid.cu) and the FAST kernel (fast.cu) to stage image tiles with their halos in shared memory:
// one elected thread issues the copy, nobody spends instructions on addresses or loads, out-of-image parts are zero filled.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace gvs {

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// one 3-D tile (x bytes, y rows, z image) of a tensor map into shared memory; completion is signalled on the mbarrier
__device__ __forceinline__ void tmaLoadTile3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

}  // namespace gvs
