// bulk.cuh — 1-D bulk asynchronous copies (the TMA engine without a tensor map: cp.async.bulk, SASS UBLKCP) and the
// mbarrier handshake around them, for the CSR streaming kernels (K1 major_stats.cu, K6 pca.cu). A producer thread streams a
// contiguous run of the nnz arrays tile by tile into a shared-memory ring; the copy engine signals the stage's "full"
// barrier with the byte count, consumer warps signal "empty" when they are done with the stage. The loads of many tiles
// are in flight without occupying a register or an issue slot of the consumer warps.
#pragma once
#include <cstdint>

namespace srb {
namespace bulk {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (then __syncthreads before first use)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is counted on `bar`
__device__ __forceinline__ void copy_g2s(uint32_t dst_smem, const void *src_global, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_global), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace bulk
}  // namespace srb
