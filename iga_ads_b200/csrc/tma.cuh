// tma.cuh -- thin wrappers over the sm_100a async-copy PTX used by the TMA-fed kernels: shared-memory
// mbarriers, bulk copies (cp.async.bulk, SASS UBLKCP) and their completion mechanisms.
#ifndef ADSB_TMA_CUH
#define ADSB_TMA_CUH

#include <cstdint>

namespace adsb {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const void* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(smem_addr(src))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}


// Programmatic dependent launch (griddepcontrol): a kernel launched with launch_ex(..., pdl = true) may start
// while its predecessor in the stream is still draining -- its prologue (barrier init, staging of constant
// tables) overlaps the predecessor's tail -- and must call pdl_wait() before it touches any memory the
// predecessor reads or writes.  pdl_launch() in the predecessor lets the dependent grid be scheduled from
// that point on (otherwise: when the predecessor's blocks exit).  Both are no-ops in an ordinary launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace adsb

#endif
