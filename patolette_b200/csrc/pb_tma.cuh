// pb_tma.cuh - Hopper/Blackwell bulk asynchronous copies (TMA, non-tensor form) and the mbarrier they signal.
// 1-D `cp.async.bulk`: both addresses 16-byte aligned, size a multiple of 16 bytes; one thread issues a copy of
// kilobytes that the TMA unit carries out while the SM does something else (SASS: UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t pb_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// writes done with ordinary stores become visible to the async proxy (the TMA unit) - every writing thread, before
// the barrier that precedes the copy
__device__ __forceinline__ void pb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared -> global, completion tracked by bulk groups
__device__ __forceinline__ void pb_bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(pb_smem_addr(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the sources of all committed groups have been read (shared memory may be reused / the CTA may exit)
__device__ __forceinline__ void pb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// global -> shared, completion signalled on an mbarrier (transaction bytes)
__device__ __forceinline__ void pb_mbar_init(unsigned long long *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pb_smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pb_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pb_bulk_load(void *sdst, const void *gsrc, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pb_smem_addr(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(pb_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void pb_mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(pb_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
