// tma.cuh -- mbarrier + TMA bulk-copy (cp.async.bulk) helpers, inline PTX for sm_100a.
#pragma once
#include <stdint.h>

namespace uavrl {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order generic-proxy smem accesses before subsequent async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one contiguous global -> shared bulk copy (TMA engine); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// bulk copies are limited by the mbarrier tx-count (2^20-1 bytes); our images are < 128 KB, but
// split anyway so several TMA requests are in flight.
__device__ __forceinline__ void bulk_g2s_chunked(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    mbar_arrive_expect_tx(bar, bytes);
    const uint32_t chunk = 16384;
    for (uint32_t off = 0; off < bytes; off += chunk) {
        const uint32_t n = (bytes - off < chunk) ? (bytes - off) : chunk;
        bulk_g2s(static_cast<char *>(smem_dst) + off, static_cast<const char *>(gsrc) + off, n, bar);
    }
}

}  // namespace uavrl
