// pq_tma.cuh -- mbarrier and 1-D bulk-copy (TMA, cp.async.bulk) primitives shared by the ring kernels.
// SASS: UBLKCP.S.G (global -> shared), UBLKCP.G.S (shared -> global), SYNCS.* (mbarrier).
#pragma once

#include <cstdint>

namespace pq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {   // try_wait suspends the thread in hardware for a bounded time; loop until the phase flips
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (no tensor map needed);
// addresses and size must be multiples of 16 bytes
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// all but the N most recent bulk groups are complete
template <int N>
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// order generic-proxy shared-memory writes before async-proxy (TMA) reads of the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barrier 1 over the consumer warps only (the producer warp never joins)
template <int THREADS>
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

// ---- dynamic tile assignment for the persistent ring kernels --------------------------------------------
// A static tile -> CTA map makes the whole grid wait for its slowest SM (the two dies of a B200 differ by
// ~10 % in memory latency): 6.7 vs 7.2 TB/s on a 4:1 read:write stream (profiles/r1_sched_probe_*.txt).  The
// producer thread of each CTA therefore draws tile indices from a global counter.  sched[0] = next tile,
// sched[1] = CTAs that have finished drawing; the last one resets both, so the pair is zero between launches.
// Tiles are drawn kSchedChunk at a time (one atomic per 8 tiles keeps the single counter far from its
// same-address throughput limit) and the NEXT chunk is requested while the current one is being issued.
constexpr int kSchedChunk = 8;
__device__ __forceinline__ long long sched_next_chunk(unsigned long long* sched) {
    return static_cast<long long>(atomicAdd(sched, static_cast<unsigned long long>(kSchedChunk)));
}
__device__ __forceinline__ void sched_cta_done(unsigned long long* sched) {
    __threadfence();
    if (atomicAdd(sched + 1, 1ull) == gridDim.x - 1) {
        sched[0] = 0ull;
        sched[1] = 0ull;
        __threadfence();
    }
}

}  // namespace pq
