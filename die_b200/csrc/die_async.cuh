// die_async.cuh -- the five asynchronous-copy primitives the bulk field kernel needs (sm_90+ PTX): an mbarrier in
// shared memory that counts transaction bytes, and 1-D bulk copies global -> shared (the TMA unit, no tensor map:
// SASS UBLKCP) that complete on it.
//
//   mbar_init(bar, n)            one thread; n = arrivals per phase
//   mbar_fence_init()            makes the initialised barrier visible to the async proxy
//   mbar_arrive_expect_tx(bar,b) one arrival + "b more bytes will land in this phase"
//   bulk_g2s(dst, src, b, bar)   copy b bytes (multiple of 16, both addresses 16-byte aligned); completes b bytes on bar
//   mbar_wait(bar, parity)       spin until the phase with this parity has completed; bounded: a barrier that never
//                                completes traps instead of hanging the GPU
//   proxy_fence_async()          orders this thread's earlier generic-proxy accesses to shared memory before later
//                                async-proxy ones (a buffer the threads wrote is about to be overwritten by a bulk copy)
//
// Under the CPU emulator (tests/hostsim, DIE_HOSTSIM) the same names are provided by tests/hostsim/cuda_runtime.h:
// copies are deferred until somebody waits and their destination is poisoned meanwhile, byte counts and phases are
// checked, so a protocol error shows up as wrong results or a reported deadlock.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace die {

#if !defined(DIE_HOSTSIM)

typedef unsigned long long mbar_t;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(mbar_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(mbar_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, mbar_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void mbar_wait(mbar_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((spins & 63u) == 63u) {             // two seconds of waiting for a few hundred KB: the protocol is broken
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}

__device__ __forceinline__ void proxy_fence_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

#endif  // !DIE_HOSTSIM

}  // namespace die
