// tile_pipe.cuh — shared-memory staging of SoA tiles with 1-D bulk async copies (TMA, cp.async.bulk) completing on
// mbarriers, plus programmatic-dependent-launch controls. An elected thread keeps NSTAGE tiles of every row a kernel
// reads in flight per CTA, so the bytes in flight per SM are set by shared memory (tens of KB), not by how many loads
// the register file can hold — which is what an HBM-latency-bound stream of small per-thread loads runs out of.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

namespace pipe {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy; `bytes` and both addresses must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes),
                 "r"(s32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = s32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!ok);
}

// Programmatic dependent launch: `launch_dependents` lets the next kernel of the stream start occupying SMs as this
// grid's CTAs retire; `grid_wait` blocks until the previous kernel of the stream has completed and flushed. Every kernel
// launched with the attribute calls grid_wait() before its first global-memory access.
__device__ __forceinline__ void launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- CTA-to-CTA chaining across launches --------------------------------------------------------------------------
// Consecutive launches of one solver use the same grid and the same tile -> CTA mapping, and trajectories are
// independent, so CTA b of launch n+1 depends only on CTA b of launch n. Instead of waiting for the whole previous grid
// (griddepcontrol.wait), a chained launch waits for its own predecessor's generation flag: the ramp of launch n+1
// overlaps the tail of launch n. Safe with programmatic dependent launch because a dependent grid is only scheduled
// after every CTA of the primary has started, i.e. the CTA being waited for is already resident and waits for nothing.
// The spin is bounded: a time-out falls back to griddepcontrol.wait (correct, just slower), so a broken chain cannot hang.
struct Chain {
    uint32_t* flags;  // one generation counter per CTA
    uint32_t gen;     // generation of this launch (> 0)
    int chained;      // 1: wait for flags[b] >= gen-1; 0: wait for the previous grid
};

__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Entry protocol; every thread calls it (contains a __syncthreads()).
__device__ __forceinline__ void chain_enter(const Chain& ch) {
    launch_dependents();
    if (!ch.chained) {
        grid_wait();
    } else {
        if (threadIdx.x == 0) {
            bool ok = false;
            for (int spin = 0; spin < (1 << 16); ++spin) {
                if ((int32_t)(ld_acquire(ch.flags + blockIdx.x) - (ch.gen - 1)) >= 0) {
                    ok = true;
                    break;
                }
                __nanosleep(64);
            }
            if (!ok) grid_wait();
            asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies that follow read what the predecessor stored
        }
    }
    __syncthreads();
}

// Exit protocol; every thread calls it after its last global store.
__device__ __forceinline__ void chain_exit(const Chain& ch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        st_release(ch.flags + blockIdx.x, ch.gen);
    }
}

}  // namespace pipe
