// rk_small_launch.cuh — host-side dispatch of rk_small_kernel over (stage count, arithmetic mode, control mode).
// Each RHS family is instantiated in its own translation unit (rk_small_<family>.cu) so they compile in parallel.
#pragma once
#include <unordered_map>

#include "rk_small.cuh"

constexpr int RK_SMALL_THREADS = 128;

struct SmallLaunch {
    vo_ctx ctx;
    double* x;
    int64_t N;
    const TableauDev* tb;
    const RhsParams* rp;
    CtlArrays ca;
    const CtlShared* cs;   // per-trajectory control (sl == nullptr)
    const StepList* sl;    // lock-step fixed steps (cs == nullptr)
    EvSlot* ev;
};

// Persistent grid: every thread gets the same number of trajectories (no partial last wave), all CTAs resident.
template <class K> static unsigned persistent_grid(vo_ctx c, K kernel, int64_t N) {
    static std::unordered_map<const void*, int> cache;  // resident CTAs per SM of each kernel instantiation
    int& bps = cache[(const void*)kernel];
    if (bps == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, RK_SMALL_THREADS, 0) != cudaSuccess || bps < 1) bps = 1;
    }
    const int64_t slots = (int64_t)c->sm_count * bps * RK_SMALL_THREADS;
    const int64_t iters = ceil_div(N, slots);
    return (unsigned)ceil_div(N, iters * RK_SMALL_THREADS);
}

template <class RHS, int S, bool STRICT> static void launch_one(const SmallLaunch& L) {
    if (L.sl) {
        auto k = rk_fixed_kernel<RHS, S, STRICT>;
        k<<<persistent_grid(L.ctx, k, L.N), RK_SMALL_THREADS, 0, L.ctx->stream>>>(L.x, L.N, *L.tb, *L.rp, *L.sl);
    } else {
        auto k = rk_ctl_kernel<RHS, S, STRICT>;
        k<<<persistent_grid(L.ctx, k, L.N), RK_SMALL_THREADS, 0, L.ctx->stream>>>(L.x, L.N, *L.tb, *L.rp, L.ca, *L.cs, L.ev);
    }
}

template <class RHS, int S> static void launch_s(const SmallLaunch& L) {
    if (L.ctx->arith == VO_ARITH_STRICT) launch_one<RHS, S, true>(L);
    else launch_one<RHS, S, false>(L);
}

// Stage counts with a fully unrolled, register-resident instantiation; anything else (<= VO_MAX_STAGES)
// runs the generic S = 0 instantiation.
template <class RHS> static int32_t launch_family(const SmallLaunch& L) {
    switch (L.tb->s) {
        case 4: launch_s<RHS, 4>(L); break;
        case 6: launch_s<RHS, 6>(L); break;
        case 7: launch_s<RHS, 7>(L); break;
        default: launch_s<RHS, 0>(L); break;
    }
    return VO_OK;
}

// one per family TU
int32_t launch_small_diag(const SmallLaunch& L, int d);
int32_t launch_small_harmonic(const SmallLaunch& L);
int32_t launch_small_lorenz(const SmallLaunch& L);
int32_t launch_small_vdp(const SmallLaunch& L);
