// rk_small_launch.cuh — host-side dispatch of rk_small_kernel over (stage count, arithmetic mode, control mode).
// Each RHS family is instantiated in its own translation unit (rk_small_<family>.cu) so they compile in parallel.
#pragma once
#include "rk_small.cuh"

constexpr int RK_SMALL_THREADS = 128;

struct SmallLaunch {
    vo_ctx ctx;
    double* x;
    int64_t N;
    const TableauDev* tb;
    const RhsParams* rp;
    CtlArrays ca;
    const CtlShared* cs;
    EvSlot* ev;
    bool uniform;
};

template <class RHS, int S, bool STRICT, bool UNIFORM> static void launch_one(const SmallLaunch& L) {
    const unsigned grid = (unsigned)ceil_div(L.N, RK_SMALL_THREADS);
    rk_small_kernel<RHS, S, STRICT, UNIFORM><<<grid, RK_SMALL_THREADS, 0, L.ctx->stream>>>(L.x, L.N, *L.tb, *L.rp, L.ca, *L.cs, L.ev);
}

template <class RHS, int S> static void launch_s(const SmallLaunch& L) {
    const bool strict = L.ctx->arith == VO_ARITH_STRICT;
    if (strict && L.uniform) launch_one<RHS, S, true, true>(L);
    else if (strict) launch_one<RHS, S, true, false>(L);
    else if (L.uniform) launch_one<RHS, S, false, true>(L);
    else launch_one<RHS, S, false, false>(L);
}

// Stage counts with a fully unrolled, register-resident instantiation; anything else (<= VO_SMALL_MAX_STAGES)
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
