// rk_small_launch.cuh — host-side dispatch of rk_small_kernel over (stage count, arithmetic mode, control mode).
// Each RHS family is instantiated in its own translation unit (rk_small_<family>.cu) so they compile in parallel.
#pragma once
#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "rk_small_blk.cuh"

constexpr int RK_SMALL_THREADS = 128;

// Host side of pipe::Chain for one solver: which staged kernel wrote the generation flags last, and with which grid. A launch
// may skip the grid-wide wait only if the launch before it (of this solver) was the SAME kernel instantiation on the SAME
// grid — then CTA b of both launches owns the same tiles (the tile -> CTA map is a function of the grid size) — and nothing
// else has touched the solver's state in between (`live`, kept by solver.cu from the ctx's epoch).
struct ChainState {
    uint32_t* flags = nullptr;
    uint32_t gen = 0;
    const void* fn = nullptr;
    unsigned grid = 0;
    bool live = false;
    pipe::Chain next(const void* kernel, unsigned g, bool staged) {
        pipe::Chain ch{flags, ++gen, (staged && live && fn == kernel && grid == g) ? 1 : 0};
        fn = kernel, grid = g, live = staged;
        return ch;
    }
};

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
    ChainState* cst;       // chaining state of the owning solver
    const BlkView* blk = nullptr;  // non-NULL: the state lives in the tile-blocked copy (rk_small_blk.cuh) and x / ca are stale
};

// The staged kernels need 16-byte aligned SoA rows (N even) and at least one full tile.
static inline bool small_path_is_staged(int64_t N) { return (N % 2 == 0) && N >= RK_SMALL_THREADS; }

// Persistent grid: every CTA gets the same number of 128-trajectory tiles (no partial last wave), all CTAs resident.
// The shared-memory size depends on run-time facts (how many RHS parameters are per-trajectory arrays), and both the large
// shared-memory opt-in and the occupancy are per device, so resident CTAs per SM are cached per (device, kernel, smem).
template <class K> static unsigned persistent_grid(vo_ctx c, K kernel, int64_t N, size_t smem, int tile = RK_SMALL_THREADS) {
    static std::map<std::tuple<int, const void*, size_t>, int> cache;
    static std::mutex mu;  // contexts may be driven from different host threads
    int bps;
    {
        std::lock_guard<std::mutex> lock(mu);
        int& slot = cache[std::make_tuple(c->device, (const void*)kernel, smem)];
        if (slot == 0) {
            vo_ensure_smem_attr(c->device, (const void*)kernel, smem);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&slot, kernel, RK_SMALL_THREADS, smem) != cudaSuccess || slot < 1) slot = 1;
        }
        bps = slot;
    }
    const int64_t tiles = ceil_div(N, tile);
    const int64_t iters = ceil_div(tiles, (int64_t)c->sm_count * bps);
    return (unsigned)ceil_div(tiles, iters);
}

// A CHAINED launch (see pipe::Chain) is submitted with programmatic stream serialisation, so its CTAs are scheduled as
// the previous grid's CTAs retire and each starts as soon as its own predecessor CTA has published its generation flag.
// An unchained launch is an ordinary launch: it starts after everything before it in the stream has completed.
// The kernel's last parameter is the pipe::Chain, which is made here from the solver's ChainState once the grid is known.
template <class... KArgs, class... Args>
static void launch_staged(ChainState* cst, void (*kernel)(KArgs...), unsigned grid, size_t smem, cudaStream_t stream, Args&&... args) {
    const pipe::Chain ch = cst->next((const void*)kernel, grid, true);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(RK_SMALL_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = ch.chained ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)..., ch);
}

// When the one-event adaptive sweep may run on the tile-blocked copy of the state: the dispatch condition of the
// two-trajectory control kernels (below) for a compiled-in RHS family with an unrolled stage count.
static inline bool small_path_blocked_ok(int64_t N, int stages, const CtlShared& cs) {
    const bool unrolled = stages == 4 || stages == 6 || stages == 7;
    return unrolled && small_path_is_staged(N) && N >= 4 * VO_TILE_CTL && cs.adaptive && cs.use_err && cs.norm_kind == VO_NORM_L2 && cs.k_events == 1;
}

static inline int per_traj_rows(const RhsParams& rp, int np) {
    int n = 0;
    for (int q = 0; q < np; ++q) n += rp.per_traj[q] ? 1 : 0;
    return n;
}

template <class RHS, int S, bool STRICT> static void launch_one(const SmallLaunch& L) {
    const bool staged = small_path_is_staged(L.N);
    if (L.sl) {
        if (staged) {
            if constexpr (S > 0) {
                if (L.N >= 4 * VO_TILE2) {  // two trajectories per thread (rk_small2.cuh)
                    const size_t smem2 = (size_t)VO_STAGES * (RHS::D + per_traj_rows(*L.rp, RHS::NP)) * VO_TILE2 * sizeof(double);
#ifdef VO_FIXED2_CTA  // A/B switch (tools/build_variant.sh): the CTA-staged version of the kernel
                    auto k2 = rk_fixed2_staged_kernel<RHS, S, STRICT>;
#else
                    auto k2 = rk_fixed2w_staged_kernel<RHS, S, STRICT>;
#endif
                    launch_staged(L.cst, k2, persistent_grid(L.ctx, k2, L.N, smem2, VO_TILE2), smem2, L.ctx->stream, L.x, L.N, *L.tb, *L.rp, *L.sl);
                    return;
                }
            }
            const size_t smem = (size_t)VO_STAGES * (RHS::D + per_traj_rows(*L.rp, RHS::NP)) * VO_TILE * sizeof(double);
            auto k = rk_fixed_staged_kernel<RHS, S, STRICT>;
            launch_staged(L.cst, k, persistent_grid(L.ctx, k, L.N, smem), smem, L.ctx->stream, L.x, L.N, *L.tb, *L.rp, *L.sl);
        } else {
            auto k = rk_fixed_kernel<RHS, S, STRICT>;
            L.cst->live = false;
            k<<<persistent_grid(L.ctx, k, L.N, 0), RK_SMALL_THREADS, 0, L.ctx->stream>>>(L.x, L.N, *L.tb, *L.rp, *L.sl);
        }
    } else {
        if (staged) {
            const size_t smem = (size_t)VO_STAGES * ((RHS::D + 2 + per_traj_rows(*L.rp, RHS::NP)) * VO_TILE * sizeof(double) + 3 * VO_TILE * sizeof(uint32_t));
            const bool common = L.cs->adaptive && L.cs->use_err && L.cs->norm_kind == VO_NORM_L2;
            if constexpr (S > 0) {
                if (L.blk) {  // the one-event sweep on the tile-blocked state (solver.cu decided with small_path_blocked_ok)
                    const size_t smemb = (size_t)4 * VO_BLK_NST * L.blk->read_bytes;
                    auto kb = rk_ctl2b_kernel<RHS, S, STRICT>;
                    launch_staged(L.cst, kb, persistent_grid(L.ctx, kb, L.blk->n_wtiles * VO_BLK_WT, smemb, VO_TILE_CTL), smemb, L.ctx->stream, *L.blk, L.N, *L.tb, *L.rp,
                                  *L.cs, L.ev);
                    return;
                }
                if (common && L.cs->k_events == 1 && L.N >= 4 * VO_TILE_CTL) {  // several trajectories per thread (rk_small2.cuh)
                    const size_t smem2 = (size_t)VO_STAGES * ((RHS::D + 2 + per_traj_rows(*L.rp, RHS::NP)) * VO_TILE_CTL * sizeof(double) + 3 * VO_TILE_CTL * sizeof(uint32_t));
#ifdef VO_CTL2_CTA  // A/B switch (tools/build_variant.sh): the CTA-staged version of the kernel
                    auto k2 = rk_ctl2_staged_kernel<RHS, S, STRICT>;
#else
                    auto k2 = rk_ctl2w_staged_kernel<RHS, S, STRICT>;  // warp-autonomous staging: 15.9 against 16.9 us on the VdP sweep
#endif
                    launch_staged(L.cst, k2, persistent_grid(L.ctx, k2, L.N, smem2, VO_TILE_CTL), smem2, L.ctx->stream, L.x, L.N, *L.tb, *L.rp, L.ca,
                                  *L.cs, L.ev);
                    return;
                }
            }
            auto k = common ? rk_ctl_staged_kernel<RHS, S, STRICT, 1> : rk_ctl_staged_kernel<RHS, S, STRICT, 0>;
            launch_staged(L.cst, k, persistent_grid(L.ctx, k, L.N, smem), smem, L.ctx->stream, L.x, L.N, *L.tb, *L.rp, L.ca, *L.cs, L.ev);
        } else {
            auto k = rk_ctl_kernel<RHS, S, STRICT>;
            L.cst->live = false;
            k<<<persistent_grid(L.ctx, k, L.N, 0), RK_SMALL_THREADS, 0, L.ctx->stream>>>(L.x, L.N, *L.tb, *L.rp, L.ca, *L.cs, L.ev);
        }
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
// user RHS compiled at run time (nvrtc_rhs.cu)
int32_t launch_small_custom(const SmallLaunch& L, vo_rhs_s* r);
