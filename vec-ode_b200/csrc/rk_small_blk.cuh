// rk_small_blk.cuh — the one-event adaptive sweep on a TILE-BLOCKED copy of the solver's state.
//
// The one-event control kernel of rk_small2.cuh (rk_ctl2w_staged_kernel) reads and writes nine separate arrays per attempt —
// the SoA rows of x, per-trajectory parameters, t, h, prev_h, dx_norm, the status word and two counters — each through its own
// 64-bit base address: a third of its instructions computed addresses, staged rows and issued copies, and instruction issue,
// not HBM, bounds that kernel. Here everything one warp needs for its 64 trajectories is ONE contiguous block:
//
//     tile (64 trajectories) = [ x_0 | .. | x_{D-1} | t | h | p_q (per-trajectory parameters) ]   rows of 64 doubles, 512 B each
//                              [ word | n_accept | n_reject ]                                       rows of 64 u32,    256 B each
//                              ----------------------------------------------------------------  read part: `read_bytes`
//                              [ prev_h | dx_norm ]                                                 rows of 64 doubles (written; prev_h read only at checkpoints)
//
// so a warp's whole input arrives with ONE bulk copy (TMA, cp.async.bulk + a per-warp mbarrier pipeline, no CTA barrier in the
// tile loop) and every load and store of a lane is `tile + constant`. The block holds exactly what the public layout holds
// (SoA ensemble + CtlArrays); solver.cu packs it when a solver enters a run of one-event adaptive launches and unpacks it
// before anything else looks at the state (vo_current, vo_solver_stats, another kernel, the stage path, ...). Arithmetic,
// event logic and the tile -> CTA map (CTA b owns tiles 4b .. 4b+3 of every grid-stride round) are those of
// rk_ctl2w_staged_kernel, so results are bit-identical to it (tests/test_gpu_rk.py) and the launches chain the same way.
//
// MEASURED (round 2, 10^6 Van der Pol oscillators, DoPri5, one attempt per launch; gpurun_out r2c): 10.24 M warp instructions per
// launch against 10.67 M on the public layout (fast; 13.22 M against 13.74 M strict), the same 52 MB of DRAM reads — and 18.4 us
// against 17.1 us per launch (fast), 21.2 against 20.5 (strict). The sweep is not bound by instruction issue after all: it moves
// 92-96 B per attempt (52 read: state 16, t 8, h 8, mu 8, word + two counters 12; ~42 written), which at 17.1 us is 5.5 TB/s =
// 85 % of the measured copy peak. Fewer instructions buy nothing there, and one 3.3 KB bulk copy per warp tile has a longer
// latency than the seven 16-byte cp.async of the public-layout kernel. Kept as an option (vo_solver_set_blocked), off by default.
#pragma once
#include "rk_small2.cuh"

#define VO_BLK_WT 64   // trajectories per warp tile
#define VO_BLK_NST 4   // tiles in flight per warp

struct BlkView {
    unsigned char* base;   // [n_wtiles][stride] bytes
    int64_t n_wtiles;      // warp tiles, a multiple of 4 (padded slots carry VO_TRAJ_DONE)
    uint32_t read_bytes;   // R * 512 + 768
    uint32_t stride;       // read_bytes + 1024
    int R;                 // rows of doubles in the read part: D + 2 + per-trajectory parameters
};

static __host__ __device__ __forceinline__ uint32_t blk_read_bytes(int R) { return (uint32_t)R * 512u + 768u; }

// one trajectory of a tile (slot 0..63) for ctl_lane
struct BlkAcc {
    unsigned char* tile;  // this trajectory's tile in global memory
    int slot, D, R;
    int64_t i, N;
    __device__ __forceinline__ double* row(int r) const { return reinterpret_cast<double*>(tile + 512 * r) + slot; }
    __device__ __forceinline__ uint32_t* wrow(int r) const { return reinterpret_cast<uint32_t*>(tile + 512 * R + 256 * r) + slot; }
    __device__ __forceinline__ double* orow(int r) const { return reinterpret_cast<double*>(tile + 512 * R + 768 + 512 * r) + slot; }
    __device__ __forceinline__ int64_t gidx() const { return i; }
    __device__ __forceinline__ int64_t count() const { return N; }
    __device__ __forceinline__ void st_x(int c, double v) const { *row(c) = v; }
    __device__ __forceinline__ void st_t(double v) const { *row(D) = v; }
    __device__ __forceinline__ void st_h(double v) const { *row(D + 1) = v; }
    __device__ __forceinline__ double ld_prev_h() const { return *orow(0); }
    __device__ __forceinline__ void st_prev_h(double v) const { *orow(0) = v; }
    __device__ __forceinline__ void st_dxn(double v) const { *orow(1) = v; }
    __device__ __forceinline__ void st_word(uint32_t v) const { *wrow(0) = v; }
    __device__ __forceinline__ void st_acc(uint32_t v) const { *wrow(1) = v; }
    __device__ __forceinline__ void st_rej(uint32_t v) const { *wrow(2) = v; }
};

// ---- public layout <-> tiles (one thread per slot; both sides coalesced: 64 consecutive trajectories per row) ----------------
struct BlkPackArgs {
    const double* x;  // SoA [D][N]
    int64_t N;
    CtlArrays ca;
    const double* par[VO_MAX_PARAMS];  // the per-trajectory parameter arrays, compacted (npt of them)
    int D, npt;
};

static __global__ void blk_pack_kernel(BlkView bv, BlkPackArgs a) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // slot number = trajectory number (padded)
    if (j >= bv.n_wtiles * VO_BLK_WT) return;
    const BlkAcc t{bv.base + (j / VO_BLK_WT) * (int64_t)bv.stride, (int)(j % VO_BLK_WT), a.D, bv.R, j, a.N};
    if (j < a.N) {
        for (int c = 0; c < a.D; ++c) t.st_x(c, a.x[c * a.N + j]);
        t.st_t(a.ca.t[j]), t.st_h(a.ca.h[j]);
        for (int q = 0; q < a.npt; ++q) *t.row(a.D + 2 + q) = a.par[q][j];
        t.st_word(a.ca.word[j]), t.st_acc(a.ca.n_accept[j]), t.st_rej(a.ca.n_reject[j]);
        t.st_prev_h(a.ca.prev_h[j]), t.st_dxn(a.ca.dx_norm[j]);
    } else {  // padding: a finished trajectory
        for (int r = 0; r < bv.R; ++r) *t.row(r) = 0.0;
        t.st_word((uint32_t)VO_TRAJ_DONE << VO_WORD_STATUS_SHIFT), t.st_acc(0), t.st_rej(0), t.st_prev_h(0.0), t.st_dxn(0.0);
    }
}

static __global__ void blk_unpack_kernel(BlkView bv, BlkPackArgs a, double* __restrict__ x_out) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= a.N) return;
    const BlkAcc t{bv.base + (j / VO_BLK_WT) * (int64_t)bv.stride, (int)(j % VO_BLK_WT), a.D, bv.R, j, a.N};
    for (int c = 0; c < a.D; ++c) x_out[c * a.N + j] = *t.row(c);
    a.ca.t[j] = *t.row(a.D), a.ca.h[j] = *t.row(a.D + 1);
    a.ca.word[j] = *t.wrow(0), a.ca.n_accept[j] = *t.wrow(1), a.ca.n_reject[j] = *t.wrow(2);
    a.ca.prev_h[j] = *t.orow(0), a.ca.dx_norm[j] = *t.orow(1);
}

// ---- the kernel --------------------------------------------------------------------------------------------------------------
template <class RHS, int S, bool STRICT>
__global__ void __launch_bounds__(128, VO_CTL2_MIN_BLOCKS + 1) rk_ctl2b_kernel(const BlkView bv, int64_t N, const __grid_constant__ TableauDev tb,
                                                                               const __grid_constant__ RhsParams rp, const __grid_constant__ CtlShared cs,
                                                                               EvSlot* __restrict__ ev, const pipe::Chain ch) {
    constexpr int D = RHS::D, NP = RHS::NP, U = 2, NST = VO_BLK_NST;
    extern __shared__ __align__(128) unsigned char sraw[];  // [4 warps][NST][read_bytes]
    __shared__ __align__(8) uint64_t full[4][NST];
    __shared__ double s_tl[VO_INLINE_TLIST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t RB = bv.read_bytes;
    unsigned char* wbase = sraw + (size_t)warp * NST * RB;
    const int64_t n_ctiles = bv.n_wtiles / 4, G = gridDim.x, first = blockIdx.x;
    const int64_t my_count = first < n_ctiles ? (n_ctiles - first + G - 1) / G : 0;
    const TList tl = tlist_stage(cs, s_tl);
    unsigned c_step = 0, c_chkpt = 0, c_rej = 0, c_end = 0, c_stuck = 0;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < NST; ++st) pipe::mbar_init(&full[warp][st], 1);
        pipe::fence_mbar_init();
    }
    pipe::chain_enter(ch);  // __syncthreads() inside: barriers and the t_list copy are visible
    if (lane == 0) asm volatile("fence.proxy.async;" ::: "memory");  // this lane's bulk copies read what the previous launch stored
    auto issue = [&](int64_t k) {  // lane 0: the whole read part of this warp's k-th tile in one bulk copy
        if (lane == 0 && k < my_count) {
            const int st = (int)(k % NST);
            const int64_t wt = (first + k * G) * 4 + warp;
            pipe::mbar_expect_tx(&full[warp][st], RB);
            pipe::bulk_g2s(wbase + (size_t)st * RB, bv.base + wt * (int64_t)bv.stride, RB, &full[warp][st]);
        }
    };
    for (int64_t k = 0; k < NST - 1; ++k) issue(k);
    for (int64_t k = 0; k < my_count; ++k) {
        const int64_t wt = (first + k * G) * 4 + warp;
        unsigned char* gt = bv.base + wt * (int64_t)bv.stride;  // this tile in global memory
        const int64_t i0 = wt * VO_BLK_WT + 2 * lane;           // this lane's two trajectories: i0, i0 + 1
        __syncwarp();  // every lane is done with the buffer the next copy goes to
        issue(k + NST - 1);
        pipe::mbar_wait(&full[warp][k % NST], (uint32_t)((k / NST) & 1));
        const unsigned char* stage = wbase + (size_t)(k % NST) * RB;
        double xc[U][D], p[U][NP], t[U], h[U];
        uint32_t word[U], n_acc[U], n_rej[U];
        int R = 0;
        {
            auto row2 = [&](double& a, double& b) {
                const double2 v = *reinterpret_cast<const double2*>(stage + 512 * (R++) + 16 * lane);
                a = v.x, b = v.y;
            };
#pragma unroll
            for (int c = 0; c < D; ++c) row2(xc[0][c], xc[1][c]);
            row2(t[0], t[1]);
            row2(h[0], h[1]);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (rp.per_traj[q]) row2(p[0][q], p[1][q]);
                else p[0][q] = p[1][q] = rp.shared[q];
            }
            const unsigned char* w = stage + 512 * R + 8 * lane;
            const uint2 a = *reinterpret_cast<const uint2*>(w), b = *reinterpret_cast<const uint2*>(w + 256), c2 = *reinterpret_cast<const uint2*>(w + 512);
            word[0] = a.x, word[1] = a.y, n_acc[0] = b.x, n_acc[1] = b.y, n_rej[0] = c2.x, n_rej[1] = c2.y;
        }
        unsigned char* gw = gt + 512 * R + 8 * lane;         // this lane's slots of the word rows
        unsigned char* go = gt + 512 * R + 768 + 16 * lane;  // ... and of prev_h / dx_norm
        bool live[U];
        bool pair = cs.k_events == 1;
#pragma unroll
        for (int u = 0; u < U; ++u) live[u] = !((word[u] >> VO_WORD_STATUS_SHIFT) & VO_TRAJ_DONE), pair = pair && live[u];
        double dt[U] = {0.0, 0.0}, t_tgt[U] = {0.0, 0.0};
        if (pair) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int tgt = (int)(word[u] & VO_WORD_TGT_MASK);
                if (tgt >= cs.n_tlist) {
                    pair = false;
                } else {
                    t_tgt[u] = tl.at(tgt);
                    const double rem = t_tgt[u] - t[u];  // step_size_of (ode.rs:165-176) + check_step (ode.rs:389-399)
                    if (fabs(rem) <= 2.220446049250313e-16) pair = false;
                    dt[u] = rem < h[u] ? rem : h[u];
                }
            }
        }
        if (pair) {
            double xf[U][D], xe[U][D];
            rk_attempt_n<RHS, S, STRICT, U>(tb, t, dt, xc, p, xf, xe);
            double dxn[U] = {0.0, 0.0}, new_h[U], acc[U];
            bool rej[U], nonfin[U];
#pragma unroll
            for (int u = 0; u < U; ++u) acc[u] = err_sumsq<STRICT, D>(xe[u]);
            controller_l2_n<STRICT, U>(acc, h, cs, cs.record_dx_norm != 0, dxn, new_h, rej, nonfin);  // handle_step_adaptive, ode.rs:311-334
            const double t_new[U] = {t[0] + dt[0], t[1] + dt[1]};                                      // advance, ode.rs:184-188
            // apply_step (ode.rs:402-428) + masked write-back: 128-bit stores where both trajectories commit the same way
            if (!rej[0] && !rej[1]) {
#pragma unroll
                for (int c = 0; c < D; ++c) *reinterpret_cast<double2*>(gt + 512 * c + 16 * lane) = make_double2(xf[0][c], xf[1][c]);
                *reinterpret_cast<double2*>(gt + 512 * D + 16 * lane) = make_double2(t_new[0], t_new[1]);
                *reinterpret_cast<uint2*>(gw + 256) = make_uint2(n_acc[0] + 1, n_acc[1] + 1);
                c_step += 2;
            } else if (rej[0] && rej[1]) {
                *reinterpret_cast<uint2*>(gw + 512) = make_uint2(n_rej[0] + 1, n_rej[1] + 1);
                c_rej += 2;
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (rej[u]) {
                        *reinterpret_cast<uint32_t*>(gw + 512 + 4 * u) = n_rej[u] + 1, ++c_rej;
                    } else {
#pragma unroll
                        for (int c = 0; c < D; ++c) *reinterpret_cast<double*>(gt + 512 * c + 16 * lane + 8 * u) = xf[u][c];
                        *reinterpret_cast<double*>(gt + 512 * D + 16 * lane + 8 * u) = t_new[u];
                        *reinterpret_cast<uint32_t*>(gw + 256 + 4 * u) = n_acc[u] + 1, ++c_step;
                    }
                }
            }
            *reinterpret_cast<double2*>(gt + 512 * (D + 1) + 16 * lane) = make_double2(new_h[0], new_h[1]);  // update_step_size, ode.rs:202-205
            if (!cs.lazy_prev_h) *reinterpret_cast<double2*>(go) = make_double2(h[0], h[1]);
            if (cs.record_dx_norm) *reinterpret_cast<double2*>(go + 512) = make_double2(dxn[0], dxn[1]);
            // the rare rest: status bits, and prev_h where a checkpoint comes next (its one reader, ode.rs:192-195)
            bool special = nonfin[0] || nonfin[1];
#pragma unroll
            for (int u = 0; u < U; ++u) special = special || (rej[u] && h[u] <= cs.min_dt) || (!rej[u] && fabs(t_tgt[u] - t_new[u]) <= 2.220446049250313e-16);
            if (special) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t status = word[u] >> VO_WORD_STATUS_SHIFT;
                    if (nonfin[u]) status |= VO_TRAJ_NONFINITE;
                    if (rej[u] && h[u] <= cs.min_dt) status |= VO_TRAJ_STUCK, ++c_stuck;
                    if (cs.lazy_prev_h && !rej[u] && fabs(t_tgt[u] - t_new[u]) <= 2.220446049250313e-16) *reinterpret_cast<double*>(go + 8 * u) = h[u];
                    const uint32_t nw = (word[u] & VO_WORD_TGT_MASK) | (status << VO_WORD_STATUS_SHIFT);
                    if (nw != word[u]) *reinterpret_cast<uint32_t*>(gw + 4 * u) = nw;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (live[u])
                    ctl_lane<RHS, S, STRICT, 1>(BlkAcc{gt, 2 * lane + u, D, R, i0 + u, N}, tb, cs, tl, word[u], xc[u], p[u], t[u], h[u], n_acc[u], n_rej[u], c_step,
                                                c_chkpt, c_rej, c_end, c_stuck);
        }
    }
    ctl_count_events(cs, ev, c_step, c_chkpt, c_rej, c_end, c_stuck);
    pipe::chain_exit(ch);
}
