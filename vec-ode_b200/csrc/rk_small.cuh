// rk_small.cuh — whole-attempt Runge-Kutta kernel for ensembles of small systems (d <= 8).
//
// One thread owns one trajectory. A launch executes up to `k_events` calls of the reference's
// `step()` / `step_adaptive()` (src/base/ode.rs:249-253, 337-341) for that trajectory with the state, all
// s stage derivatives and the controller scalars in REGISTERS: state is read from HBM once and written
// once per launch (SoA, coalesced: lane i touches component c at x[c*N + i]). The tableau is a by-value
// kernel parameter, so after full unrolling every coefficient is a constant-bank operand.
//
// Everything the reference does between two calls is reproduced per trajectory:
//   step_size_of/check_step (ode.rs:165-176, 389-399)  ->  Step(dt) | Chkpt | End
//   rk_step (rk.rs:90-155)                             ->  X_b, X_berr, x_err in the reference's op order
//   handle_step_adaptive (ode.rs:311-334)              ->  norm, f = rtol/|err|, h update, Reject iff f <= 1
//   apply_step (ode.rs:402-428)                        ->  advance / checkpoint_update
// Rejected or finished lanes are masked: they neither advance nor write state.
#pragma once
#include "common.cuh"
#include "rhs.cuh"

#define VO_WORD_TGT_MASK 0xffffu
#define VO_WORD_STATUS_SHIFT 16
#define VO_EV_SLOTS 32
#define VO_INLINE_TLIST 8

// Per-trajectory controller state (device arrays of length N).
struct CtlArrays {
    double* t;
    double* h;
    double* prev_h;
    double* dx_norm;
    uint32_t* n_accept;
    uint32_t* n_reject;
    uint32_t* word;  // bits 0..15: tgt_t (ode.rs:90); bits 16..: VO_TRAJ_* status
};

// Controller constants shared by the ensemble (ODEAdaptiveData, ode.rs:98-110) + launch options.
struct CtlShared {
    double rtol, alpha, pw, min_dt, max_dt;
    double t_list_inline[VO_INLINE_TLIST];
    const double* t_list;  // device copy when n_tlist > VO_INLINE_TLIST
    int n_tlist;
    int norm_kind;
    int adaptive;      // step_adaptive() vs step()
    int use_err;       // tableau has b_err AND the solver still holds x_err (rk.rs:136-151)
    int k_events;      // calls fused into this launch
    int count_events;  // accumulate vo_step_result counters
    int pw_is_third;   // pw == 1.0/3.0 exactly (the order RK45Solver hard-wires, rk.rs:258-260)
    int record_dx_norm;  // keep ODEAdaptiveData.dx_norm (ode.rs:104) per trajectory
};

// Event counters, VO_EV_SLOTS copies 128 bytes apart to spread the atomics (host sums the slots).
struct EvSlot {
    unsigned long long n_step, n_chkpt, n_reject, n_end, n_stuck;
    unsigned long long pad[11];
};

template <bool STRICT, int D> __device__ __forceinline__ double err_norm(const double (&e)[D], int kind) {
    using A = Ar<STRICT>;
    double acc = 0.0;
    if (kind == VO_NORM_L2) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc = A::add(acc, A::mul(e[c], e[c]));
        return sqrt(acc);
    }
    if (kind == VO_NORM_LINF) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc = fmax(acc, fabs(e[c]));
        return acc;
    }
    if (kind == VO_NORM_L1) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc = A::add(acc, fabs(e[c]));
        return acc;
    }
    // VO_NORM_HYPOT
    if (D == 2) return hypot(e[0], e[D - 1]);
#pragma unroll
    for (int c = 0; c + 1 < D; c += 2) {
        const double m = hypot(e[c], e[c + 1]);
        acc = A::add(acc, A::mul(m, m));
    }
    return sqrt(acc);
}

// v[c] = (sum_{j<n} k[j]*K[j][c]) * dt + x0[c]  in the reference's order (lc.rs:20-54, rk.rs:123-124).
template <bool STRICT, int D, int SM>
__device__ __forceinline__ void combine(const double* __restrict__ k, int n, const double (&K)[SM][D], double dt, const double (&x0)[D],
                                        double (&v)[D]) {
    using A = Ar<STRICT>;
    if (STRICT) {
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = A::mul(k[0], K[0][c]);
#pragma unroll
        for (int j = 1; j < SM; ++j)
            if (j < n) {
#pragma unroll
                for (int c = 0; c < D; ++c) v[c] = A::axpy(v[c], k[j], K[j][c]);
            }
    } else {
        // FMA chain. Zero coefficients are NOT tested for here: a data-dependent skip costs a DSETP + selects per term,
        // more than the FMA it would save (the stage-path kernels, which save a whole HBM pass per zero, do skip).
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = k[0] * K[0][c];
#pragma unroll
        for (int j = 1; j < SM; ++j)
            if (j < n) {
#pragma unroll
                for (int c = 0; c < D; ++c) v[c] = fma(k[j], K[j][c], v[c]);
            }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) v[c] = A::add(A::mul(v[c], dt), x0[c]);
}

// rk_step (src/base/rk.rs:90-155). On return xf is the state the reference propagates (X_berr when the error
// branch runs, else X_b) and xe = X_b - X_berr.
template <class RHS, int S, bool STRICT>
__device__ __forceinline__ void rk_attempt(const TableauDev& tb, bool use_err, double t, double dt, const double (&x0)[RHS::D],
                                           const double (&p)[RHS::NP], double (&xf)[RHS::D], double (&xe)[RHS::D]) {
    using A = Ar<STRICT>;
    constexpr int D = RHS::D;
    constexpr int SM = S > 0 ? S : VO_MAX_STAGES;
    const int s = S > 0 ? S : tb.s;
    double K[SM][D];
    RHS::template eval<STRICT>(t, x0, K[0], p);  // rk.rs:111
#pragma unroll
    for (int i = 1; i < SM; ++i) {
        if (i < s) {
            const double* row = &tb.ac[i * s];
            const double ti = A::add(t, A::mul(row[i], dt));  // rk.rs:119
            double xs[D];
            combine<STRICT, D, SM>(row, i, K, dt, x0, xs);    // rk.rs:121-124
            RHS::template eval<STRICT>(ti, xs, K[i], p);      // rk.rs:127
        }
    }
    combine<STRICT, D, SM>(tb.b, s, K, dt, x0, xf);           // rk.rs:131-133
    if (use_err) {                                            // rk.rs:136-151
#pragma unroll
        for (int c = 0; c < D; ++c) xe[c] = xf[c];            // swap: xe := X_b
        combine<STRICT, D, SM>(tb.b_err, s, K, dt, x0, xf);   // xf := X_berr
#pragma unroll
        for (int c = 0; c < D; ++c) xe[c] = A::sub(xe[c], xf[c]);
    }
}

// ---- lock-step fixed-step kernel -----------------------------------------------------------------------------
// Every trajectory takes the same (t, dt) sequence, which the host has already derived with the reference's state
// machine (ode.rs:165-176, 389-399), so the kernel carries no control arithmetic at all: per trajectory it moves
// 2*d*8 bytes and executes the rk_step FLOPs, nothing else. Persistent grid: each thread walks trajectories
// i, i+stride, ... and loads the NEXT trajectory's state before integrating the current one, so the HBM latency of
// iteration n+1 hides behind the FP64 work of iteration n.
#define VO_MAX_FUSED 32
struct StepList {
    int n;        // steps fused into this launch (<= VO_MAX_FUSED)
    int use_err;  // propagate X_berr (rk.rs:142-146) instead of X_b
    double t[VO_MAX_FUSED];
    double dt[VO_MAX_FUSED];
};

template <class RHS> __device__ __forceinline__ void lane_load(const double* __restrict__ x, int64_t N, const RhsParams& rp, int64_t i,
                                                               double (&xc)[RHS::D], double (&p)[RHS::NP]) {
#pragma unroll
    for (int c = 0; c < RHS::D; ++c) xc[c] = x[c * N + i];
    load_params<RHS::NP>(rp, i, p);
}

template <class RHS, int S, bool STRICT>
__global__ void __launch_bounds__(128) rk_fixed_kernel(double* __restrict__ x, int64_t N, const __grid_constant__ TableauDev tb,
                                                       const __grid_constant__ RhsParams rp, const __grid_constant__ StepList sl) {
    constexpr int D = RHS::D;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double xc[D], p[RHS::NP], xn[D], pn[RHS::NP];
    if (i < N) lane_load<RHS>(x, N, rp, i, xc, p);
    while (i < N) {
        const int64_t j = i + stride;
        if (j < N) lane_load<RHS>(x, N, rp, j, xn, pn);  // prefetch the next trajectory of this thread
        for (int e = 0; e < sl.n; ++e) {
            double xf[D], xe[D];
            rk_attempt<RHS, S, STRICT>(tb, sl.use_err != 0, sl.t[e], sl.dt[e], xc, p, xf, xe);
#pragma unroll
            for (int c = 0; c < D; ++c) xc[c] = xf[c];
        }
#pragma unroll
        for (int c = 0; c < D; ++c) x[c * N + i] = xc[c];
#pragma unroll
        for (int c = 0; c < D; ++c) xc[c] = xn[c];
#pragma unroll
        for (int q = 0; q < RHS::NP; ++q) p[q] = pn[q];
        i = j;
    }
}

// ---- per-trajectory control kernel ---------------------------------------------------------------------------
// alpha * f^pw (ode.rs:133-135). STRICT keeps the general pow of the reference's `powf`; FAST takes the cube root
// directly when pw is exactly 1/3 (the order RK45Solver hard-wires, rk.rs:258-260) — a few ulp apart, 4x fewer FP64 ops.
template <bool STRICT> __device__ __forceinline__ double step_size_mul(double alpha, double f, double pw, int pw_is_third) {
    if (!STRICT && pw_is_third) return alpha * cbrt(f);
    return alpha * pow(f, pw);
}

template <class RHS, int S, bool STRICT>
__global__ void __launch_bounds__(128) rk_ctl_kernel(double* __restrict__ x, int64_t N, const __grid_constant__ TableauDev tb,
                                                     const __grid_constant__ RhsParams rp, const CtlArrays ca, const __grid_constant__ CtlShared cs,
                                                     EvSlot* __restrict__ ev) {
    constexpr int D = RHS::D;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned c_step = 0, c_chkpt = 0, c_rej = 0, c_end = 0, c_stuck = 0;
    const double* tl = cs.n_tlist > VO_INLINE_TLIST ? cs.t_list : cs.t_list_inline;
    // current / prefetched lane
    uint32_t word = 0, word_n = 0;
    double xc[D], p[RHS::NP], t = 0.0, h = 0.0, xn[D], pn[RHS::NP], t_n = 0.0, h_n = 0.0;
    bool live = false, live_n = false;
    if (i < N) {
        word = ca.word[i];
        live = !((word >> VO_WORD_STATUS_SHIFT) & VO_TRAJ_DONE);
        if (live) lane_load<RHS>(x, N, rp, i, xc, p), t = ca.t[i], h = ca.h[i];
    }
    while (i < N) {
        const int64_t j = i + stride;
        if (j < N) {
            word_n = ca.word[j];
            live_n = !((word_n >> VO_WORD_STATUS_SHIFT) & VO_TRAJ_DONE);
            if (live_n) lane_load<RHS>(x, N, rp, j, xn, pn), t_n = ca.t[j], h_n = ca.h[j];
        }
        if (live) {
            double prev_h = 0.0;
            bool prev_h_loaded = false, ctl_dirty = false, moved = false;
            int tgt = (int)(word & VO_WORD_TGT_MASK);
            uint32_t status = word >> VO_WORD_STATUS_SHIFT;
            double dxn = 0.0;
            bool dxn_set = false;
            unsigned l_step = 0, l_rej = 0;
            for (int e = 0; e < cs.k_events; ++e) {
                // ---- step_size_of (ode.rs:165-176) + check_step (ode.rs:389-399)
                int evk;
                double dt = 0.0;
                if (tgt >= cs.n_tlist) {
                    evk = VO_EV_END;
                } else {
                    const double rem = tl[tgt] - t;
                    if (fabs(rem) <= 2.220446049250313e-16) evk = (tgt >= cs.n_tlist - 1) ? VO_EV_END : VO_EV_CHKPT;
                    else dt = rem < h ? rem : h, evk = VO_EV_STEP;
                }
                if (evk == VO_EV_STEP) {
                    double xf[D], xe[D];
                    rk_attempt<RHS, S, STRICT>(tb, cs.use_err != 0, t, dt, xc, p, xf, xe);
                    if (cs.adaptive) {  // handle_step_adaptive, ode.rs:311-334
                        dxn = err_norm<STRICT, D>(xe, cs.norm_kind), dxn_set = true;
                        const double f = cs.rtol / dxn;
                        const double fp_lim = fmin(fmax(step_size_mul<STRICT>(cs.alpha, f, cs.pw, cs.pw_is_third), 0.3), 2.0);
                        const double new_h = fmin(fmax(fp_lim * h, cs.min_dt), cs.max_dt);
                        if (!(dxn == dxn)) status |= VO_TRAJ_NONFINITE;
                        if (f <= 1.0) {
                            evk = VO_EV_REJECT;
                            if (h <= cs.min_dt) status |= VO_TRAJ_STUCK, ++c_stuck;
                        }
                        prev_h = h, h = new_h, prev_h_loaded = true, ctl_dirty = true;  // update_step_size, ode.rs:202-205
                    }
                    if (evk == VO_EV_STEP) {  // accept_step -> advance, ode.rs:184-188
#pragma unroll
                        for (int c = 0; c < D; ++c) xc[c] = xf[c];
                        t += dt;
                        moved = true, ++l_step;
                    } else {
                        ++l_rej;
                    }
                } else {  // Chkpt / End -> checkpoint_update, ode.rs:192-195
                    if (!prev_h_loaded) prev_h = ca.prev_h[i], prev_h_loaded = true;
                    tgt += 1, h = prev_h, ctl_dirty = true;
                    if (evk == VO_EV_END) {
                        status |= VO_TRAJ_DONE, ++c_end;
                        break;
                    }
                    ++c_chkpt;
                }
            }
            if (moved) {
#pragma unroll
                for (int c = 0; c < D; ++c) x[c * N + i] = xc[c];
                ca.t[i] = t;
            }
            if (ctl_dirty) {
                ca.h[i] = h;
                ca.prev_h[i] = prev_h;
            }
            if (dxn_set && cs.record_dx_norm) ca.dx_norm[i] = dxn;
            if (l_step) ca.n_accept[i] += l_step;
            if (l_rej) ca.n_reject[i] += l_rej;
            const uint32_t nw = ((uint32_t)tgt & VO_WORD_TGT_MASK) | (status << VO_WORD_STATUS_SHIFT);
            if (nw != word) ca.word[i] = nw;
            c_step += l_step, c_rej += l_rej;
        }
        // rotate the prefetched lane in
        word = word_n, live = (j < N) && live_n, t = t_n, h = h_n;
#pragma unroll
        for (int c = 0; c < D; ++c) xc[c] = xn[c];
#pragma unroll
        for (int q = 0; q < RHS::NP; ++q) p[q] = pn[q];
        i = j;
    }
    if (cs.count_events) {
        // block-level reduction, then at most five atomics per block, spread over VO_EV_SLOTS lines
        __shared__ unsigned sm[5];
        if (threadIdx.x < 5) sm[threadIdx.x] = 0;
        __syncthreads();
        c_step = __reduce_add_sync(0xffffffffu, c_step), c_chkpt = __reduce_add_sync(0xffffffffu, c_chkpt);
        c_rej = __reduce_add_sync(0xffffffffu, c_rej), c_end = __reduce_add_sync(0xffffffffu, c_end);
        c_stuck = __reduce_add_sync(0xffffffffu, c_stuck);
        if ((threadIdx.x & 31) == 0) {
            if (c_step) atomicAdd(&sm[0], c_step);
            if (c_chkpt) atomicAdd(&sm[1], c_chkpt);
            if (c_rej) atomicAdd(&sm[2], c_rej);
            if (c_end) atomicAdd(&sm[3], c_end);
            if (c_stuck) atomicAdd(&sm[4], c_stuck);
        }
        __syncthreads();
        if (threadIdx.x < 5 && sm[threadIdx.x]) {
            EvSlot* slot = ev + (blockIdx.x % VO_EV_SLOTS);
            unsigned long long* dst = threadIdx.x == 0   ? &slot->n_step
                                      : threadIdx.x == 1 ? &slot->n_chkpt
                                      : threadIdx.x == 2 ? &slot->n_reject
                                      : threadIdx.x == 3 ? &slot->n_end
                                                         : &slot->n_stuck;
            atomicAdd(dst, (unsigned long long)sm[threadIdx.x]);
        }
    }
}
