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
    // uniform (lock-step) control: every trajectory shares these scalars, nothing is loaded or stored
    double u_t, u_h, u_prev_h;
    int u_tgt;
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
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = 0.0;
#pragma unroll
        for (int j = 0; j < SM; ++j)
            if (j < n && k[j] != 0.0) {  // uniform predicate on a constant-bank value
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

template <class RHS, int S, bool STRICT, bool UNIFORM>
__global__ void __launch_bounds__(128) rk_small_kernel(double* __restrict__ x, int64_t N, const __grid_constant__ TableauDev tb,
                                                       const __grid_constant__ RhsParams rp, const CtlArrays ca,
                                                       const __grid_constant__ CtlShared cs, EvSlot* __restrict__ ev) {
    constexpr int D = RHS::D;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned c_step = 0, c_chkpt = 0, c_rej = 0, c_end = 0, c_stuck = 0;
    if (i < N) {
        uint32_t word = UNIFORM ? (uint32_t)cs.u_tgt : ca.word[i];
        if (!((word >> VO_WORD_STATUS_SHIFT) & VO_TRAJ_DONE)) {
            double xc[D], p[RHS::NP];
#pragma unroll
            for (int c = 0; c < D; ++c) xc[c] = x[c * N + i];
            load_params<RHS::NP>(rp, i, p);
            double t = UNIFORM ? cs.u_t : ca.t[i];
            double h = UNIFORM ? cs.u_h : ca.h[i];
            double prev_h = UNIFORM ? cs.u_prev_h : 0.0;
            bool prev_h_loaded = UNIFORM, ctl_dirty = false, moved = false;
            int tgt = (int)(word & VO_WORD_TGT_MASK);
            uint32_t status = word >> VO_WORD_STATUS_SHIFT;
            double dxn = 0.0;
            bool dxn_set = false;
            const double* tl = cs.n_tlist > VO_INLINE_TLIST ? cs.t_list : cs.t_list_inline;
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
                    if (!UNIFORM && cs.adaptive) {  // handle_step_adaptive, ode.rs:311-334
                        dxn = err_norm<STRICT, D>(xe, cs.norm_kind), dxn_set = true;
                        const double f = cs.rtol / dxn;
                        const double fp_lim = fmin(fmax(cs.alpha * pow(f, cs.pw), 0.3), 2.0);
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
                        moved = true, ++c_step;
                    } else {
                        ++c_rej;
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
            }
            if (!UNIFORM) {
                if (moved) ca.t[i] = t;
                if (ctl_dirty) {
                    ca.h[i] = h;
                    ca.prev_h[i] = prev_h;
                }
                if (dxn_set) ca.dx_norm[i] = dxn;
                if (c_step) ca.n_accept[i] += c_step;
                if (c_rej) ca.n_reject[i] += c_rej;
                const uint32_t nw = ((uint32_t)tgt & VO_WORD_TGT_MASK) | (status << VO_WORD_STATUS_SHIFT);
                if (nw != word) ca.word[i] = nw;
            }
        }
    }
    if (!UNIFORM && cs.count_events) {
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
