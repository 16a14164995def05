// rk_stage_pointwise.cuh — the stage-path kernel for pointwise RHS families (see rk_stage.cuh for the path as a whole).
// Kept free of host headers so that the run-time compiled user RHS modules (nvrtc_rhs.cu) can instantiate it too.
#pragma once
#include "common.cuh"
#include "rhs.cuh"

struct StageArgs {
    const double* K[VO_MAX_STAGES];  // stage derivative buffers K_0 .. K_{s-1}
    double a[VO_MAX_STAGES];         // row i of the tableau (a_i0 .. a_i,i-1)
    double b[VO_MAX_STAGES];         // tail only
    double b_err[VO_MAX_STAGES];     // tail only
    int i;                           // stage index (number of terms in the row)
    int s;                           // number of stages
    int use_err;                     // tail: also produce X_berr / x_err
    double t_i;                      // t + c_i * dt (lock-step control)
    double dt;
    double c_i;                      // ac[i][i]
    // per-trajectory control (ensemble stage path): when non-NULL they override t_i / dt, and lanes whose
    // event is not Step are masked out.
    const double* tv;
    const double* dtv;
    const uint8_t* evv;
};

// acc = (sum_{j<n} k_j v_j) * dt + x0 for one element, reference order; FAST skips zero coefficients.
// The first eight stage derivatives are fetched by an unrolled, predicated block of loads before any arithmetic, so they are all in
// flight together (a loop with a run-time trip count made every load wait for the previous term's use: the generic stage path ran
// at a quarter of the HBM rate); tableaux with more than eight stages finish in a plain loop.
template <bool STRICT> __device__ __forceinline__ double stage_elem(const StageArgs& sa, const double* k, int n, int64_t e, double x0, double dt) {
    using A = Ar<STRICT>;
    double kv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) kv[j] = (j < n && (STRICT || k[j] != 0.0)) ? sa.K[j][e] : 0.0;
    double acc;
    if (STRICT) {
        acc = A::mul(k[0], kv[0]);
#pragma unroll
        for (int j = 1; j < 8; ++j)
            if (j < n) acc = A::axpy(acc, k[j], kv[j]);
        for (int j = 8; j < n; ++j) acc = A::axpy(acc, k[j], sa.K[j][e]);
    } else {
        acc = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < n && k[j] != 0.0) acc = fma(k[j], kv[j], acc);
        for (int j = 8; j < n; ++j)
            if (k[j] != 0.0) acc = fma(k[j], sa.K[j][e], acc);
    }
    return A::add(A::mul(acc, dt), x0);
}

// ---- pointwise RHS ------------------------------------------------------------------------------------
// STAGE0: K_0 = f(t, x0).  !TAIL: K_i.  TAIL: last stage in registers + b / b_err combinations.
template <class RHS, bool STRICT, bool TAIL>
__global__ void __launch_bounds__(128) stage_pointwise_kernel(const double* __restrict__ x0, int64_t N, const __grid_constant__ StageArgs sa,
                                                              const __grid_constant__ RhsParams rp, double* __restrict__ k_out,
                                                              double* __restrict__ next_x, double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int D = RHS::D;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (sa.evv && sa.evv[i] != VO_EV_STEP) return;  // masked lane: Chkpt / End / Done
    const double dt = sa.dtv ? sa.dtv[i] : sa.dt;
    const double t_i = sa.tv ? A::add(sa.tv[i], A::mul(sa.c_i, dt)) : sa.t_i;  // rk.rs:119
    double xc[D], xs[D], kl[D], p[RHS::NP];
#pragma unroll
    for (int c = 0; c < D; ++c) xc[c] = x0[c * N + i];
    load_params<RHS::NP>(rp, i, p);
    if (sa.i == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) xs[c] = xc[c];
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) xs[c] = stage_elem<STRICT>(sa, sa.a, sa.i, c * N + i, xc[c], dt);
    }
    RHS::template eval<STRICT>(t_i, xs, kl, p);
    if (!TAIL) {
#pragma unroll
        for (int c = 0; c < D; ++c) k_out[c * N + i] = kl[c];
        return;
    }
    // tail: sum_j b_j K_j with K_{s-1} = kl held in registers (same left-to-right order as lc.rs:20-35)
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const int64_t e = c * N + i;
        double xb, xbe = 0.0;
        const int s = sa.s;
        if (STRICT) {
            xb = A::mul(sa.b[0], s == 1 ? kl[c] : sa.K[0][e]);
            for (int j = 1; j < s; ++j) xb = A::axpy(xb, sa.b[j], j == s - 1 ? kl[c] : sa.K[j][e]);
        } else {
            xb = 0.0;
            for (int j = 0; j < s; ++j)
                if (sa.b[j] != 0.0) xb = fma(sa.b[j], j == s - 1 ? kl[c] : sa.K[j][e], xb);
        }
        xb = A::add(A::mul(xb, dt), xc[c]);
        if (sa.use_err) {
            if (STRICT) {
                xbe = A::mul(sa.b_err[0], s == 1 ? kl[c] : sa.K[0][e]);
                for (int j = 1; j < s; ++j) xbe = A::axpy(xbe, sa.b_err[j], j == s - 1 ? kl[c] : sa.K[j][e]);
            } else {
                for (int j = 0; j < s; ++j)
                    if (sa.b_err[j] != 0.0) xbe = fma(sa.b_err[j], j == s - 1 ? kl[c] : sa.K[j][e], xbe);
            }
            xbe = A::add(A::mul(xbe, dt), xc[c]);
            next_x[e] = xbe;                 // the reference propagates X_berr (rk.rs:142-146)
            x_err[e] = A::sub(xb, xbe);      // rk.rs:147
        } else {
            next_x[e] = xb;
        }
        if (k_out) k_out[e] = kl[c];
    }
}
