// rk_stage.cuh — stage-granular Runge-Kutta kernels: ONE fused kernel per stage,
//     K_i = f(t + c_i h, x0 + h * sum_{j<i} a_ij K_j)                      (src/base/rk.rs:118-128)
// and one for the tail, which evaluates the last stage in registers and forms
//     next_x = x0 + h * sum_j b_j K_j   [, X_berr and x_err = X_b - X_berr] (src/base/rk.rs:131-151)
// without ever writing K_{s-1} or the stage argument xf to memory. Per stage the kernel reads x0 and the
// K_j it needs once and writes K_i once: (s-1)(s-2)/2 + 3s - 1 vector passes per step instead of the ~124 the
// reference's LinearCombination chain makes for s = 6. This is the path for large states (heat equation,
// N = 1, d = 2^26) and the generic fallback for any (RHS, tableau) the register-resident kernel does not cover.
//
// Two RHS shapes:
//  * pointwise families (small d, SoA [d][N]): one thread per trajectory, components in registers;
//  * HEAT1D (3-point periodic stencil along the component axis): one CTA per tile of TILE consecutive
//    elements; the stage argument of the tile plus a one-cell halo is built in shared memory from 128-bit
//    global loads, then the stencil reads shared memory only.
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
template <bool STRICT> __device__ __forceinline__ double stage_elem(const StageArgs& sa, const double* k, int n, int64_t e, double x0, double dt) {
    using A = Ar<STRICT>;
    double acc;
    if (STRICT) {
        acc = A::mul(k[0], sa.K[0][e]);
        for (int j = 1; j < n; ++j) acc = A::axpy(acc, k[j], sa.K[j][e]);
    } else {
        acc = 0.0;
        for (int j = 0; j < n; ++j)
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

// ---- HEAT1D, single large state (N == 1): shared-memory tile with one-cell halo -----------------------
constexpr int HEAT_THREADS = 256;
constexpr int HEAT_VEC = 4;                             // elements per thread (two double2)
constexpr int HEAT_TILE = HEAT_THREADS * HEAT_VEC;      // 1024 elements = 8 KiB per tile

template <bool STRICT, bool TAIL>
__global__ void __launch_bounds__(HEAT_THREADS) stage_heat_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ StageArgs sa,
                                                                  double kappa, double* __restrict__ k_out, double* __restrict__ next_x,
                                                                  double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    __shared__ double tile[HEAT_TILE + 2];
    const int nterm = sa.i;
    const int64_t n_tiles = (d + HEAT_TILE - 1) / HEAT_TILE;
    for (int64_t tix = blockIdx.x; tix < n_tiles; tix += gridDim.x) {
        const int64_t base = tix * HEAT_TILE;
        const int64_t e0 = base + (int64_t)threadIdx.x * HEAT_VEC;
        double xc[HEAT_VEC], kj[VO_MAX_STAGES > 8 ? 8 : VO_MAX_STAGES][HEAT_VEC];
        const bool full = e0 + HEAT_VEC <= d;
        // own elements of x0 and of every K_j (kept in registers for the tail combination when s <= 8)
        if (full) {
            const double2 a = reinterpret_cast<const double2*>(x0 + e0)[0], b = reinterpret_cast<const double2*>(x0 + e0)[1];
            xc[0] = a.x, xc[1] = a.y, xc[2] = b.x, xc[3] = b.y;
        } else {
#pragma unroll
            for (int v = 0; v < HEAT_VEC; ++v) xc[v] = e0 + v < d ? x0[e0 + v] : 0.0;
        }
        const int nload = TAIL ? sa.s - 1 : nterm;  // the tail needs K_0..K_{s-2} for the b-combination
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nload) {
                const bool needed = STRICT || TAIL || sa.a[j] != 0.0;
                if (needed && full) {
                    const double2 a = reinterpret_cast<const double2*>(sa.K[j] + e0)[0], b = reinterpret_cast<const double2*>(sa.K[j] + e0)[1];
                    kj[j][0] = a.x, kj[j][1] = a.y, kj[j][2] = b.x, kj[j][3] = b.y;
                } else {
#pragma unroll
                    for (int v = 0; v < HEAT_VEC; ++v) kj[j][v] = (needed && e0 + v < d) ? sa.K[j][e0 + v] : 0.0;
                }
            }
        }
        // stage argument xs = x0 + dt * sum a_j K_j for own elements -> shared tile
#pragma unroll
        for (int v = 0; v < HEAT_VEC; ++v) {
            double acc;
            if (nterm == 0) {
                acc = xc[v];
            } else {
                if (STRICT) {
                    acc = A::mul(sa.a[0], kj[0][v]);
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j < nterm) acc = A::axpy(acc, sa.a[j], kj[j][v]);
                } else {
                    acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < nterm && sa.a[j] != 0.0) acc = fma(sa.a[j], kj[j][v], acc);
                }
                acc = A::add(A::mul(acc, sa.dt), xc[v]);
            }
            if (e0 + v < d) tile[1 + threadIdx.x * HEAT_VEC + v] = acc;  // cells past d belong to the halo writer
        }
        // halo cells (periodic): two threads rebuild the neighbours' stage argument from global memory
        if (threadIdx.x < 2) {
            const int64_t last = min(base + (int64_t)HEAT_TILE, d) - 1;
            int64_t e = threadIdx.x == 0 ? base - 1 : last + 1;
            if (e < 0) e = d - 1;
            if (e >= d) e = 0;
            const double xh = x0[e];
            tile[threadIdx.x == 0 ? 0 : (int)(last - base) + 2] = nterm == 0 ? xh : stage_elem<STRICT>(sa, sa.a, nterm, e, xh, sa.dt);
        }
        __syncthreads();
        double kl[HEAT_VEC];
#pragma unroll
        for (int v = 0; v < HEAT_VEC; ++v) {
            const int q = 1 + threadIdx.x * HEAT_VEC + v;
            // kappa * ((u_{j-1} + u_{j+1}) - 2 u_j)
            kl[v] = A::mul(kappa, A::sub(A::add(tile[q - 1], tile[q + 1]), A::mul(2.0, tile[q])));
        }
        __syncthreads();
        if (!TAIL) {
            if (full) {
                reinterpret_cast<double2*>(k_out + e0)[0] = make_double2(kl[0], kl[1]);
                reinterpret_cast<double2*>(k_out + e0)[1] = make_double2(kl[2], kl[3]);
            } else {
#pragma unroll
                for (int v = 0; v < HEAT_VEC; ++v)
                    if (e0 + v < d) k_out[e0 + v] = kl[v];
            }
            continue;
        }
        double ox[HEAT_VEC], oe[HEAT_VEC];
        const int s = sa.s;
#pragma unroll
        for (int v = 0; v < HEAT_VEC; ++v) {
            double xb, xbe = 0.0;
            if (STRICT) {
                xb = A::mul(sa.b[0], s == 1 ? kl[v] : kj[0][v]);
#pragma unroll
                for (int j = 1; j < 8; ++j)
                    if (j < s) xb = A::axpy(xb, sa.b[j], j == s - 1 ? kl[v] : kj[j][v]);
            } else {
                xb = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < s && sa.b[j] != 0.0) xb = fma(sa.b[j], j == s - 1 ? kl[v] : kj[j][v], xb);
            }
            xb = A::add(A::mul(xb, sa.dt), xc[v]);
            if (sa.use_err) {
                if (STRICT) {
                    xbe = A::mul(sa.b_err[0], s == 1 ? kl[v] : kj[0][v]);
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j < s) xbe = A::axpy(xbe, sa.b_err[j], j == s - 1 ? kl[v] : kj[j][v]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < s && sa.b_err[j] != 0.0) xbe = fma(sa.b_err[j], j == s - 1 ? kl[v] : kj[j][v], xbe);
                }
                xbe = A::add(A::mul(xbe, sa.dt), xc[v]);
                ox[v] = xbe, oe[v] = A::sub(xb, xbe);
            } else {
                ox[v] = xb;
            }
        }
        if (full) {
            reinterpret_cast<double2*>(next_x + e0)[0] = make_double2(ox[0], ox[1]);
            reinterpret_cast<double2*>(next_x + e0)[1] = make_double2(ox[2], ox[3]);
            if (sa.use_err) {
                reinterpret_cast<double2*>(x_err + e0)[0] = make_double2(oe[0], oe[1]);
                reinterpret_cast<double2*>(x_err + e0)[1] = make_double2(oe[2], oe[3]);
            }
            if (k_out) {
                reinterpret_cast<double2*>(k_out + e0)[0] = make_double2(kl[0], kl[1]);
                reinterpret_cast<double2*>(k_out + e0)[1] = make_double2(kl[2], kl[3]);
            }
        } else {
#pragma unroll
            for (int v = 0; v < HEAT_VEC; ++v)
                if (e0 + v < d) {
                    next_x[e0 + v] = ox[v];
                    if (sa.use_err) x_err[e0 + v] = oe[v];
                    if (k_out) k_out[e0 + v] = kl[v];
                }
        }
    }
}

// HEAT1D for ensembles (N > 1): neighbours along the component axis are N elements apart; plain loads
// (the re-reads of neighbouring rows hit L2). One thread per element.
template <bool STRICT, bool TAIL>
__global__ void __launch_bounds__(256) stage_heat_ens_kernel(const double* __restrict__ x0, int64_t d, int64_t N, const __grid_constant__ StageArgs sa,
                                                             double kappa, double* __restrict__ k_out, double* __restrict__ next_x,
                                                             double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= d * N) return;
    const int64_t c = e / N;
    const int64_t el = (c == 0 ? d - 1 : c - 1) * N + (e - c * N), er = (c + 1 == d ? 0 : c + 1) * N + (e - c * N);
    const double xc = x0[e];
    double um, ul, ur;
    if (sa.i == 0) um = xc, ul = x0[el], ur = x0[er];
    else um = stage_elem<STRICT>(sa, sa.a, sa.i, e, xc, sa.dt), ul = stage_elem<STRICT>(sa, sa.a, sa.i, el, x0[el], sa.dt), ur = stage_elem<STRICT>(sa, sa.a, sa.i, er, x0[er], sa.dt);
    const double kl = A::mul(kappa, A::sub(A::add(ul, ur), A::mul(2.0, um)));
    if (!TAIL) {
        k_out[e] = kl;
        return;
    }
    const int s = sa.s;
    double xb, xbe = 0.0;
    if (STRICT) {
        xb = A::mul(sa.b[0], s == 1 ? kl : sa.K[0][e]);
        for (int j = 1; j < s; ++j) xb = A::axpy(xb, sa.b[j], j == s - 1 ? kl : sa.K[j][e]);
    } else {
        xb = 0.0;
        for (int j = 0; j < s; ++j)
            if (sa.b[j] != 0.0) xb = fma(sa.b[j], j == s - 1 ? kl : sa.K[j][e], xb);
    }
    xb = A::add(A::mul(xb, sa.dt), xc);
    if (sa.use_err) {
        if (STRICT) {
            xbe = A::mul(sa.b_err[0], s == 1 ? kl : sa.K[0][e]);
            for (int j = 1; j < s; ++j) xbe = A::axpy(xbe, sa.b_err[j], j == s - 1 ? kl : sa.K[j][e]);
        } else {
            for (int j = 0; j < s; ++j)
                if (sa.b_err[j] != 0.0) xbe = fma(sa.b_err[j], j == s - 1 ? kl : sa.K[j][e], xbe);
        }
        xbe = A::add(A::mul(xbe, sa.dt), xc);
        next_x[e] = xbe, x_err[e] = A::sub(xb, xbe);
    } else {
        next_x[e] = xb;
    }
    if (k_out) k_out[e] = kl;
}
