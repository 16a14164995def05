// rk_stage_stencil.cuh — the stage path (rk_stage.cuh) for a USER-DEFINED stencil right-hand side on one large periodic grid
// state: the closure `f(t, &x, &mut dx)` of src/base/rk.rs:97 when V is a grid (`Array1<f64>`, src/impls/ndarray.rs:8-33) and dx_j
// depends on x_{j-R} .. x_{j+R}. The closure's statements cross the C ABI as source (vo_rhs_create_custom_stencil) and are compiled
// at run time into this kernel. One fused kernel per RK stage, like the compiled-in heat kernels: a CTA builds the stage argument
//     x0 + dt * sum_{j<i} a_ij K_j          (rk.rs:121-124, reference order, zeros kept in STRICT arithmetic)
// of a tile of 256 grid points plus R halo points per side in shared memory (periodic wrap), then every thread applies the stencil
// to its point from shared memory and writes K_i — or, in the tail launch, forms next_x / x_err from K_0..K_{s-2} and the
// last stage held in a register (rk.rs:131-151). Each array is read once per stage (plus 2R halo points per tile).
// Kept free of host headers: nvrtc_rhs.cu instantiates it with the generated functor.
#pragma once
#include "rk_stage_pointwise.cuh"

constexpr int ST_THREADS = 256;

template <class ST, bool STRICT, bool TAIL>
__global__ void __launch_bounds__(ST_THREADS) stage_stencil_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ StageArgs sa,
                                                                   const __grid_constant__ RhsParams rp, double* __restrict__ k_out,
                                                                   double* __restrict__ next_x, double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int R = ST::R, W = ST_THREADS + 2 * R;
    __shared__ double tile[W];
    double p[ST::NP];
#pragma unroll
    for (int q = 0; q < ST::NP; ++q) p[q] = rp.shared[q];
    const int64_t n_tiles = (d + ST_THREADS - 1) / ST_THREADS;
    const int s_ = sa.s;
    const int nload = TAIL ? (s_ - 1 < 8 ? s_ - 1 : 8) : (sa.i < 8 ? sa.i : 8);  // the tail also needs K_0..K_{s-2} for the b-combination
    for (int64_t ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
        const int64_t base = ti * ST_THREADS;
        // own point: x0 and the stage derivatives it needs, fetched together (kept for the tail's b-combination)
        int64_t e = base + threadIdx.x;
        const bool act = e < d;
        while (e >= d) e -= d;  // past the end of the grid (last tile only): a periodic image, needed as halo by the last points
        const double xc = x0[e];
        double kj[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) kj[j] = (j < nload && (STRICT || TAIL || sa.a[j] != 0.0)) ? sa.K[j][e] : 0.0;
        {
            double acc = xc;
            if (sa.i > 0) {
                if (STRICT) {
                    acc = A::mul(sa.a[0], kj[0]);
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j < sa.i) acc = A::axpy(acc, sa.a[j], kj[j]);
                    for (int j = 8; j < sa.i; ++j) acc = A::axpy(acc, sa.a[j], sa.K[j][e]);
                } else {
                    acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < sa.i && sa.a[j] != 0.0) acc = fma(sa.a[j], kj[j], acc);
                    for (int j = 8; j < sa.i; ++j)
                        if (sa.a[j] != 0.0) acc = fma(sa.a[j], sa.K[j][e], acc);
                }
                acc = A::add(A::mul(acc, sa.dt), xc);  // rk.rs:123-124
            }
            tile[R + threadIdx.x] = acc;
        }
        if (threadIdx.x < 2 * R) {  // the 2R halo points of the tile (periodic), rebuilt from global memory: same operations, same order
            const int k = threadIdx.x < R ? threadIdx.x : ST_THREADS + threadIdx.x;
            int64_t eh = base - R + k;
            while (eh < 0) eh += d;
            while (eh >= d) eh -= d;
            const double xh = x0[eh];
            tile[k] = sa.i == 0 ? xh : stage_elem<STRICT>(sa, sa.a, sa.i, eh, xh, sa.dt);
        }
        __syncthreads();
        if (act) {
            double u[2 * R + 1];
#pragma unroll
            for (int k = 0; k < 2 * R + 1; ++k) u[k] = tile[threadIdx.x + k];
            const double kl = ST::template eval<STRICT>(sa.t_i, (long long)e, (long long)d, u, p);
            if (!TAIL) {
                k_out[e] = kl;
            } else {  // sum_j b_j K_j with K_{s-1} = kl in a register, left to right (lc.rs:20-35)
                const int s = sa.s;
                auto kval = [&](int j) { return j == s - 1 ? kl : (j < 8 ? kj[j] : sa.K[j][e]); };
                double xb, xbe = 0.0;
                if (STRICT) {
                    xb = A::mul(sa.b[0], kval(0));
#pragma unroll
                    for (int j = 1; j < 8; ++j)
                        if (j < s) xb = A::axpy(xb, sa.b[j], kval(j));
                    for (int j = 8; j < s; ++j) xb = A::axpy(xb, sa.b[j], kval(j));
                } else {
                    xb = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < s && sa.b[j] != 0.0) xb = fma(sa.b[j], kval(j), xb);
                    for (int j = 8; j < s; ++j)
                        if (sa.b[j] != 0.0) xb = fma(sa.b[j], kval(j), xb);
                }
                xb = A::add(A::mul(xb, sa.dt), xc);
                if (sa.use_err) {
                    if (STRICT) {
                        xbe = A::mul(sa.b_err[0], kval(0));
#pragma unroll
                        for (int j = 1; j < 8; ++j)
                            if (j < s) xbe = A::axpy(xbe, sa.b_err[j], kval(j));
                        for (int j = 8; j < s; ++j) xbe = A::axpy(xbe, sa.b_err[j], kval(j));
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < s && sa.b_err[j] != 0.0) xbe = fma(sa.b_err[j], kval(j), xbe);
                        for (int j = 8; j < s; ++j)
                            if (sa.b_err[j] != 0.0) xbe = fma(sa.b_err[j], kval(j), xbe);
                    }
                    xbe = A::add(A::mul(xbe, sa.dt), xc);
                    next_x[e] = xbe;             // the reference propagates X_berr (rk.rs:142-146)
                    x_err[e] = A::sub(xb, xbe);  // rk.rs:147
                } else {
                    next_x[e] = xb;
                }
                if (k_out) k_out[e] = kl;
            }
        }
        __syncthreads();  // the tile is rebuilt in the next iteration
    }
}
