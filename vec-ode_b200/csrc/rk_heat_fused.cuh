// rk_heat_fused.cuh — one whole explicit Runge-Kutta step of the periodic 1-D heat equation in ONE kernel.
//
// The stage path (rk_stage.cuh) makes one pass over HBM per stage: 13 vector passes = 104 B per grid point for RK4. But
// the right-hand side is a 3-point stencil, so the new value of a grid point depends on x0 at its s nearest neighbours on
// each side only: a CTA that holds a tile of L points can run ALL s stages on it without leaving the SM and still gets the
// inner L - 2s points right. Traffic drops to read x0 once, write next_x once: 16 B per grid point (24 B with x_err).
//
//   * thread t owns the HF_PPT points p = t + 256 q of the tile: x0 and every K_j of those points stay in REGISTERS
//     (the stage argument x0 + dt sum_j a_ij K_j is pointwise, rk.rs:121-124);
//   * only the stage argument travels: each stage writes it to one of two shared-memory lines, one __syncthreads(), and
//     the stencil reads the two neighbours from there (consecutive lanes read consecutive doubles: no bank conflicts);
//   * the points within s of a tile end see wrong neighbours from stage to stage and are simply not stored: tiles overlap
//     by 2s points (0.4 % redundant work for RK4 at L = 2048), the periodic wrap of the grid is index arithmetic on the
//     loads;
//   * the next tile's x0 is loaded into registers while the current one is integrated.
//
// Per point the operations and their order are those of stage_heat_kernel / heat_tail_point, so in STRICT arithmetic the
// result is bit-identical to the stage path (and to the reference's un-fused code).
#pragma once
#include "rk_stage.cuh"

constexpr int HF_THREADS = 256;

template <int S, bool STRICT, int PPT>
__global__ void __launch_bounds__(HF_THREADS, (S <= 4 ? 2 : 1))
    heat_fused_step_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ TableauDev tb, const __grid_constant__ StageArgs sa,
                           double kappa, double* __restrict__ next_x, double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int L = HF_THREADS * PPT, T = L - 2 * S;  // tile length, owned points per tile
    __shared__ double buf[2][L + 2];                     // stage arguments, one guard cell at each end
    const int tid = threadIdx.x;
    const int64_t n_tiles = (d + T - 1) / T;
    if (tid == 0) buf[0][0] = buf[0][L + 1] = buf[1][0] = buf[1][L + 1] = 0.0;
    auto gidx = [&](int64_t tile, int p) {  // global index of local point p of a tile, periodic
        int64_t g = tile * T - S + p;
        if (g < 0) g += d;
        else if (g >= d) g -= d;
        return g;
    };
    double xn[PPT];
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) {
#pragma unroll
        for (int q = 0; q < PPT; ++q) xn[q] = x0[gidx(tile, tid + HF_THREADS * q)];
    }
    for (; tile < n_tiles; tile += gridDim.x) {
        double xc[PPT], K[S][PPT];
        __syncthreads();  // the previous tile's last stencil reads are done
#pragma unroll
        for (int q = 0; q < PPT; ++q) xc[q] = xn[q], buf[0][1 + tid + HF_THREADS * q] = xc[q];
        __syncthreads();
        if (tile + gridDim.x < n_tiles) {  // next tile of this CTA: in flight while this one is integrated
#pragma unroll
            for (int q = 0; q < PPT; ++q) xn[q] = x0[gidx(tile + gridDim.x, tid + HF_THREADS * q)];
        }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {  // K_0 = f(x0): kappa * ((u_{j-1} + u_{j+1}) - 2 u_j)
            const int p = tid + HF_THREADS * q;
            K[0][q] = A::mul(kappa, A::sub(A::add(buf[0][p], buf[0][p + 2]), A::mul(2.0, xc[q])));
        }
#pragma unroll
        for (int i = 1; i < S; ++i) {
            const double* row = &tb.ac[i * S];
            double* line = buf[i & 1];
            double xs[PPT];
#pragma unroll
            for (int q = 0; q < PPT; ++q) {  // stage argument (rk.rs:121-124), same order as stage_heat_kernel
                double acc;
                if (STRICT) {
                    acc = A::mul(row[0], K[0][q]);
#pragma unroll
                    for (int j = 1; j < i; ++j) acc = A::axpy(acc, row[j], K[j][q]);
                } else {
                    acc = 0.0;
#pragma unroll
                    for (int j = 0; j < i; ++j)
                        if (row[j] != 0.0) acc = fma(row[j], K[j][q], acc);
                }
                xs[q] = A::add(A::mul(acc, sa.dt), xc[q]);
                line[1 + tid + HF_THREADS * q] = xs[q];
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                const int p = tid + HF_THREADS * q;
                K[i][q] = A::mul(kappa, A::sub(A::add(line[p], line[p + 2]), A::mul(2.0, xs[q])));
            }
        }
        // b / b_err combinations (rk.rs:131-151) for the owned points of the tile
        const int64_t own_lo = tile * T, own_hi = own_lo + T < d ? own_lo + T : d;
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            const int p = tid + HF_THREADS * q;
            const int64_t g = own_lo - S + p;
            if (p >= S && g < own_hi) {
                double kj[8], ox, oe = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) kj[j] = j < S - 1 ? K[j < S - 1 ? j : 0][q] : 0.0;
                heat_tail_point<STRICT>(sa, kj, K[S - 1][q], xc[q], &ox, &oe);
                next_x[g] = ox;
                if (sa.use_err) x_err[g] = oe;
            }
        }
    }
}
