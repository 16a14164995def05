// rk_heat_fused.cuh — one whole explicit Runge-Kutta step of the periodic 1-D heat equation in ONE kernel.
//
// The stage path (rk_stage.cuh) makes one pass over HBM per stage: 13 vector passes = 104 B per grid point for RK4. But
// the right-hand side is a 3-point stencil, so the new value of a grid point depends on x0 at its s nearest neighbours on
// each side only: a CTA that holds a tile of L points can run ALL s stages on it without leaving the SM and still gets the
// inner L - 2s points right. Traffic drops to read x0 once, write next_x once: 16 B per grid point (24 B with x_err).
//
//   * thread t owns PPT CONSECUTIVE points of the tile: x0 and every K_j of those points stay in REGISTERS (the stage
//     argument x0 + dt sum_j a_ij K_j is pointwise, rk.rs:121-124), and so do the stencil neighbours of all but the two end
//     points of the thread's run;
//   * only those two end values travel, once per stage: each thread publishes the first and last stage argument of its
//     run in shared memory (two lines alternate, one __syncthreads() per stage) and reads its neighbours' — 3 shared-memory
//     instructions per thread per stage instead of 3 per POINT, which is what bounded the first version of this kernel;
//   * the points within s of a tile end see wrong neighbours from stage to stage and are simply not stored: tiles overlap
//     by 2s points (0.4 % redundant work for RK4 at L = 2048);
//   * x0 tiles arrive through a 3-deep TMA pipeline (one cp.async.bulk of the whole tile per stage buffer, mbarrier
//     completion; tile_pipe.cuh) so that ~100 KB per SM are in flight however few registers are free; only the tiles that
//     touch the two ends of the periodic grid are filled by ordinary wrapped loads.
//
// Per point the operations and their order are those of stage_heat_kernel / heat_tail_point, so in STRICT arithmetic the
// result is bit-identical to the stage path (and to the reference's un-fused code).
#pragma once
#include "rk_stage.cuh"

constexpr int HF_THREADS = 256;
constexpr int HF_NST = 3;  // tiles in flight per CTA

template <int S, bool STRICT, int PPT>
__global__ void __launch_bounds__(HF_THREADS, 2)
    heat_fused_step_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ TableauDev tb, const __grid_constant__ StageArgs sa,
                           double kappa, double* __restrict__ next_x, double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int HS = (S + 1) & ~1;                     // halo per side, even so that tile windows start on 16-byte boundaries
    constexpr int L = HF_THREADS * PPT, T = L - 2 * HS;  // tile length, owned points per tile
    static_assert(PPT % 2 == 0, "runs are moved two points at a time");
    extern __shared__ __align__(128) double sbuf[];      // [HF_NST][L]: x0 tiles
    __shared__ double e_first[2][HF_THREADS + 2], e_last[2][HF_THREADS + 2];  // end values of every thread's run, guard cell at each end
    __shared__ __align__(8) uint64_t full[HF_NST];
    const int tid = threadIdx.x;
    const int64_t n_tiles = (d + T - 1) / T, G = gridDim.x, first = blockIdx.x;
    const int64_t my_count = first < n_tiles ? (n_tiles - first + G - 1) / G : 0;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 2; ++k) e_first[k][0] = e_first[k][HF_THREADS + 1] = e_last[k][0] = e_last[k][HF_THREADS + 1] = 0.0;
#pragma unroll
        for (int k = 0; k < HF_NST; ++k) pipe::mbar_init(&full[k], 1);
        pipe::fence_mbar_init();
    }
    __syncthreads();
    auto window = [&](int64_t k) { return (first + k * G) * T - HS; };               // global index of local point 0 of the CTA's k-th tile
    auto inside = [&](int64_t k) { return window(k) >= 0 && window(k) + L <= d; };  // no periodic wrap inside the tile
    auto issue = [&](int64_t k) {  // thread 0: the whole tile in one bulk copy
        if (!inside(k)) return;
        const int st = (int)(k % HF_NST);
        pipe::mbar_expect_tx(&full[st], (uint32_t)(L * sizeof(double)));
        pipe::bulk_g2s(sbuf + (size_t)st * L, x0 + window(k), L * sizeof(double), &full[st]);
    };
    if (tid == 0)
        for (int64_t k = 0; k < my_count && k < HF_NST; ++k) issue(k);
    uint32_t phase = 0;  // bit st: parity of the next completion of stage buffer st
    // K = kappa * ((u_{j-1} + u_{j+1}) - 2 u_j) over the thread's run, the two outer neighbours through shared memory
    auto stencil = [&](const double (&v)[PPT], double (&k_out)[PPT], int par) {
        e_first[par][tid + 1] = v[0], e_last[par][tid + 1] = v[PPT - 1];
        __syncthreads();
        const double left = e_last[par][tid], right = e_first[par][tid + 2];
#pragma unroll
        for (int q = 0; q < PPT; ++q)
            k_out[q] = A::mul(kappa, A::sub(A::add(q == 0 ? left : v[q == 0 ? 0 : q - 1], q == PPT - 1 ? right : v[q == PPT - 1 ? q : q + 1]), A::mul(2.0, v[q])));
    };
    for (int64_t k = 0; k < my_count; ++k) {
        const int st = (int)(k % HF_NST);
        const int64_t tile = first + k * G, w0 = window(k);
        const int64_t base = w0 + (int64_t)tid * PPT;  // global index of the first point of this thread's run (before wrapping)
        double* stage = sbuf + (size_t)st * L;
        if (inside(k)) {
            pipe::mbar_wait(&full[st], (phase >> st) & 1u);
            phase ^= 1u << st;
        } else {  // a tile at an end of the periodic grid: wrapped loads, all threads
#pragma unroll
            for (int q = 0; q < PPT; ++q) {
                int64_t g = w0 + tid + HF_THREADS * q;
                if (g < 0) g += d;
                else if (g >= d) g -= d;
                stage[tid + HF_THREADS * q] = x0[g];
            }
            __syncthreads();
        }
        double xc[PPT], K[S][PPT];
#pragma unroll
        for (int q = 0; q < PPT; q += 2) {
            const double2 v = *reinterpret_cast<const double2*>(stage + tid * PPT + q);
            xc[q] = v.x, xc[q + 1] = v.y;
        }
        stencil(xc, K[0], 0);  // K_0 = f(x0); its barrier also says that every thread has taken its run out of the stage buffer
        if (tid == 0 && k + HF_NST < my_count) issue(k + HF_NST);
#pragma unroll
        for (int i = 1; i < S; ++i) {
            const double* row = &tb.ac[i * S];
            double xs[PPT];
            if (STRICT) {  // stage argument (rk.rs:121-124), same order as stage_heat_kernel
#pragma unroll
                for (int q = 0; q < PPT; ++q) xs[q] = A::mul(row[0], K[0][q]);
#pragma unroll
                for (int j = 1; j < i; ++j)
#pragma unroll
                    for (int q = 0; q < PPT; ++q) xs[q] = A::axpy(xs[q], row[j], K[j][q]);
            } else {
#pragma unroll
                for (int q = 0; q < PPT; ++q) xs[q] = 0.0;
#pragma unroll
                for (int j = 0; j < i; ++j)
                    if (row[j] != 0.0) {  // uniform: whole runs of FMAs are skipped for the zeros of the tableau
#pragma unroll
                        for (int q = 0; q < PPT; ++q) xs[q] = fma(row[j], K[j][q], xs[q]);
                    }
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) xs[q] = A::add(A::mul(xs[q], sa.dt), xc[q]);
            stencil(xs, K[i], i & 1);
        }
        // b / b_err combinations (rk.rs:131-151) for the owned points of the tile
        const int64_t own_hi = tile * T + T < d ? tile * T + T : d;
        // (heat_tail_point's operations with the stage count known at compile time and the zero tests hoisted out of the run)
        double ox[PPT], oe[PPT];
        auto weigh = [&](const double* w, double (&out)[PPT]) {  // (sum_j w_j K_j) * dt + x0, lc.rs:20-35 order
            if (STRICT) {
#pragma unroll
                for (int q = 0; q < PPT; ++q) out[q] = A::mul(w[0], K[0][q]);
#pragma unroll
                for (int j = 1; j < S; ++j)
#pragma unroll
                    for (int q = 0; q < PPT; ++q) out[q] = A::axpy(out[q], w[j], K[j][q]);
            } else {
#pragma unroll
                for (int q = 0; q < PPT; ++q) out[q] = 0.0;
#pragma unroll
                for (int j = 0; j < S; ++j)
                    if (w[j] != 0.0) {
#pragma unroll
                        for (int q = 0; q < PPT; ++q) out[q] = fma(w[j], K[j][q], out[q]);
                    }
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) out[q] = A::add(A::mul(out[q], sa.dt), xc[q]);
        };
        weigh(sa.b, ox);
        if (sa.use_err) {  // the reference propagates X_berr and keeps x_err = X_b - X_berr (rk.rs:142-147)
            double xbe[PPT];
            weigh(sa.b_err, xbe);
#pragma unroll
            for (int q = 0; q < PPT; ++q) oe[q] = A::sub(ox[q], xbe[q]), ox[q] = xbe[q];
        } else {
#pragma unroll
            for (int q = 0; q < PPT; ++q) oe[q] = 0.0;
        }
        const int p0 = tid * PPT;
        if (p0 >= HS && base + PPT <= own_hi) {  // the whole run is owned (and 16-byte aligned: w0, PPT even)
#pragma unroll
            for (int q = 0; q < PPT; q += 2) {
                *reinterpret_cast<double2*>(next_x + base + q) = make_double2(ox[q], ox[q + 1]);
                if (sa.use_err) *reinterpret_cast<double2*>(x_err + base + q) = make_double2(oe[q], oe[q + 1]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < PPT; ++q)
                if (p0 + q >= HS && base + q < own_hi) {
                    next_x[base + q] = ox[q];
                    if (sa.use_err) x_err[base + q] = oe[q];
                }
        }
        if ((S - 1) % 2 == 0) __syncthreads();  // the last stage read line 0, which the next tile's first stencil rewrites
    }
}
