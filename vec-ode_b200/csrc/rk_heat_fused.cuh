// rk_heat_fused.cuh — one whole explicit Runge-Kutta step of the periodic 1-D heat equation in ONE kernel.
//
// The stage path (rk_stage.cuh) makes one pass over HBM per stage: 13 vector passes = 104 B per grid point for RK4. But
// the right-hand side is a 3-point stencil, so the new value of a grid point depends on x0 at its s nearest neighbours on
// each side only: a CTA that holds a tile of L points can run ALL s stages on it without leaving the SM and still gets the
// inner L - 2s points right. Traffic drops to read x0 once, write next_x once: 16 B per grid point (24 B with x_err).
//
//   * every WARP works alone on its own tile of 256 points (8 consecutive points per lane): x0 and every K_j of a lane's run
//     stay in REGISTERS (the stage argument x0 + dt sum_j a_ij K_j is pointwise, rk.rs:121-124), and so do the stencil
//     neighbours of all but the run's two end points, which come from the adjacent lanes by warp shuffle. No block barrier
//     and no shared-memory exchange: the first versions of this kernel spent their time on exactly those
//     (one __syncthreads() per stage: "barrier" and "short scoreboard" were 60 % of the stall samples);
//   * the points within s of a tile end see wrong neighbours from stage to stage and are simply not stored: warp tiles
//     overlap by 2s points (3 % redundant work for RK4);
//   * x0 tiles arrive through a 3-deep cp.async pipeline per warp (16-byte copies, coalesced on the global side, written to
//     shared memory in an XOR-swizzled order so that each lane then reads its 64-byte run with conflict-free 128-bit
//     loads); ~100 KB per SM are in flight however few registers are free. Only the tiles that touch the two ends of the
//     periodic grid are filled by ordinary wrapped loads.
//
// Per point the operations and their order are those of stage_heat_kernel / heat_tail_point, so in STRICT arithmetic the
// result is bit-identical to the stage path (and to the reference's un-fused code).
#pragma once
#include "rk_stage.cuh"

constexpr int HF_THREADS = 256;
constexpr int HF_NST = 3;     // tiles in flight per warp
constexpr int HF_PPT = 8;     // points per lane
constexpr int HF_WL = 32 * HF_PPT;  // points per warp tile

// physical 16-byte chunk of logical chunk c of a warp tile: the 4 chunks of lane tt's run are permuted by (tt >> 1) & 3, which
// spreads the 8 lanes of a quarter-warp over all 8 bank groups for the 128-bit run loads (and for the copies that fill it)
__device__ __forceinline__ int hf_chunk(int c) { return (c & ~3) | ((c & 3) ^ ((c >> 3) & 3)); }

template <int S, bool STRICT>
__global__ void __launch_bounds__(HF_THREADS, (S <= 4 ? 2 : 1))  // 7 stages x 8 points = 112 registers of K alone
    heat_fused_step_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ TableauDev tb, const __grid_constant__ StageArgs sa,
                           double kappa, double* __restrict__ next_x, double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int PPT = HF_PPT, L = HF_WL;
    constexpr int HS = (S + 1) & ~1;  // halo per side, even so that tile windows start on 16-byte boundaries
    constexpr int T = L - 2 * HS;     // owned points per warp tile
    extern __shared__ __align__(128) double sbuf[];  // [warps][HF_NST][L]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_tiles = (d + T - 1) / T;
    const int64_t W = (int64_t)gridDim.x * (HF_THREADS / 32), first = (int64_t)blockIdx.x * (HF_THREADS / 32) + warp;
    const int64_t my_count = first < n_tiles ? (n_tiles - first + W - 1) / W : 0;
    double* wbuf = sbuf + (size_t)warp * HF_NST * L;
    auto window = [&](int64_t k) { return (first + k * W) * T - HS; };               // global index of local point 0 of the warp's k-th tile
    auto inside = [&](int64_t k) { return window(k) >= 0 && window(k) + L <= d; };  // no periodic wrap inside the tile
    auto issue = [&](int64_t k) {  // every lane: 4 of the tile's 128 chunks; always commits a group so that the group count is uniform
        if (k < my_count && inside(k)) {
            const double* src = x0 + window(k);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(wbuf + (size_t)(k % HF_NST) * L);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = lane + 32 * r;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)hf_chunk(c)), "l"(src + 2 * c) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int64_t k = 0; k < HF_NST - 1; ++k) issue(k);
    for (int64_t k = 0; k < my_count; ++k) {
        const int64_t tile = first + k * W, w0 = window(k);
        const int64_t base = w0 + (int64_t)lane * PPT;  // global index of the first point of this lane's run (before wrapping)
        double* stage = wbuf + (size_t)(k % HF_NST) * L;
        __syncwarp();  // every lane is done with the buffer the next copies go to (read two iterations ago)
        issue(k + HF_NST - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(HF_NST - 1) : "memory");
        if (!inside(k)) {  // a tile at an end of the periodic grid: wrapped loads into the same swizzled order
#pragma unroll
            for (int r = 0; r < PPT; ++r) {
                const int e = lane + 32 * r;
                int64_t g = w0 + e;
                if (g < 0) g += d;
                else if (g >= d) g -= d;
                stage[2 * hf_chunk(e >> 1) + (e & 1)] = x0[g];
            }
        }
        __syncwarp();
        double xc[PPT], K[S][PPT];
#pragma unroll
        for (int q = 0; q < PPT; q += 2) {
            const double2 v = *reinterpret_cast<const double2*>(stage + 2 * hf_chunk(lane * 4 + (q >> 1)));
            xc[q] = v.x, xc[q + 1] = v.y;
        }
        // K = kappa * ((u_{j-1} + u_{j+1}) - 2 u_j) over the lane's run, the two outer neighbours from the adjacent lanes
        auto stencil = [&](const double (&v)[PPT], double (&k_out)[PPT]) {
            const double left = __shfl_up_sync(0xffffffffu, v[PPT - 1], 1), right = __shfl_down_sync(0xffffffffu, v[0], 1);
#pragma unroll
            for (int q = 0; q < PPT; ++q)  // (an FMA form of this, kappa * fma(-2, v, l + r), measured 5 % slower under the power cap)
                k_out[q] = A::mul(kappa, A::sub(A::add(q == 0 ? left : v[q == 0 ? 0 : q - 1], q == PPT - 1 ? right : v[q == PPT - 1 ? q : q + 1]), A::mul(2.0, v[q])));
        };
        stencil(xc, K[0]);  // K_0 = f(x0)
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
                // The stage path skips the zeros of the tableau in FAST arithmetic. Here a skipped term would save one FMA and
                // cost a branch: a zero coefficient adds an exact 0 (for finite K), so the straight chain gives the same bits.
#pragma unroll
                for (int q = 0; q < PPT; ++q) xs[q] = row[0] * K[0][q];
#pragma unroll
                for (int j = 1; j < i; ++j)
#pragma unroll
                    for (int q = 0; q < PPT; ++q) xs[q] = fma(row[j], K[j][q], xs[q]);
            }
#pragma unroll
            for (int q = 0; q < PPT; ++q) xs[q] = A::add(A::mul(xs[q], sa.dt), xc[q]);
            stencil(xs, K[i]);
        }
        // b / b_err combinations (rk.rs:131-151) for the owned points of the tile
        // (heat_tail_point's operations with the stage count known at compile time and the zero tests hoisted out of the run)
        const int64_t own_hi = tile * T + T < d ? tile * T + T : d;
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
                for (int q = 0; q < PPT; ++q) out[q] = w[0] * K[0][q];
#pragma unroll
                for (int j = 1; j < S; ++j)
#pragma unroll
                    for (int q = 0; q < PPT; ++q) out[q] = fma(w[j], K[j][q], out[q]);
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
        const int p0 = lane * PPT;
#pragma unroll
        for (int q = 0; q < PPT; q += 2) {  // HS, PPT and w0 are even: a pair of points is owned or not as a whole, except at an odd end of the grid
            if (p0 + q >= HS && p0 + q < L - HS) {
                if (base + q + 1 < own_hi) {
                    *reinterpret_cast<double2*>(next_x + base + q) = make_double2(ox[q], ox[q + 1]);
                    if (sa.use_err) *reinterpret_cast<double2*>(x_err + base + q) = make_double2(oe[q], oe[q + 1]);
                } else if (base + q < own_hi) {
                    next_x[base + q] = ox[q];
                    if (sa.use_err) x_err[base + q] = oe[q];
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
