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

// ---- the same stage, TMA-staged (large grids) ------------------------------------------------------------------------------------
// The plain kernel above keeps one 8-byte load per array per thread in flight and sits at ~2.6 TB/s, like the compiled-in heat
// kernel's plain version. This one is the structure of stage_heat_tma_kernel (rk_stage.cuh) with the stencil as a functor: thread 0
// of each persistent CTA keeps several 1024-point tiles of every row the stage reads (x0 and the NK stage derivatives with a
// non-skipped coefficient, compacted by the host into StencilArgs), each with an HL-point halo on both sides (HL = R rounded up to
// even, so that every piece stays 16-byte aligned; periodic wrap), in flight with cp.async.bulk + mbarrier; a thread owns four
// consecutive grid points, builds the stage argument at those and at the R points to each side, applies the stencil four times
// and, in the tail launch, forms next_x / x_err with the last stage in registers. Needs d even and R <= 4.
#include "tile_pipe.cuh"

constexpr int SX_THREADS = 256;
constexpr int SX_EPT = 4;
constexpr int SX_TILE = SX_THREADS * SX_EPT;
constexpr int SX_STAGES_MAX = 8;

struct StencilArgs {
    const double* K[8];     // the NK stage-derivative buffers this launch reads, in increasing stage order
    double a[8];            // their stage coefficients (zero-padded terms kept in STRICT mode)
    double b[8], b_err[8];  // tail: weights of those K's; the last stage's own weight is b_last / b_err_last
    double b_last, b_err_last;
    int nterm;              // number of leading rows that enter the stage argument
    int use_err, nst;
    double dt, t_i;
};

template <class ST, bool STRICT, bool TAIL, int NK>
__global__ void __launch_bounds__(SX_THREADS, ((NK <= 3 && ST::R <= 2) ? 2 : 1))
    stage_stencil_tma_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ StencilArgs ha, const __grid_constant__ RhsParams rp,
                             double* __restrict__ k_out, double* __restrict__ next_x, double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int R = ST::R, HL = (R + 1) & ~1, ROW = SX_TILE + 2 * HL, NV = SX_EPT + 2 * HL, NR = NK + 1;
    extern __shared__ __align__(128) double sb[];  // [nst][NR][ROW]
    __shared__ __align__(8) uint64_t full[SX_STAGES_MAX];
    double p[ST::NP];
#pragma unroll
    for (int q = 0; q < ST::NP; ++q) p[q] = rp.shared[q];
    const int nst = ha.nst;
    const int64_t n_tiles = (d + SX_TILE - 1) / SX_TILE, G = gridDim.x, first = blockIdx.x;
    const int64_t my_count = first < n_tiles ? (n_tiles - first + G - 1) / G : 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < SX_STAGES_MAX; ++q) pipe::mbar_init(&full[q], 1);
        pipe::fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int64_t k) {
        const int st = (int)(k % nst);
        const int64_t base = (first + k * G) * SX_TILE;
        const int64_t cnt = min((int64_t)SX_TILE, d - base);
        double* dst = sb + (size_t)st * NR * ROW;
        pipe::mbar_expect_tx(&full[st], (uint32_t)(NR * (cnt + 2 * HL) * sizeof(double)));
        // periodic halos: the left one never straddles the end of the grid (base is 0 or >= HL); the right one does when fewer than HL
        // points remain after the tile (HL = 4 with a two-point last tile), and is then copied in two pieces (all counts are even)
        const int64_t lh = base == 0 ? d - HL : base - HL, after = base + cnt;
        const int64_t r1 = after >= d ? 0 : min((int64_t)HL, d - after);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const double* src = r == 0 ? x0 : ha.K[r - 1];
            double* rd = dst + (size_t)r * ROW;
            pipe::bulk_g2s(rd, src + lh, HL * sizeof(double), &full[st]);
            pipe::bulk_g2s(rd + HL, src + base, (uint32_t)(cnt * sizeof(double)), &full[st]);
            if (r1 > 0) pipe::bulk_g2s(rd + HL + cnt, src + after, (uint32_t)(r1 * sizeof(double)), &full[st]);
            if (r1 < HL) pipe::bulk_g2s(rd + HL + cnt + r1, src, (uint32_t)((HL - r1) * sizeof(double)), &full[st]);
        }
    };
    if (threadIdx.x == 0)
        for (int64_t k = 0; k < my_count && k < nst; ++k) issue(k);
    const int li = threadIdx.x * SX_EPT;
    for (int64_t k = 0; k < my_count; ++k) {
        const int st = (int)(k % nst);
        const int64_t base = (first + k * G) * SX_TILE;
        const int64_t cnt = min((int64_t)SX_TILE, d - base);
        pipe::mbar_wait(&full[st], (uint32_t)((k / nst) & 1));
        // NV consecutive values per row: points li - HL .. li + 3 + HL of the tile (row index 0 is point -HL)
        const double* src = sb + (size_t)st * NR * ROW + li;
        double v[NR][NV];
        const bool act = li < cnt;
        if (act) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double* rs = src + (size_t)r * ROW;
#pragma unroll
                for (int h = 0; h < NV / 2; ++h) {
                    const double2 b2 = *reinterpret_cast<const double2*>(rs + 2 * h);
                    v[r][2 * h] = b2.x, v[r][2 * h + 1] = b2.y;
                }
            }
        }
        __syncthreads();  // every lane holds its values: the stage may be refilled
        if (threadIdx.x == 0 && k + nst < my_count) issue(k + nst);
        if (!act) continue;
        double sg[NV];  // stage argument x0 + dt * sum a_j K_j (rk.rs:121-124) at the points the four stencils reach
#pragma unroll
        for (int q = HL - R; q < HL + SX_EPT + R; ++q) {
            double acc = v[0][q];
            if (NK > 0 && ha.nterm > 0) {
                if (STRICT) {
                    acc = A::mul(ha.a[0], v[NK > 0 ? 1 : 0][q]);
#pragma unroll
                    for (int r = 1; r < NK; ++r)
                        if (r < ha.nterm) acc = A::axpy(acc, ha.a[r], v[r + 1][q]);
                } else {
                    acc = ha.a[0] * v[NK > 0 ? 1 : 0][q];
#pragma unroll
                    for (int r = 1; r < NK; ++r)
                        if (r < ha.nterm) acc = fma(ha.a[r], v[r + 1][q], acc);
                }
                acc = A::add(A::mul(acc, ha.dt), v[0][q]);
            }
            sg[q] = acc;
        }
        const int64_t e0 = base + li;
        double kl[SX_EPT];
#pragma unroll
        for (int pt = 0; pt < SX_EPT; ++pt) {
            double u[2 * R + 1];
#pragma unroll
            for (int c = 0; c < 2 * R + 1; ++c) u[c] = sg[HL - R + pt + c];
            kl[pt] = ST::template eval<STRICT>(ha.t_i, (long long)(e0 + pt), (long long)d, u, p);
        }
        const bool full4 = li + SX_EPT <= cnt;  // cnt is even and li a multiple of 4: otherwise exactly 2 points remain
        if (!TAIL) {
            *reinterpret_cast<double2*>(k_out + e0) = make_double2(kl[0], kl[1]);
            if (full4) *reinterpret_cast<double2*>(k_out + e0 + 2) = make_double2(kl[2], kl[3]);
            continue;
        }
        double ox[SX_EPT], oe[SX_EPT] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int pt = 0; pt < SX_EPT; ++pt) {  // sum_j b_j K_j with K_{s-1} = kl in registers, left to right (lc.rs:20-35)
            const int q = HL + pt;
            double xb, xbe = 0.0;
            if (NK == 0) {
                xb = A::mul(ha.b_last, kl[pt]);
            } else if (STRICT) {
                xb = A::mul(ha.b[0], v[NK > 0 ? 1 : 0][q]);
#pragma unroll
                for (int r = 1; r < NK; ++r) xb = A::axpy(xb, ha.b[r], v[r + 1][q]);
                xb = A::axpy(xb, ha.b_last, kl[pt]);
            } else {
                xb = ha.b[0] * v[NK > 0 ? 1 : 0][q];
#pragma unroll
                for (int r = 1; r < NK; ++r) xb = fma(ha.b[r], v[r + 1][q], xb);
                xb = fma(ha.b_last, kl[pt], xb);
            }
            xb = A::add(A::mul(xb, ha.dt), v[0][q]);
            if (ha.use_err) {
                if (NK == 0) {
                    xbe = A::mul(ha.b_err_last, kl[pt]);
                } else if (STRICT) {
                    xbe = A::mul(ha.b_err[0], v[NK > 0 ? 1 : 0][q]);
#pragma unroll
                    for (int r = 1; r < NK; ++r) xbe = A::axpy(xbe, ha.b_err[r], v[r + 1][q]);
                    xbe = A::axpy(xbe, ha.b_err_last, kl[pt]);
                } else {
                    xbe = ha.b_err[0] * v[NK > 0 ? 1 : 0][q];
#pragma unroll
                    for (int r = 1; r < NK; ++r) xbe = fma(ha.b_err[r], v[r + 1][q], xbe);
                    xbe = fma(ha.b_err_last, kl[pt], xbe);
                }
                xbe = A::add(A::mul(xbe, ha.dt), v[0][q]);
                ox[pt] = xbe, oe[pt] = A::sub(xb, xbe);  // the reference propagates X_berr (rk.rs:142-147)
            } else {
                ox[pt] = xb;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 1 && !full4) break;
            *reinterpret_cast<double2*>(next_x + e0 + 2 * h) = make_double2(ox[2 * h], ox[2 * h + 1]);
            if (ha.use_err) *reinterpret_cast<double2*>(x_err + e0 + 2 * h) = make_double2(oe[2 * h], oe[2 * h + 1]);
            if (k_out) *reinterpret_cast<double2*>(k_out + e0 + 2 * h) = make_double2(kl[2 * h], kl[2 * h + 1]);
        }
    }
}
