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
#include "tile_pipe.cuh"
#include "rk_stage_pointwise.cuh"


// ---- HEAT1D, single large state (N == 1) ---------------------------------------------------------------------------
// A pure stream: each lane owns one PAIR of consecutive grid points per iteration (128-bit loads of x0 and of every K_j it
// needs), builds the stage argument of its pair in registers, and gets the two neighbouring values of the stencil from
// the adjacent lanes with warp shuffles. Only the two edge lanes of a warp rebuild their halo value from global memory
// (same operations, same order, so the bits are identical; those loads hit L1/L2). No shared memory, no block barrier.
constexpr int HEAT_THREADS = 256;
constexpr int HEAT_TILE = HEAT_THREADS * 2;  // grid points per CTA per iteration

template <bool STRICT> __device__ __forceinline__ double heat_stage_point(const double* __restrict__ x0, const StageArgs& sa, int64_t e) {
    const double xh = x0[e];
    return sa.i == 0 ? xh : stage_elem<STRICT>(sa, sa.a, sa.i, e, xh, sa.dt);
}

// b / b_err combinations of the tail for one grid point whose K_0..K_{s-2} are in kj[] and K_{s-1} = kl (lc.rs:20-35 order)
template <bool STRICT> __device__ __forceinline__ void heat_tail_point(const StageArgs& sa, const double (&kj)[8], double kl, double xc, double* ox, double* oe) {
    using A = Ar<STRICT>;
    const int s = sa.s;
    double xb, xbe = 0.0;
    if (STRICT) {
        xb = A::mul(sa.b[0], s == 1 ? kl : kj[0]);
#pragma unroll
        for (int j = 1; j < 8; ++j)
            if (j < s) xb = A::axpy(xb, sa.b[j], j == s - 1 ? kl : kj[j]);
    } else {
        xb = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < s && sa.b[j] != 0.0) xb = fma(sa.b[j], j == s - 1 ? kl : kj[j], xb);
    }
    xb = A::add(A::mul(xb, sa.dt), xc);
    if (sa.use_err) {
        if (STRICT) {
            xbe = A::mul(sa.b_err[0], s == 1 ? kl : kj[0]);
#pragma unroll
            for (int j = 1; j < 8; ++j)
                if (j < s) xbe = A::axpy(xbe, sa.b_err[j], j == s - 1 ? kl : kj[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < s && sa.b_err[j] != 0.0) xbe = fma(sa.b_err[j], j == s - 1 ? kl : kj[j], xbe);
        }
        xbe = A::add(A::mul(xbe, sa.dt), xc);
        *ox = xbe, *oe = A::sub(xb, xbe);  // the reference propagates X_berr (rk.rs:142-147)
    } else {
        *ox = xb;
    }
}

template <bool STRICT, bool TAIL>
__global__ void __launch_bounds__(HEAT_THREADS) stage_heat_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ StageArgs sa,
                                                                  double kappa, double* __restrict__ k_out, double* __restrict__ next_x,
                                                                  double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    const int nterm = sa.i;
    const int nload = TAIL ? sa.s - 1 : nterm;  // the tail needs K_0..K_{s-2} for the b-combination
    const int64_t npairs = d >> 1;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // warp-uniform loop: all 32 lanes iterate while the warp's first pair is in range, so the shuffles are always converged
    for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x - lane; p0 < npairs; p0 += stride) {
        const int64_t p = p0 + lane, e0 = 2 * p;
        const bool act = p < npairs;
        double xc[2] = {0.0, 0.0}, kj[2][8], sg[2] = {0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 8; ++j) kj[0][j] = kj[1][j] = 0.0;
        if (act) {
            const double2 xv = *reinterpret_cast<const double2*>(x0 + e0);
            xc[0] = xv.x, xc[1] = xv.y;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < nload && (STRICT || TAIL || sa.a[j] != 0.0)) {
                    const double2 kv = *reinterpret_cast<const double2*>(sa.K[j] + e0);
                    kj[0][j] = kv.x, kj[1][j] = kv.y;
                }
#pragma unroll
            for (int v = 0; v < 2; ++v) {  // stage argument x0 + dt * sum a_j K_j of the own pair (rk.rs:121-124)
                double acc;
                if (nterm == 0) {
                    acc = xc[v];
                } else {
                    if (STRICT) {
                        acc = A::mul(sa.a[0], kj[v][0]);
#pragma unroll
                        for (int j = 1; j < 8; ++j)
                            if (j < nterm) acc = A::axpy(acc, sa.a[j], kj[v][j]);
                    } else {
                        acc = 0.0;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nterm && sa.a[j] != 0.0) acc = fma(sa.a[j], kj[v][j], acc);
                    }
                    acc = A::add(A::mul(acc, sa.dt), xc[v]);
                }
                sg[v] = acc;
            }
        }
        double left = __shfl_up_sync(0xffffffffu, sg[1], 1), right = __shfl_down_sync(0xffffffffu, sg[0], 1);
        if (!act) continue;
        if (lane == 0) left = heat_stage_point<STRICT>(x0, sa, e0 == 0 ? d - 1 : e0 - 1);
        if (lane == 31 || p + 1 >= npairs) right = heat_stage_point<STRICT>(x0, sa, e0 + 2 >= d ? e0 + 2 - d : e0 + 2);
        // kappa * ((u_{j-1} + u_{j+1}) - 2 u_j)
        const double kl0 = A::mul(kappa, A::sub(A::add(left, sg[1]), A::mul(2.0, sg[0])));
        const double kl1 = A::mul(kappa, A::sub(A::add(sg[0], right), A::mul(2.0, sg[1])));
        if (!TAIL) {
            *reinterpret_cast<double2*>(k_out + e0) = make_double2(kl0, kl1);
            continue;
        }
        double ox[2], oe[2] = {0.0, 0.0};
        heat_tail_point<STRICT>(sa, kj[0], kl0, xc[0], &ox[0], &oe[0]);
        heat_tail_point<STRICT>(sa, kj[1], kl1, xc[1], &ox[1], &oe[1]);
        *reinterpret_cast<double2*>(next_x + e0) = make_double2(ox[0], ox[1]);
        if (sa.use_err) *reinterpret_cast<double2*>(x_err + e0) = make_double2(oe[0], oe[1]);
        if (k_out) *reinterpret_cast<double2*>(k_out + e0) = make_double2(kl0, kl1);
    }
    // odd d: the last grid point, one thread
    if ((d & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t e = d - 1;
        const double m = heat_stage_point<STRICT>(x0, sa, e), l = heat_stage_point<STRICT>(x0, sa, e - 1), r = heat_stage_point<STRICT>(x0, sa, 0);
        const double kl = A::mul(kappa, A::sub(A::add(l, r), A::mul(2.0, m)));
        if (!TAIL) {
            k_out[e] = kl;
        } else {
            double kjl[8], ox, oe = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) kjl[j] = j < nload ? sa.K[j][e] : 0.0;
            heat_tail_point<STRICT>(sa, kjl, kl, x0[e], &ox, &oe);
            next_x[e] = ox;
            if (sa.use_err) x_err[e] = oe;
            if (k_out) k_out[e] = kl;
        }
    }
}

// ---- HEAT1D, single large state, TMA-staged ------------------------------------------------------------------------
// Same arithmetic as stage_heat_kernel; the stage buffers reach the SM through shared memory instead of registers:
// thread 0 of each persistent CTA keeps several tiles of every row the stage reads (x0 and the NK stage derivatives with
// a non-skipped coefficient, compacted by the host into HeatArgs), each with a two-point halo on both sides (periodic
// wrap), in flight with cp.async.bulk + mbarrier. Bytes in flight per SM are then ~200 KB whatever the register budget;
// the shuffle kernel above, with one 128-bit load per array per lane in flight, sits at ~3 TB/s. NK is a template
// parameter so that every per-row loop is fully unrolled with no predicates. Needs d even (16-byte aligned halo copies).
constexpr int HT_THREADS = 256;
constexpr int HT_EPT = 4;                      // grid points per thread
constexpr int HT_TILE = HT_THREADS * HT_EPT;   // 1024 points = 8 KiB per row
constexpr int HT_ROW = HT_TILE + 4;            // + 2-point halo each side (keeps every piece 16-byte aligned)
constexpr int HT_STAGES_MAX = 8;

struct HeatArgs {
    const double* K[8];  // the NK stage-derivative buffers this launch reads, in increasing stage order
    double a[8];         // their stage coefficients a_ij (zero-padded terms kept in STRICT mode)
    double b[8], b_err[8];  // tail: weights of those K's; the last stage's own weight is b_last / b_err_last
    double b_last, b_err_last;
    int nterm;           // number of leading rows that enter the stage argument (the tail may load more rows than that)
    int use_err, nst;
    double dt, kappa;
};

template <bool STRICT, bool TAIL, int NK>
__global__ void __launch_bounds__(HT_THREADS, (NK <= 3 ? 2 : 1)) stage_heat_tma_kernel(const double* __restrict__ x0, int64_t d, const __grid_constant__ HeatArgs ha,
                                                                                        double* __restrict__ k_out, double* __restrict__ next_x,
                                                                                        double* __restrict__ x_err) {
    using A = Ar<STRICT>;
    constexpr int NR = NK + 1;
    extern __shared__ __align__(128) double sb[];  // [nst][NR][HT_ROW]
    __shared__ __align__(8) uint64_t full[HT_STAGES_MAX];
    const int nst = ha.nst;
    const int64_t n_tiles = (d + HT_TILE - 1) / HT_TILE, G = gridDim.x, first = blockIdx.x;
    const int64_t my_count = first < n_tiles ? (n_tiles - first + G - 1) / G : 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < HT_STAGES_MAX; ++q) pipe::mbar_init(&full[q], 1);
        pipe::fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int64_t k) {
        const int st = (int)(k % nst);
        const int64_t base = (first + k * G) * HT_TILE;
        const int64_t cnt = min((int64_t)HT_TILE, d - base);
        double* dst = sb + (size_t)st * NR * HT_ROW;
        pipe::mbar_expect_tx(&full[st], (uint32_t)(NR * (cnt + 4) * sizeof(double)));
        const int64_t lh = base == 0 ? d - 2 : base - 2, rh = base + cnt >= d ? 0 : base + cnt;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const double* src = r == 0 ? x0 : ha.K[r - 1];
            double* rd = dst + (size_t)r * HT_ROW;
            pipe::bulk_g2s(rd, src + lh, 16, &full[st]);
            pipe::bulk_g2s(rd + 2, src + base, (uint32_t)(cnt * sizeof(double)), &full[st]);
            pipe::bulk_g2s(rd + 2 + cnt, src + rh, 16, &full[st]);
        }
    };
    if (threadIdx.x == 0)
        for (int64_t k = 0; k < my_count && k < nst; ++k) issue(k);
    const int li = threadIdx.x * HT_EPT;
    for (int64_t k = 0; k < my_count; ++k) {
        const int st = (int)(k % nst);
        const int64_t base = (first + k * G) * HT_TILE;
        const int64_t cnt = min((int64_t)HT_TILE, d - base);
        pipe::mbar_wait(&full[st], (uint32_t)((k / nst) & 1));
        // six consecutive values per row: points li-1 .. li+4 of the tile (the halo-extended row starts at point -2)
        const double* src = sb + (size_t)st * NR * HT_ROW + 2 + li;
        double v[NR][6];
        const bool act = li < cnt;
        if (act) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double* rs = src + (size_t)r * HT_ROW;
                const double2 b0 = *reinterpret_cast<const double2*>(rs), b1 = *reinterpret_cast<const double2*>(rs + 2);
                v[r][0] = rs[-1], v[r][1] = b0.x, v[r][2] = b0.y, v[r][3] = b1.x, v[r][4] = b1.y, v[r][5] = rs[4];
            }
        }
        __syncthreads();  // every lane holds its values: the stage may be refilled
        if (threadIdx.x == 0 && k + nst < my_count) issue(k + nst);
        if (!act) continue;
        double sg[6];  // stage argument x0 + dt * sum a_j K_j at the six points (rk.rs:121-124)
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            double acc = v[0][p];
            if (NK > 0 && ha.nterm > 0) {
                if (STRICT) {
                    acc = A::mul(ha.a[0], v[1][p]);
#pragma unroll
                    for (int r = 1; r < NK; ++r)
                        if (r < ha.nterm) acc = A::axpy(acc, ha.a[r], v[r + 1][p]);
                } else {
                    acc = ha.a[0] * v[1][p];
#pragma unroll
                    for (int r = 1; r < NK; ++r)
                        if (r < ha.nterm) acc = fma(ha.a[r], v[r + 1][p], acc);
                }
                acc = A::add(A::mul(acc, ha.dt), v[0][p]);
            }
            sg[p] = acc;
        }
        double kl[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) kl[p] = A::mul(ha.kappa, A::sub(A::add(sg[p], sg[p + 2]), A::mul(2.0, sg[p + 1])));
        const int64_t e0 = base + li;
        const bool full4 = li + HT_EPT <= cnt;  // cnt is even and li a multiple of 4: otherwise exactly 2 points remain
        if (!TAIL) {
            *reinterpret_cast<double2*>(k_out + e0) = make_double2(kl[0], kl[1]);
            if (full4) *reinterpret_cast<double2*>(k_out + e0 + 2) = make_double2(kl[2], kl[3]);
            continue;
        }
        double ox[4], oe[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int p = 0; p < 4; ++p) {  // sum_j b_j K_j with K_{s-1} = kl in registers, same left-to-right order as lc.rs:20-35
            double xb, xbe = 0.0;
            if (NK == 0) {
                xb = A::mul(ha.b_last, kl[p]);
            } else if (STRICT) {
                xb = A::mul(ha.b[0], v[1][p + 1]);
#pragma unroll
                for (int r = 1; r < NK; ++r) xb = A::axpy(xb, ha.b[r], v[r + 1][p + 1]);
                xb = A::axpy(xb, ha.b_last, kl[p]);
            } else {
                xb = ha.b[0] * v[1][p + 1];
#pragma unroll
                for (int r = 1; r < NK; ++r) xb = fma(ha.b[r], v[r + 1][p + 1], xb);
                xb = fma(ha.b_last, kl[p], xb);
            }
            xb = A::add(A::mul(xb, ha.dt), v[0][p + 1]);
            if (ha.use_err) {
                if (NK == 0) {
                    xbe = A::mul(ha.b_err_last, kl[p]);
                } else if (STRICT) {
                    xbe = A::mul(ha.b_err[0], v[1][p + 1]);
#pragma unroll
                    for (int r = 1; r < NK; ++r) xbe = A::axpy(xbe, ha.b_err[r], v[r + 1][p + 1]);
                    xbe = A::axpy(xbe, ha.b_err_last, kl[p]);
                } else {
                    xbe = ha.b_err[0] * v[1][p + 1];
#pragma unroll
                    for (int r = 1; r < NK; ++r) xbe = fma(ha.b_err[r], v[r + 1][p + 1], xbe);
                    xbe = fma(ha.b_err_last, kl[p], xbe);
                }
                xbe = A::add(A::mul(xbe, ha.dt), v[0][p + 1]);
                ox[p] = xbe, oe[p] = A::sub(xb, xbe);  // the reference propagates X_berr (rk.rs:142-147)
            } else {
                ox[p] = xb;
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
