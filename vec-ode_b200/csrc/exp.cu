// exp.cu — exponential integrators (src/exp) for ensembles of dense complex systems on a shared operator basis.
//
// What the reference fixes is the SCHEME: midpoint (exp/magnus.rs:10-26), 4th-order Magnus with one commutator
// (exp/magnus.rs:28-83) and the commutator-free CFM4 of cfm_general (exp/cfm.rs:43-100 with the tables of
// dat/mod.rs:4, 67-74), each wrapped in the ODEData / ODEAdaptiveData state machine (base/ode.rs). What it leaves to the
// user is the operator algebra behind ExponentialSplit / Commutator (exp/mod.rs:11-54): lin_zero, exp, map_exp,
// commutator, norm. This file supplies that algebra for L_i = sum_m coef[i][m] * B_m with M complex n x n matrices B_m
// shared by the ensemble:
//   * an operator `L` is its M complex coefficients, so LinearCombination on L is arithmetic on coefficients;
//   * exp(L) is lazy (U = L) and map_exp(U, x) applies the scaled Taylor series of exp(L) to x without forming U
//     (each U is used exactly once per step in all three schemes);
//   * commutator(La, Lb) is expanded on the basis through a structure tensor supplied by the caller.
// Per Taylor term every system of a tile needs B_m x for all m: that is a GEMM  [n x n] x [n x TB]  with the basis as the
// shared operand. It runs on the FP64 tensor cores (mma.sync m8n8k4 DMMA; tcgen05 has no f64 kind): the basis lives in
// shared memory in A-fragment order for the whole kernel, the Taylor term of the tile is re-published to shared memory
// once per term in B-fragment-friendly layout, accumulators and the running sum stay in registers.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "rk_small.cuh"  // CtlArrays / EvSlot / status-word layout shared with the RK solver

#define VO_EXP_MAX_M 4
#define VO_EXP_MAX_E 3  // exponentials per step: CFM4 adaptive = 2 + 1, Magnus adaptive = 1 + 1

struct vo_split_s {
    vo_ctx ctx = nullptr;
    int n = 0, M = 0;
    double* frag_dev = nullptr;  // [M][2 planes][n/8 row blocks][n/4 k-steps][32 lanes]
    double norm1[VO_EXP_MAX_M];  // induced 1-norm of each basis matrix
    double cs[VO_EXP_MAX_M * VO_EXP_MAX_M * VO_EXP_MAX_M];
    bool has_cs = false;
    int taylor_deg = 0;
};

struct vo_expsolver_s {
    vo_ctx ctx = nullptr;
    vo_split sp = nullptr;
    int scheme = 0, M_gen = 0;
    int64_t N = 0;
    double t0 = 0, tf = 0, h_init = 0;
    double2* psi = nullptr;   // [N][n]
    double2* psi0 = nullptr;  // copy of the initial state (reset)
    double* gp = nullptr;     // [N][M_gen-1][3]
    CtlArrays ca{};
    EvSlot* ev_dev = nullptr;
    EvSlot* ev_host = nullptr;
    EvSlot ev_seen{};
    int64_t n_done = 0;
    bool want_err = true;  // alph_err / x_err present (exp/cfm.rs:157-161)
    unsigned split_mask = 1u;  // VO_EXP_SPLIT_MIDPOINT: which basis matrices form split A (default: B_0)
    double atol = 1.0e-6, rtol = 1.0e-4, alpha = 0.9, pw = 1.0 / 3.0, min_dt = 1.0e-6, max_dt = 1.0;
};

namespace {

struct ExpKP {
    int n, M, M_gen, scheme, adaptive, want_err, taylor_deg, mode;  // mode 0: solver event, 1: bare map_exp
    int pw_is_third, count_events;
    int nseq;             // mode 1: exponentials applied one after the other, coefficient sets [nseq][N][M]
    unsigned split_mask;  // VO_EXP_SPLIT_MIDPOINT: bit m set <=> basis matrix m belongs to split A
    double t_end, t_start;
    double rtol, alpha, pw, min_dt, max_dt;
    double norm1[VO_EXP_MAX_M];
    double cs[VO_EXP_MAX_M * VO_EXP_MAX_M * VO_EXP_MAX_M];
    int64_t N;
};

// dat/mod.rs:4, 67-74 (same literals as the reference)
__constant__ double C_GL4[2] = {0.21132486540518711775, 0.78867513459481288225};
__constant__ double CFM_R4[4] = {0.53867513459481288225, -0.038675134594812882255, -0.038675134594812882255, 0.53867513459481288225};
__constant__ double CFM_R2[2] = {0.5, 0.5};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Taylor plan of map_exp (same rule as the CPU restatement): sub-steps so that theta/sq <= 1, then the smallest degree
// whose term falls below 2^-53.
__device__ __forceinline__ void taylor_plan(double theta, int forced_deg, int* sq, int* deg) {
    int s = theta > 1.0 ? (int)ceil(theta) : 1;
    const double th = theta / s;
    double term = 1.0;
    int k = 0;
    while (k < 60) {
        ++k;
        term = term * th / k;
        if (term <= 1.1102230246251565e-16) break;
    }
    *sq = s, *deg = forced_deg > 0 ? forced_deg : k;
}

template <int NDIM, int M, int TB> struct Geo {
    static constexpr int NW = NDIM / 8;      // row blocks of 8 (one warp each per column group)
    static constexpr int NK = NDIM / 4;      // k-steps of 4
    static constexpr int NCG = TB / 16;      // column groups of 16 systems
    static constexpr int NT = 2;             // n-tiles of 8 columns per warp
    static constexpr int LDT = NDIM + 4;     // padded row of the term buffer: conflict-free B-fragment loads
    static constexpr int THREADS = NW * NCG * 32;
    static constexpr int NBUF = (M * 2 * NDIM * NDIM * 8 + 2 * 2 * TB * LDT * 8 > 200 * 1024) ? 1 : 2;
    static constexpr size_t SMEM_B = (size_t)M * 2 * NDIM * NDIM * sizeof(double);
    static constexpr size_t SMEM_T = (size_t)NBUF * 2 * TB * LDT * sizeof(double);
    static constexpr size_t SMEM_COEF = (size_t)VO_EXP_MAX_E * M * TB * sizeof(double2);
    static constexpr size_t SMEM_MISC = (size_t)(NW * TB + 8 * TB) * sizeof(double) + 64 * sizeof(int);
    static constexpr size_t SMEM = SMEM_B + SMEM_T + SMEM_COEF + SMEM_MISC;
};

// x <- exp(sum_m coef[m][s] B_m) x for the tile, state in C-fragment layout:
// lane l of warp (w, cg) owns row 8w + l/4 and columns 16cg + 8j + 2(l%4) + q, j,q in {0,1}.
template <int NDIM, int M, int TB>
__device__ __forceinline__ void map_exp_tile(const double* __restrict__ sB, double* __restrict__ sT, const double2* __restrict__ sCoefE, int sq, int deg,
                                             double (&xr)[2][2], double (&xi)[2][2], int& buf) {
    using G = Geo<NDIM, M, TB>;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int w = wi % G::NW, cg = wi / G::NW;
    const int row = 8 * w + (lane >> 2);
    const double inv_sq = 1.0 / sq;
    double cr[M][2][2], ci[M][2][2];  // coefficients / sq of this lane's four systems
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double2 c = sCoefE[m * TB + 16 * cg + 8 * j + 2 * (lane & 3) + q];
                cr[m][j][q] = c.x * inv_sq, ci[m][j][q] = c.y * inv_sq;
            }
    for (int rep = 0; rep < sq; ++rep) {
        double ar[2][2], ai[2][2], tr[2][2], ti[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) ar[j][q] = tr[j][q] = xr[j][q], ai[j][q] = ti[j][q] = xi[j][q];
        for (int k = 1; k <= deg; ++k) {
            // publish the current term: planar [plane][column][LDT], row fastest
            if (G::NBUF == 1) __syncthreads();
            double* Tr = sT + (size_t)buf * 2 * TB * G::LDT;
            double* Ti = Tr + TB * G::LDT;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col = 16 * cg + 8 * j + 2 * (lane & 3) + q;
                    Tr[col * G::LDT + row] = tr[j][q], Ti[col * G::LDT + row] = ti[j][q];
                }
            __syncthreads();
            double Wr[M][2][2], Wi[M][2][2];
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int j = 0; j < 2; ++j) Wr[m][j][0] = Wr[m][j][1] = Wi[m][j][0] = Wi[m][j][1] = 0.0;
            const double* bA = sB + ((size_t)w * G::NK) * 32 + lane;  // + ((m*2 + plane) * NW) * NK * 32 + kk * 32
            const double* bX = Tr + (16 * cg + (lane >> 2)) * G::LDT + (lane & 3);
#pragma unroll 4
            for (int kk = 0; kk < G::NK; ++kk) {
                double fr[2], fi[2], nfi[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    fr[j] = bX[8 * j * G::LDT + 4 * kk];
                    fi[j] = bX[TB * G::LDT + 8 * j * G::LDT + 4 * kk];
                    nfi[j] = -fi[j];
                }
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const double a_re = bA[((size_t)(m * 2 + 0) * G::NW) * G::NK * 32 + kk * 32];
                    const double a_im = bA[((size_t)(m * 2 + 1) * G::NW) * G::NK * 32 + kk * 32];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma(Wr[m][j][0], Wr[m][j][1], a_re, fr[j]);   // Re += Br Xr
                        dmma(Wr[m][j][0], Wr[m][j][1], a_im, nfi[j]);  // Re -= Bi Xi
                        dmma(Wi[m][j][0], Wi[m][j][1], a_im, fr[j]);   // Im += Bi Xr
                        dmma(Wi[m][j][0], Wi[m][j][1], a_re, fi[j]);   // Im += Br Xi
                    }
                }
            }
            const double ik = 1.0 / k;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    double sr = 0.0, si = 0.0;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        sr += cr[m][j][q] * Wr[m][j][q] - ci[m][j][q] * Wi[m][j][q];
                        si += cr[m][j][q] * Wi[m][j][q] + ci[m][j][q] * Wr[m][j][q];
                    }
                    tr[j][q] = sr * ik, ti[j][q] = si * ik;
                    ar[j][q] += tr[j][q], ai[j][q] += ti[j][q];
                }
            if (G::NBUF == 2) buf ^= 1;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) xr[j][q] = ar[j][q], xi[j][q] = ai[j][q];
    }
}

// generator family: L(t) = B_0 + sum_{m=1}^{M_gen-1} amp_m cos(omega_m t + phase_m) B_m ; coefficients beyond M_gen are 0
template <int M> __device__ __forceinline__ void gen_coef(const double* __restrict__ gp, int M_gen, double t, double (&c)[M]) {
#pragma unroll
    for (int m = 0; m < M; ++m) c[m] = 0.0;
    c[0] = 1.0;
#pragma unroll
    for (int m = 1; m < M; ++m)
        if (m < M_gen) c[m] = gp[(m - 1) * 3 + 0] * cos(gp[(m - 1) * 3 + 1] * t + gp[(m - 1) * 3 + 2]);
}

template <int NDIM, int M, int TB>
__global__ void __launch_bounds__(Geo<NDIM, M, TB>::THREADS, 1)
exp_step_kernel(const __grid_constant__ ExpKP kp, const double* __restrict__ frag, double2* __restrict__ psi, double2* __restrict__ psi_out,
                const double* __restrict__ gp, const double2* __restrict__ coef_in, const CtlArrays ca, EvSlot* __restrict__ ev) {
    using G = Geo<NDIM, M, TB>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sB = reinterpret_cast<double*>(smem_raw);
    double* sT = reinterpret_cast<double*>(smem_raw + G::SMEM_B);
    double2* sCoef = reinterpret_cast<double2*>(smem_raw + G::SMEM_B + G::SMEM_T);               // [E][M][TB]
    double* sNorm = reinterpret_cast<double*>(smem_raw + G::SMEM_B + G::SMEM_T + G::SMEM_COEF);  // [NW][TB]
    double* sTheta = sNorm + G::NW * TB;                                                         // [E][TB]
    double* sDt = sTheta + VO_EXP_MAX_E * TB;                                                    // [TB]
    int* sEv = reinterpret_cast<int*>(sDt + TB + 4 * TB);                                        // [TB] event, then [8] plan, [1] any
    int* sPlan = sEv + TB;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int w = wi % G::NW, cg = wi / G::NW;
    const int row = 8 * w + (lane >> 2);

    // the basis, already in A-fragment order, stays in shared memory for the life of the CTA
    for (size_t i = threadIdx.x; i < G::SMEM_B / sizeof(double2); i += blockDim.x)
        reinterpret_cast<double2*>(sB)[i] = reinterpret_cast<const double2*>(frag)[i];
    __syncthreads();

    const int64_t n_tiles = (kp.N + TB - 1) / TB;
    int buf = 0;
    unsigned c_step = 0, c_chkpt = 0, c_rej = 0, c_end = 0, c_stuck = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TB;
        const bool embedded = kp.mode == 0 && kp.want_err && (kp.scheme == VO_EXP_CFM4 || kp.scheme == VO_EXP_MAGNUS42);
        const int nbase = kp.mode == 1 ? 1 : (kp.scheme == VO_EXP_CFM4 ? 2 : (kp.scheme == VO_EXP_SPLIT_MIDPOINT ? 3 : 1));
        const int nexp = nbase + (embedded ? 1 : 0);
        // ---- phase A: per-system control and exponent coefficients (one thread per system)
        if (threadIdx.x < TB) {
            const int s = threadIdx.x;
            const int64_t sys = base + s;
            int evk = 255;  // not live
            double dt = 0.0;
            double2 ce[VO_EXP_MAX_E][M];
#pragma unroll
            for (int e = 0; e < VO_EXP_MAX_E; ++e)
#pragma unroll
                for (int m = 0; m < M; ++m) ce[e][m] = make_double2(0.0, 0.0);
            if (sys < kp.N) {
                if (kp.mode == 1) {
                    evk = VO_EV_STEP;
#pragma unroll
                    for (int m = 0; m < M; ++m) ce[0][m] = coef_in[sys * M + m];
                } else {
                    const uint32_t word = ca.word[sys];
                    if (!((word >> VO_WORD_STATUS_SHIFT) & VO_TRAJ_DONE)) {
                        const int tgt = (int)(word & VO_WORD_TGT_MASK);
                        const double t = ca.t[sys], h = ca.h[sys];
                        // step_size_of (ode.rs:165-176) with t_list = [t0, tf]
                        if (tgt >= 2) {
                            evk = VO_EV_END;
                        } else {
                            const double rem = (tgt == 0 ? kp.t_start : kp.t_end) - t;
                            if (fabs(rem) <= 2.220446049250313e-16) evk = tgt >= 1 ? VO_EV_END : VO_EV_CHKPT;
                            else dt = rem < h ? rem : h, evk = VO_EV_STEP;
                        }
                        if (evk == VO_EV_STEP) {
                            const double* g = gp + sys * (kp.M_gen - 1) * 3;
                            if (kp.scheme == VO_EXP_MIDPOINT) {  // exp/magnus.rs:10-26
                                double l[M];
                                gen_coef<M>(g, kp.M_gen, t + dt * 0.5, l);
#pragma unroll
                                for (int m = 0; m < M; ++m) ce[0][m].x = l[m] * dt;
                            } else if (kp.scheme == VO_EXP_CFM4) {  // exp/cfm.rs:43-100, cfm_exp :20-40
                                double v0[M], v1[M];
                                gen_coef<M>(g, kp.M_gen, t + C_GL4[0] * dt, v0);
                                gen_coef<M>(g, kp.M_gen, t + C_GL4[1] * dt, v1);
#pragma unroll
                                for (int m = 0; m < M; ++m) {
                                    ce[0][m].x = (CFM_R4[0] * v0[m] + (CFM_R4[1] * v1[m])) * dt;
                                    ce[1][m].x = (CFM_R4[2] * v0[m] + (CFM_R4[3] * v1[m])) * dt;
                                    ce[2][m].x = (CFM_R2[0] * v0[m] + (CFM_R2[1] * v1[m])) * dt;  // error scheme, :83-97
                                }
                            } else if (kp.scheme == VO_EXP_SPLIT_MIDPOINT) {  // split_exp_midpoint, exp/split_exp.rs:520-562
                                // literal: the generator is sampled at t (not t + dt/2) and BOTH splits are scaled by dt/2
                                // (KA[0] and KB[0] by dt0, :540-548), applied as A, B, A (:556-559)
                                double l[M];
                                gen_coef<M>(g, kp.M_gen, t, l);
                                const double dt0 = dt * 0.5;
#pragma unroll
                                for (int m = 0; m < M; ++m) {
                                    const bool in_a = (kp.split_mask >> m) & 1u;
                                    ce[0][m].x = in_a ? l[m] * dt0 : 0.0;
                                    ce[1][m].x = in_a ? 0.0 : l[m] * dt0;
                                    ce[2][m].x = ce[0][m].x;
                                }
                            } else {  // magnus_42, exp/magnus.rs:28-83
                                const double c_mid = 0.288675134594812882254574390251;
                                const double b1 = dt * 0.5, b2 = dt * dt * -0.144337567297406441127287195125;
                                const double mid_t = t + b1;
                                double l0[M], l1[M], w2[M];
                                gen_coef<M>(g, kp.M_gen, mid_t - c_mid * dt, l0);
                                gen_coef<M>(g, kp.M_gen, mid_t + c_mid * dt, l1);
#pragma unroll
                                for (int c = 0; c < M; ++c) w2[c] = 0.0;
#pragma unroll
                                for (int a = 0; a < M; ++a)
#pragma unroll
                                    for (int b = 0; b < M; ++b) {
                                        const double ab = l0[a] * l1[b];
#pragma unroll
                                        for (int c = 0; c < M; ++c) {
                                            const double sc = kp.cs[(a * M + b) * M + c];
                                            if (sc != 0.0) w2[c] = w2[c] + ab * sc;
                                        }
                                    }
#pragma unroll
                                for (int m = 0; m < M; ++m) {
                                    const double w1 = (l0[m] + l1[m]) * b1;
                                    ce[0][m].x = w1 + w2[m] * b2;  // u = exp(w1 + w2)
                                    ce[1][m].x = w1;               // u1 = exp(w1), the 2nd-order embedded solution
                                }
                            }
                        }
                    }
                }
            }
            sEv[s] = evk, sDt[s] = dt;
#pragma unroll
            for (int e = 0; e < VO_EXP_MAX_E; ++e) {
                double th = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    sCoef[(e * M + m) * TB + s] = ce[e][m];
                    th += hypot(ce[e][m].x, ce[e][m].y) * kp.norm1[m];
                }
                sTheta[e * TB + s] = evk == VO_EV_STEP ? th : 0.0;
            }
        }
        __syncthreads();
        // ---- phase B: tile-uniform Taylor plan per exponential
        if (threadIdx.x == 0) {
            int any = 0;
            for (int s = 0; s < TB; ++s) any |= sEv[s] == VO_EV_STEP;
            sPlan[2 * VO_EXP_MAX_E] = any;
            for (int e = 0; e < VO_EXP_MAX_E; ++e) {
                double th = 0.0;
                for (int s = 0; s < TB; ++s) th = fmax(th, sTheta[e * TB + s]);
                taylor_plan(th, kp.taylor_deg, &sPlan[2 * e], &sPlan[2 * e + 1]);
            }
        }
        __syncthreads();
        const bool any_step = sPlan[2 * VO_EXP_MAX_E] != 0;
        // ---- phase C: the exponentials
        double x0r[2][2], x0i[2][2], xfr[2][2], xfi[2][2], xer[2][2], xei[2][2];
        if (any_step) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int64_t sys = base + 16 * cg + 8 * j + 2 * (lane & 3) + q;
                    const double2 v = sys < kp.N ? psi[sys * NDIM + row] : make_double2(0.0, 0.0);
                    x0r[j][q] = xfr[j][q] = v.x, x0i[j][q] = xfi[j][q] = v.y;
                    xer[j][q] = xei[j][q] = 0.0;
                }
            map_exp_tile<NDIM, M, TB>(sB, sT, sCoef, sPlan[0], sPlan[1], xfr, xfi, buf);
            if (kp.mode == 0 && nbase >= 2) map_exp_tile<NDIM, M, TB>(sB, sT, sCoef + M * TB, sPlan[2], sPlan[3], xfr, xfi, buf);
            if (kp.mode == 0 && nbase >= 3) map_exp_tile<NDIM, M, TB>(sB, sT, sCoef + 2 * M * TB, sPlan[4], sPlan[5], xfr, xfi, buf);
            for (int q = 1; kp.mode == 1 && q < kp.nseq; ++q) {  // vo_map_exp_seq: the next exponential of the composition
                __syncthreads();
                if (threadIdx.x < TB) {
                    const int64_t sys = base + threadIdx.x;
                    double th = 0.0;
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const double2 c = sys < kp.N ? coef_in[((int64_t)q * kp.N + sys) * M + m] : make_double2(0.0, 0.0);
                        sCoef[m * TB + threadIdx.x] = c;
                        th += hypot(c.x, c.y) * kp.norm1[m];
                    }
                    sTheta[threadIdx.x] = th;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double th = 0.0;
                    for (int s = 0; s < TB; ++s) th = fmax(th, sTheta[s]);
                    taylor_plan(th, kp.taylor_deg, &sPlan[0], &sPlan[1]);
                }
                __syncthreads();
                map_exp_tile<NDIM, M, TB>(sB, sT, sCoef, sPlan[0], sPlan[1], xfr, xfi, buf);
            }
            if (embedded) {  // embedded lower-order solution from x0, then x_err = that - xf
                const int e = nexp - 1;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) xer[j][q] = x0r[j][q], xei[j][q] = x0i[j][q];
                map_exp_tile<NDIM, M, TB>(sB, sT, sCoef + e * M * TB, sPlan[2 * e], sPlan[2 * e + 1], xer, xei, buf);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) xer[j][q] -= xfr[j][q], xei[j][q] -= xfi[j][q];
            }
        }
        // ---- phase D: error norm (2-norm of x_err per system), controller, apply_step
        if (kp.mode == 0) {
            if (any_step && kp.adaptive) {
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        double v = xer[j][q] * xer[j][q] + xei[j][q] * xei[j][q];
                        v += __shfl_xor_sync(0xffffffffu, v, 4), v += __shfl_xor_sync(0xffffffffu, v, 8), v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if ((lane >> 2) == 0) sNorm[w * TB + 16 * cg + 8 * j + 2 * (lane & 3) + q] = v;
                    }
            }
            __syncthreads();
            if (threadIdx.x < TB) {
                const int s = threadIdx.x;
                const int64_t sys = base + s;
                int evk = sEv[s];
                if (evk != 255) {
                    const uint32_t word = ca.word[sys];
                    int tgt = (int)(word & VO_WORD_TGT_MASK);
                    uint32_t status = word >> VO_WORD_STATUS_SHIFT;
                    if (evk == VO_EV_STEP) {
                        const double h = ca.h[sys];
                        if (kp.adaptive) {  // handle_step_adaptive, ode.rs:311-334
                            double nn = 0.0;
                            for (int ww = 0; ww < G::NW; ++ww) nn += sNorm[ww * TB + s];
                            const double dxn = sqrt(nn);
                            const double f = kp.rtol / dxn;
                            const double mul = kp.alpha * pow(f, kp.pw);  // step_size_mul, ode.rs:133-135
                            const double fp_lim = fmin(fmax(mul, 0.3), 2.0);
                            const double new_h = fmin(fmax(fp_lim * h, kp.min_dt), kp.max_dt);
                            if (!(dxn == dxn)) status |= VO_TRAJ_NONFINITE;
                            if (f <= 1.0) {
                                evk = VO_EV_REJECT;
                                if (h <= kp.min_dt) status |= VO_TRAJ_STUCK, ++c_stuck;
                            }
                            ca.prev_h[sys] = h, ca.h[sys] = new_h, ca.dx_norm[sys] = dxn;
                        }
                        if (evk == VO_EV_STEP) ca.t[sys] += sDt[s], ca.n_accept[sys] += 1, ++c_step;
                        else ca.n_reject[sys] += 1, ++c_rej;
                    } else {  // Chkpt / End: checkpoint_update, ode.rs:192-195
                        tgt += 1, ca.h[sys] = ca.prev_h[sys];
                        if (evk == VO_EV_END) status |= VO_TRAJ_DONE, ++c_end;
                        else ++c_chkpt;
                    }
                    const uint32_t nw = ((uint32_t)tgt & VO_WORD_TGT_MASK) | (status << VO_WORD_STATUS_SHIFT);
                    if (nw != word) ca.word[sys] = nw;
                    sEv[s] = evk;
                }
            }
            __syncthreads();
        }
        // ---- phase E: masked commit (accepted systems only)
        if (any_step) {
            double2* dst = kp.mode == 1 ? psi_out : psi;
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col = 16 * cg + 8 * j + 2 * (lane & 3) + q;
                    const int64_t sys = base + col;
                    if (sys < kp.N && sEv[col] == VO_EV_STEP) dst[sys * NDIM + row] = make_double2(xfr[j][q], xfi[j][q]);
                }
        }
        __syncthreads();  // sEv / sCoef are rewritten by the next tile
    }
    if (kp.mode == 0 && kp.count_events && threadIdx.x < 32) {
        c_step = __reduce_add_sync(0xffffffffu, c_step), c_chkpt = __reduce_add_sync(0xffffffffu, c_chkpt);
        c_rej = __reduce_add_sync(0xffffffffu, c_rej), c_end = __reduce_add_sync(0xffffffffu, c_end);
        c_stuck = __reduce_add_sync(0xffffffffu, c_stuck);
        if (threadIdx.x == 0) {
            EvSlot* slot = ev + (blockIdx.x % VO_EV_SLOTS);
            if (c_step) atomicAdd(&slot->n_step, (unsigned long long)c_step);
            if (c_chkpt) atomicAdd(&slot->n_chkpt, (unsigned long long)c_chkpt);
            if (c_rej) atomicAdd(&slot->n_reject, (unsigned long long)c_rej);
            if (c_end) atomicAdd(&slot->n_end, (unsigned long long)c_end);
            if (c_stuck) atomicAdd(&slot->n_stuck, (unsigned long long)c_stuck);
        }
    }
}

// NormedExponentialSplit::norm (exp/mod.rs:37-45): 2-norm of every state, one warp per system
__global__ void exp_norm_kernel(const double2* __restrict__ psi, int n, int64_t N, double* __restrict__ out) {
    const int64_t sys = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sys >= N) return;
    double acc = 0.0;
    for (int r = threadIdx.x & 31; r < n; r += 32) {
        const double2 v = psi[sys * n + r];
        acc += v.x * v.x + v.y * v.y;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) out[sys] = sqrt(acc);
}

__global__ void exp_ctl_fill_kernel(CtlArrays ca, int64_t N, double t, double h) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    ca.t[i] = t, ca.h[i] = h, ca.prev_h[i] = h, ca.dx_norm[i] = 0.0;
    ca.n_accept[i] = 0, ca.n_reject[i] = 0, ca.word[i] = 0;
}

template <int NDIM, int M, int TB>
int32_t launch_exp(vo_ctx c, const ExpKP& kp, const double* frag, double2* psi, double2* psi_out, const double* gp, const double2* coef_in,
                   const CtlArrays& ca, EvSlot* ev) {
    using G = Geo<NDIM, M, TB>;
    auto k = exp_step_kernel<NDIM, M, TB>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM) != cudaSuccess)
            return vo_fail(c, VO_ERR_CUDA, "exp: shared-memory carve-out rejected");
        configured = true;
    }
    const int64_t tiles = ceil_div(kp.N, TB);
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, c->sm_count);
    k<<<grid, G::THREADS, G::SMEM, c->stream>>>(kp, frag, psi, psi_out, gp, coef_in, ca, ev);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

int32_t dispatch_exp(vo_split sp, const ExpKP& kp, double2* psi, double2* psi_out, const double* gp, const double2* coef_in, const CtlArrays& ca, EvSlot* ev) {
    vo_ctx c = sp->ctx;
#define VO_EXP_CASE(NDIM, MM, TB) \
    if (sp->n == NDIM && sp->M == MM) return launch_exp<NDIM, MM, TB>(c, kp, sp->frag_dev, psi, psi_out, gp, coef_in, ca, ev);
    VO_EXP_CASE(64, 2, 16) VO_EXP_CASE(64, 3, 16) VO_EXP_CASE(64, 1, 16) VO_EXP_CASE(32, 2, 16) VO_EXP_CASE(32, 3, 16) VO_EXP_CASE(16, 2, 16) VO_EXP_CASE(16, 3, 16)
#undef VO_EXP_CASE
    return vo_fail(c, VO_ERR_UNSUPPORTED, "exp: supported shapes are n in {16, 32, 64} with M in {2, 3} (and n = 64, M = 1)");
}

ExpKP make_kp(const vo_split_s* sp, int mode) {
    ExpKP kp;
    std::memset(&kp, 0, sizeof kp);
    kp.n = sp->n, kp.M = sp->M, kp.mode = mode, kp.taylor_deg = sp->taylor_deg;
    std::memcpy(kp.norm1, sp->norm1, sizeof kp.norm1);
    for (int a = 0; a < sp->M; ++a)
        for (int b = 0; b < sp->M; ++b)
            for (int cc = 0; cc < sp->M; ++cc) kp.cs[(a * sp->M + b) * sp->M + cc] = sp->has_cs ? sp->cs[(a * sp->M + b) * sp->M + cc] : 0.0;
    return kp;
}

int32_t exp_ev_read(vo_expsolver_s* s, EvSlot* out) {
    vo_ctx c = s->ctx;
    VO_CUDA(c, cudaMemcpyAsync(s->ev_host, s->ev_dev, sizeof(EvSlot) * VO_EV_SLOTS, cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    EvSlot tot;
    std::memset(&tot, 0, sizeof tot);
    for (int i = 0; i < VO_EV_SLOTS; ++i)
        tot.n_step += s->ev_host[i].n_step, tot.n_chkpt += s->ev_host[i].n_chkpt, tot.n_reject += s->ev_host[i].n_reject, tot.n_end += s->ev_host[i].n_end,
            tot.n_stuck += s->ev_host[i].n_stuck;
    out->n_step = tot.n_step - s->ev_seen.n_step, out->n_chkpt = tot.n_chkpt - s->ev_seen.n_chkpt, out->n_reject = tot.n_reject - s->ev_seen.n_reject;
    out->n_end = tot.n_end - s->ev_seen.n_end, out->n_stuck = tot.n_stuck - s->ev_seen.n_stuck;
    s->ev_seen = tot;
    return VO_OK;
}

int32_t exp_launch_event(vo_expsolver_s* s, bool adaptive) {
    vo_ctx c = s->ctx;
    if (adaptive && !s->want_err) return vo_fail(c, VO_ERR_NOT_ADAPTIVE, "adaptive step validation failed");  // ode.rs:312
    if (adaptive && (s->scheme == VO_EXP_MIDPOINT || s->scheme == VO_EXP_SPLIT_MIDPOINT))
        return vo_fail(c, VO_ERR_NOT_ADAPTIVE, "this solver has no error estimate (it only implements ODESolver)");
    ExpKP kp = make_kp(s->sp, 0);
    kp.M_gen = s->M_gen, kp.scheme = s->scheme, kp.adaptive = adaptive ? 1 : 0;
    kp.want_err = (s->want_err && adaptive) ? 1 : 0;  // the embedded solution only feeds the controller
    kp.t_start = s->t0, kp.t_end = s->tf, kp.N = s->N, kp.count_events = 1;
    kp.rtol = s->rtol, kp.alpha = s->alpha, kp.pw = s->pw, kp.min_dt = s->min_dt, kp.max_dt = s->max_dt;
    kp.pw_is_third = s->pw == 1.0 / 3.0;
    kp.split_mask = s->split_mask;
    return dispatch_exp(s->sp, kp, s->psi, nullptr, s->gp, nullptr, s->ca, s->ev_dev);
}

void exp_res_add(vo_step_result* res, const EvSlot& e, int launches) {
    if (!res) return;
    res->n_step += (int64_t)e.n_step, res->n_chkpt += (int64_t)e.n_chkpt, res->n_reject += (int64_t)e.n_reject, res->n_end += (int64_t)e.n_end;
    res->launches += launches;
}

int32_t exp_step_impl(vo_expsolver s, bool adaptive, vo_step_result* res) {
    if (!s) return VO_ERR_BAD_ARG;
    DeviceGuard g(s->ctx->device);
    if (res) std::memset(res, 0, sizeof *res);
    int32_t r = exp_launch_event(s, adaptive);
    if (r != VO_OK) return r;
    EvSlot d;
    r = exp_ev_read(s, &d);
    if (r != VO_OK) return r;
    s->n_done += (int64_t)d.n_end;
    exp_res_add(res, d, 1);
    if (res) res->n_active = s->N - s->n_done, res->state = res->n_active == 0 ? VO_STATE_DONE : VO_STATE_OK;
    return VO_OK;
}

}  // namespace

extern "C" {

int32_t vo_split_basis_create(vo_ctx c, int32_t n, int32_t M, const double* basis, vo_split* out) {
    if (!c || !basis || !out || M < 1 || M > VO_EXP_MAX_M || n < 8 || n % 8) return vo_fail(c, VO_ERR_BAD_ARG, "vo_split_basis_create: need n % 8 == 0 and 1 <= M <= 4");
    DeviceGuard g(c->device);
    vo_split sp = new vo_split_s();
    sp->ctx = c, sp->n = n, sp->M = M;
    std::memset(sp->cs, 0, sizeof sp->cs);
    const int NW = n / 8, NK = n / 4;
    std::vector<double> frag((size_t)M * 2 * n * n);
    for (int m = 0; m < M; ++m) {
        double best = 0.0;  // induced 1-norm: max column sum of |B_rc|
        for (int col = 0; col < n; ++col) {
            double sum = 0.0;
            for (int r = 0; r < n; ++r) sum += std::hypot(basis[(((size_t)m * n + r) * n + col) * 2], basis[(((size_t)m * n + r) * n + col) * 2 + 1]);
            best = std::max(best, sum);
        }
        sp->norm1[m] = best;
        for (int p = 0; p < 2; ++p)
            for (int w = 0; w < NW; ++w)
                for (int kk = 0; kk < NK; ++kk)
                    for (int l = 0; l < 32; ++l)  // A fragment of mma.m8n8k4: lane l holds A[l/4][l%4]
                        frag[((((size_t)(m * 2 + p) * NW + w) * NK + kk) * 32) + l] = basis[(((size_t)m * n + 8 * w + l / 4) * n + 4 * kk + l % 4) * 2 + p];
    }
    if (cudaMalloc(&sp->frag_dev, frag.size() * sizeof(double)) != cudaSuccess) {
        delete sp;
        return vo_fail(c, VO_ERR_ALLOC, "vo_split_basis_create: cudaMalloc failed");
    }
    VO_CUDA(c, cudaMemcpyAsync(sp->frag_dev, frag.data(), frag.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = sp;
    return VO_OK;
}

int32_t vo_split_destroy(vo_split sp) {
    if (!sp) return VO_OK;
    DeviceGuard g(sp->ctx->device);
    cudaStreamSynchronize(sp->ctx->stream);
    cudaFree(sp->frag_dev);
    delete sp;
    return VO_OK;
}

int32_t vo_split_set_commutator(vo_split sp, const double* cs) {
    if (!sp || !cs) return VO_ERR_BAD_ARG;
    std::memcpy(sp->cs, cs, sizeof(double) * sp->M * sp->M * sp->M);
    sp->has_cs = true;
    return VO_OK;
}

int32_t vo_split_set_taylor_degree(vo_split sp, int32_t deg) {
    if (!sp || deg < 0 || deg > 60) return vo_fail(sp ? sp->ctx : nullptr, VO_ERR_BAD_ARG, "vo_split_set_taylor_degree: 0 <= deg <= 60");
    sp->taylor_deg = deg;
    return VO_OK;
}

int32_t vo_split_norm(vo_split sp, const void* psi_dev, int64_t N, double* out_host) {
    if (!sp || !psi_dev || !out_host || N < 1) return vo_fail(sp ? sp->ctx : nullptr, VO_ERR_BAD_ARG, "vo_split_norm: bad argument");
    vo_ctx c = sp->ctx;
    DeviceGuard g(c->device);
    double* out_dev = nullptr;
    VO_CUDA(c, cudaMallocAsync(&out_dev, sizeof(double) * N, c->stream));
    exp_norm_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, c->stream>>>((const double2*)psi_dev, sp->n, N, out_dev);
    VO_CHECK_LAUNCH(c);
    VO_CUDA(c, cudaMemcpyAsync(out_host, out_dev, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    cudaFreeAsync(out_dev, c->stream);
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    return VO_OK;
}

// Commutator::commutator (exp/mod.rs:47-54) on coefficient vectors: out_c = sum_ab la_a lb_b cs[a][b][c]. Operators of this
// split ARE their M coefficients, so this is host-side bookkeeping on N*M complex numbers, not a device computation.
int32_t vo_split_commutator(vo_split sp, const double* la, const double* lb, int64_t N, double* out) {
    if (!sp || !la || !lb || !out || N < 1) return vo_fail(sp ? sp->ctx : nullptr, VO_ERR_BAD_ARG, "vo_split_commutator: bad argument");
    if (!sp->has_cs) return vo_fail(sp->ctx, VO_ERR_STATE, "vo_split_commutator: no structure tensor (vo_split_set_commutator)");
    const int M = sp->M;
    for (int64_t i = 0; i < N; ++i)
        for (int cc = 0; cc < M; ++cc) {
            std::complex<double> acc(0.0, 0.0);
            for (int a = 0; a < M; ++a)
                for (int b = 0; b < M; ++b) {
                    const double sc = sp->cs[(a * M + b) * M + cc];
                    if (sc != 0.0)
                        acc += std::complex<double>(la[(i * M + a) * 2], la[(i * M + a) * 2 + 1]) * std::complex<double>(lb[(i * M + b) * 2], lb[(i * M + b) * 2 + 1]) * sc;
                }
            out[(i * M + cc) * 2] = acc.real(), out[(i * M + cc) * 2 + 1] = acc.imag();
        }
    return VO_OK;
}

int32_t vo_map_exp_seq(vo_split sp, const double* coef_host, int32_t K, int64_t N, void* psi_in_dev, void* psi_out_dev) {
    if (!sp || !coef_host || N < 1 || K < 1 || !psi_in_dev || !psi_out_dev) return vo_fail(sp ? sp->ctx : nullptr, VO_ERR_BAD_ARG, "vo_map_exp_seq: bad argument");
    vo_ctx c = sp->ctx;
    DeviceGuard g(c->device);
    double2* coef_dev = nullptr;
    const size_t bytes = sizeof(double2) * (size_t)K * N * sp->M;
    VO_CUDA(c, cudaMallocAsync(&coef_dev, bytes, c->stream));
    VO_CUDA(c, cudaMemcpyAsync(coef_dev, coef_host, bytes, cudaMemcpyHostToDevice, c->stream));
    ExpKP kp = make_kp(sp, 1);
    kp.N = N, kp.nseq = K;
    CtlArrays none{};
    int32_t r = dispatch_exp(sp, kp, (double2*)psi_in_dev, (double2*)psi_out_dev, nullptr, coef_dev, none, nullptr);
    cudaFreeAsync(coef_dev, c->stream);
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    return r;
}

int32_t vo_map_exp(vo_split sp, const double* coef_host, int64_t N, void* psi_in_dev, void* psi_out_dev) {
    return vo_map_exp_seq(sp, coef_host, 1, N, psi_in_dev, psi_out_dev);
}

int32_t vo_exp_set_split_mask(vo_expsolver s, uint32_t a_mask) {
    if (!s) return VO_ERR_BAD_ARG;
    s->split_mask = a_mask;
    return VO_OK;
}

int32_t vo_exp_create(vo_ctx c, vo_split sp, int32_t scheme, int32_t M_gen, const double* gp_host, int64_t N, double t0, double tf,
                      const double* psi0_host, double h, vo_expsolver* out) {
    if (!c || !sp || !out || !psi0_host || N < 1 || scheme < 0 || scheme > VO_EXP_SPLIT_MIDPOINT || M_gen < 1 || M_gen > sp->M || (M_gen > 1 && !gp_host))
        return vo_fail(c, VO_ERR_BAD_ARG, "vo_exp_create: bad argument");
    if (scheme == VO_EXP_MAGNUS42 && !sp->has_cs) return vo_fail(c, VO_ERR_BAD_ARG, "vo_exp_create: Magnus needs vo_split_set_commutator (Commutator trait, exp/mod.rs:47-54)");
    DeviceGuard g(c->device);
    vo_expsolver s = new vo_expsolver_s();
    s->ctx = c, s->sp = sp, s->scheme = scheme, s->M_gen = M_gen, s->N = N, s->t0 = t0, s->tf = tf, s->h_init = h;
    const size_t nb = sizeof(double2) * (size_t)N * sp->n, ngp = sizeof(double) * (size_t)N * std::max(1, M_gen - 1) * 3;
    bool ok = cudaMalloc(&s->psi, nb) == cudaSuccess && cudaMalloc(&s->psi0, nb) == cudaSuccess && cudaMalloc(&s->gp, ngp) == cudaSuccess &&
              cudaMalloc(&s->ca.t, 8 * N) == cudaSuccess && cudaMalloc(&s->ca.h, 8 * N) == cudaSuccess && cudaMalloc(&s->ca.prev_h, 8 * N) == cudaSuccess &&
              cudaMalloc(&s->ca.dx_norm, 8 * N) == cudaSuccess && cudaMalloc(&s->ca.n_accept, 4 * N) == cudaSuccess &&
              cudaMalloc(&s->ca.n_reject, 4 * N) == cudaSuccess && cudaMalloc(&s->ca.word, 4 * N) == cudaSuccess &&
              cudaMalloc(&s->ev_dev, sizeof(EvSlot) * VO_EV_SLOTS) == cudaSuccess && cudaMallocHost(&s->ev_host, sizeof(EvSlot) * VO_EV_SLOTS) == cudaSuccess;
    if (!ok) {
        vo_exp_destroy(s);
        return vo_fail(c, VO_ERR_ALLOC, "vo_exp_create: allocation failed");
    }
    VO_CUDA(c, cudaMemcpyAsync(s->psi0, psi0_host, nb, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(c, cudaMemcpyAsync(s->psi, s->psi0, nb, cudaMemcpyDeviceToDevice, c->stream));
    if (M_gen > 1) VO_CUDA(c, cudaMemcpyAsync(s->gp, gp_host, ngp, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(c, cudaMemsetAsync(s->ev_dev, 0, sizeof(EvSlot) * VO_EV_SLOTS, c->stream));
    exp_ctl_fill_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, c->stream>>>(s->ca, N, t0, h);
    VO_CHECK_LAUNCH(c);
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = s;
    return VO_OK;
}

int32_t vo_exp_destroy(vo_expsolver s) {
    if (!s) return VO_OK;
    DeviceGuard g(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->psi), cudaFree(s->psi0), cudaFree(s->gp);
    cudaFree(s->ca.t), cudaFree(s->ca.h), cudaFree(s->ca.prev_h), cudaFree(s->ca.dx_norm), cudaFree(s->ca.n_accept), cudaFree(s->ca.n_reject), cudaFree(s->ca.word);
    cudaFree(s->ev_dev), cudaFreeHost(s->ev_host);
    delete s;
    return VO_OK;
}

int32_t vo_exp_no_adaptive(vo_expsolver s) {
    if (!s) return VO_ERR_BAD_ARG;
    s->want_err = false;  // exp/cfm.rs:157-161: alph_err = None
    return VO_OK;
}

int32_t vo_exp_with_tolerance(vo_expsolver s, double atol, double rtol) {
    if (!s) return VO_ERR_BAD_ARG;
    if (!(atol > 0.0) || !(rtol > 0.0)) return vo_fail(s->ctx, VO_ERR_BAD_ARG, "Invalid tolerances: atol=" + std::to_string(atol) + ", rtol=" + std::to_string(rtol));
    s->atol = atol, s->rtol = rtol;
    return VO_OK;
}

int32_t vo_exp_with_step_range(vo_expsolver s, double dt_min, double dt_max) {
    if (!s) return VO_ERR_BAD_ARG;
    if (!(dt_min > 0.0) || !(dt_max > 0.0) || !(dt_max > dt_min))
        return vo_fail(s->ctx, VO_ERR_BAD_ARG, "Invalid step range: (" + std::to_string(dt_min) + ", " + std::to_string(dt_max) + ")");
    vo_ctx c = s->ctx;
    DeviceGuard g(c->device);
    s->min_dt = dt_min, s->max_dt = dt_max;
    s->h_init = std::sqrt(dt_min * dt_max);  // ode.rs:273-280
    exp_ctl_fill_kernel<<<(unsigned)ceil_div(s->N, 256), 256, 0, c->stream>>>(s->ca, s->N, s->t0, s->h_init);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

int32_t vo_exp_step(vo_expsolver s, vo_step_result* res) { return exp_step_impl(s, false, res); }
int32_t vo_exp_step_adaptive(vo_expsolver s, vo_step_result* res) { return exp_step_impl(s, true, res); }

int32_t vo_exp_run(vo_expsolver s, int32_t adaptive, int64_t max_calls, vo_step_result* res) {
    if (!s) return VO_ERR_BAD_ARG;
    vo_ctx c = s->ctx;
    DeviceGuard g(c->device);
    vo_step_result acc;
    std::memset(&acc, 0, sizeof acc);
    int64_t calls = 0;
    int batch = 2;
    while (s->n_done < s->N && (max_calls <= 0 || calls < max_calls)) {
        int launched = 0;
        for (int b = 0; b < batch && (max_calls <= 0 || calls < max_calls); ++b, ++calls, ++launched) {
            int32_t r = exp_launch_event(s, adaptive != 0);
            if (r != VO_OK) return r;
        }
        EvSlot d;
        int32_t r = exp_ev_read(s, &d);
        if (r != VO_OK) return r;
        s->n_done += (int64_t)d.n_end;
        exp_res_add(&acc, d, launched);
        if (d.n_stuck && d.n_step == 0 && d.n_chkpt == 0 && d.n_end == 0) {
            acc.n_active = s->N - s->n_done, acc.state = VO_STATE_ERR;
            if (res) *res = acc;
            return vo_fail(c, VO_ERR_STATE, "vo_exp_run: all remaining systems are rejected at h == min_dt");
        }
        batch = std::min(batch * 2, 16);
    }
    acc.n_active = s->N - s->n_done, acc.state = acc.n_active == 0 ? VO_STATE_DONE : VO_STATE_OK;
    if (res) *res = acc;
    return VO_OK;
}

int32_t vo_exp_current(vo_expsolver s, double* t_min, double* t_max, double* psi_host) {
    if (!s) return VO_ERR_BAD_ARG;
    vo_ctx c = s->ctx;
    DeviceGuard g(c->device);
    if (t_min || t_max) {
        std::vector<double> t((size_t)s->N);
        VO_CUDA(c, cudaMemcpyAsync(t.data(), s->ca.t, 8 * (size_t)s->N, cudaMemcpyDeviceToHost, c->stream));
        VO_CUDA(c, cudaStreamSynchronize(c->stream));
        const auto mm = std::minmax_element(t.begin(), t.end());
        if (t_min) *t_min = *mm.first;
        if (t_max) *t_max = *mm.second;
    }
    if (psi_host) {
        VO_CUDA(c, cudaMemcpyAsync(psi_host, s->psi, sizeof(double2) * (size_t)s->N * s->sp->n, cudaMemcpyDeviceToHost, c->stream));
        VO_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return VO_OK;
}

int32_t vo_exp_stats(vo_expsolver s, int64_t* accepted, int64_t* rejected, double* t, double* h, double* dx_norm) {
    if (!s) return VO_ERR_BAD_ARG;
    vo_ctx c = s->ctx;
    DeviceGuard g(c->device);
    const size_t n = (size_t)s->N;
    std::vector<uint32_t> tmp(n);
    if (accepted) {
        VO_CUDA(c, cudaMemcpyAsync(tmp.data(), s->ca.n_accept, 4 * n, cudaMemcpyDeviceToHost, c->stream));
        VO_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < n; ++i) accepted[i] = tmp[i];
    }
    if (rejected) {
        VO_CUDA(c, cudaMemcpyAsync(tmp.data(), s->ca.n_reject, 4 * n, cudaMemcpyDeviceToHost, c->stream));
        VO_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < n; ++i) rejected[i] = tmp[i];
    }
    if (t) VO_CUDA(c, cudaMemcpyAsync(t, s->ca.t, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    if (h) VO_CUDA(c, cudaMemcpyAsync(h, s->ca.h, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    if (dx_norm) VO_CUDA(c, cudaMemcpyAsync(dx_norm, s->ca.dx_norm, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    return VO_OK;
}

int32_t vo_exp_reset(vo_expsolver s, const double* psi0_host) {
    if (!s) return VO_ERR_BAD_ARG;
    vo_ctx c = s->ctx;
    DeviceGuard g(c->device);
    const size_t nb = sizeof(double2) * (size_t)s->N * s->sp->n;
    if (psi0_host) VO_CUDA(c, cudaMemcpyAsync(s->psi0, psi0_host, nb, cudaMemcpyHostToDevice, c->stream));
    VO_CUDA(c, cudaMemcpyAsync(s->psi, s->psi0, nb, cudaMemcpyDeviceToDevice, c->stream));
    VO_CUDA(c, cudaMemsetAsync(s->ev_dev, 0, sizeof(EvSlot) * VO_EV_SLOTS, c->stream));
    std::memset(&s->ev_seen, 0, sizeof s->ev_seen);
    s->n_done = 0;
    exp_ctl_fill_kernel<<<(unsigned)ceil_div(s->N, 256), 256, 0, c->stream>>>(s->ca, s->N, s->t0, s->h_init);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

void* vo_exp_state_device_ptr(vo_expsolver s) { return s ? (void*)s->psi : nullptr; }

}  // extern "C"
