// rhs.cuh — device functors replacing the reference's RHS closure `f(t, &x, &mut dx)` (src/base/rk.rs:97).
// A closure cannot cross a C ABI into device code, so the RHS is a compiled-in functor selected by id with
// parameters that are shared scalars or per-trajectory arrays. Operation order is part of the contract: it is
// the same as `Rhs::operator()` in oracle/vecode_oracle.cpp (checked bit-for-bit in STRICT mode).
#pragma once
#include "common.cuh"

template <int KIND, int DIM> struct RhsF;

// dx_c = p_c * x_c  — the family of the reference's own tests (src/impls/nalgebra.rs:54-58, 74-78, 93-96).
template <int DIM> struct RhsF<VO_RHS_DIAG_LINEAR, DIM> {
    static constexpr int D = DIM, NP = DIM;
    template <bool STRICT> static __device__ __forceinline__ void eval(double, const double (&x)[D], double (&dx)[D], const double (&p)[NP]) {
#pragma unroll
        for (int c = 0; c < D; ++c) dx[c] = Ar<STRICT>::mul(p[c], x[c]);
    }
};

template <> struct RhsF<VO_RHS_HARMONIC2D, 2> {
    static constexpr int D = 2, NP = 1;
    template <bool STRICT> static __device__ __forceinline__ void eval(double, const double (&x)[2], double (&dx)[2], const double (&p)[1]) {
        dx[0] = x[1];
        dx[1] = -Ar<STRICT>::mul(p[0], x[0]);
    }
};

template <> struct RhsF<VO_RHS_LORENZ63, 3> {
    static constexpr int D = 3, NP = 3;
    template <bool STRICT> static __device__ __forceinline__ void eval(double, const double (&x)[3], double (&dx)[3], const double (&p)[3]) {
        using A = Ar<STRICT>;
        dx[0] = A::mul(p[0], A::sub(x[1], x[0]));
        dx[1] = A::sub(A::mul(x[0], A::sub(p[1], x[2])), x[1]);
        dx[2] = A::sub(A::mul(x[0], x[1]), A::mul(p[2], x[2]));
    }
};

template <> struct RhsF<VO_RHS_VDP, 2> {
    static constexpr int D = 2, NP = 1;
    template <bool STRICT> static __device__ __forceinline__ void eval(double, const double (&x)[2], double (&dx)[2], const double (&p)[1]) {
        using A = Ar<STRICT>;
        dx[0] = x[1];
        dx[1] = A::sub(A::mul(A::mul(p[0], A::sub(1.0, A::mul(x[0], x[0]))), x[1]), x[0]);
    }
};

// Loads the NP parameters of trajectory i (per-trajectory array if set, else the shared scalar).
template <int NP> static __device__ __forceinline__ void load_params(const RhsParams& rp, int64_t i, double (&p)[NP]) {
#pragma unroll
    for (int q = 0; q < NP; ++q) p[q] = rp.per_traj[q] ? __ldg(rp.per_traj[q] + i) : rp.shared[q];
}

#ifndef __CUDACC_RTC__
static inline RhsParams make_rhs_params(const vo_rhs_s* r) {
    RhsParams rp;
    for (int q = 0; q < VO_MAX_PARAMS; ++q) rp.shared[q] = r->shared[q], rp.per_traj[q] = r->per_traj[q];
    return rp;
}
#endif
