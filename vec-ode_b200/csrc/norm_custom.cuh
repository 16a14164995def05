// norm_custom.cuh — user-defined norms: the reference's `Normed<T, V>` (src/base/ode.rs:9-11; RK45Solver::norm, rk.rs:302-304)
// and ExpCFMSolver's NormFn closure (src/exp/cfm.rs:105, 214-216) are the USER's. Like the RHS closure, a norm cannot cross a C
// ABI into device code as a function pointer, so it crosses as source, in the shape every norm of a vector has on a parallel
// machine:      norm(e) = finish( JOIN_i map(e_i, i) ),     JOIN = + or max.
// The functor NF is generated from the caller's statements (nvrtc_rhs.cu: norm_source) and compiled at run time
//   * into the register-resident control kernels (rk_small.cuh: err_norm, kind VO_NORM_CUSTOM),
//   * into exp_step_kernel (exp_kernels.cuh, phase D), and
//   * into the reduction kernels below (stage path: one large state, or per-trajectory control with d > 8; vo_norm_custom).
// Components are visited left to right in STRICT arithmetic (one thread per trajectory), in a fixed tree otherwise, so results
// are reproducible run to run.
#pragma once
#include "common.cuh"

template <class NF> __device__ __forceinline__ double nf_join(double a, double b) { return NF::JOIN == 1 ? fmax(a, b) : a + b; }

// one thread per trajectory, left to right over the components (coalesced over i); also the strict-order path
template <class NF>
__global__ void norm_small_custom_kernel(const double* __restrict__ x, int64_t d, int64_t n, double* __restrict__ out, int finish) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int64_t c = 0; c < d; ++c) acc = nf_join<NF>(acc, NF::map(x[c * n + i], 0.0, (int)c, (int)d));
    out[i] = finish ? NF::finish(acc, (int)d) : acc;
}

// large d: grid (chunks, trajectories); warp-shuffle tree inside each block; fixed order
template <class NF>
__global__ void __launch_bounds__(256) norm_partial_custom_kernel(const double* __restrict__ x, int64_t d, int64_t n, int64_t i_off, int64_t d_total,
                                                                  double* __restrict__ partial) {
    const int64_t traj = blockIdx.y;
    const int64_t chunk = (d + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * chunk, hi = min(d, lo + chunk);
    double acc = 0.0;
    for (int64_t c = lo + threadIdx.x; c < hi; c += blockDim.x) acc = nf_join<NF>(acc, NF::map(x[c * n + traj], 0.0, (int)(c + i_off), (int)d_total));
    for (int off = 16; off > 0; off >>= 1) acc = nf_join<NF>(acc, __shfl_down_sync(0xffffffffu, acc, off));
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        for (int off = 4; off > 0; off >>= 1) acc = nf_join<NF>(acc, __shfl_down_sync(0xffffffffu, acc, off));
        if (threadIdx.x == 0) partial[traj * gridDim.x + blockIdx.x] = acc;
    }
}
template <class NF>
__global__ void norm_final_custom_kernel(const double* __restrict__ partial, int chunks, int64_t d_total, double* __restrict__ out, int finish) {
    const int64_t traj = blockIdx.x;
    double acc = 0.0;
    for (int c = threadIdx.x; c < chunks; c += 32) acc = nf_join<NF>(acc, partial[traj * chunks + c]);
    for (int off = 16; off > 0; off >>= 1) acc = nf_join<NF>(acc, __shfl_down_sync(0xffffffffu, acc, off));
    if (threadIdx.x == 0) out[traj] = finish ? NF::finish(acc, (int)d_total) : acc;
}
