// lc.cu — LinearCombination kernels (src/lc.rs:7-55) and per-trajectory norms (src/base/ode.rs:9-11).
//
// All of these are element-wise streams over d*N doubles: HBM-bound. Each thread moves 128-bit vectors
// (double2 -> LDG.128/STG.128), two per iteration for memory-level parallelism, in a grid-stride loop over a
// grid sized to a multiple of the SM count. The n-term reducer reads every operand once and writes once
// (n+1 passes) where the reference's chain of add_scalar_mul calls makes 2 + 3(n-1) passes.
#include "common.cuh"

namespace {

constexpr int LC_THREADS = 256;

struct LcTerms {
    const double* v[VO_MAX_TERMS];
    double k[VO_MAX_TERMS];
    int n;
};

enum LcOp { OP_SCALE, OP_SMUL_TO, OP_AXPY, OP_ADD, OP_DELTA };

template <int OP, bool STRICT> __device__ __forceinline__ double lc_apply(double v, double u, double k) {
    using A = Ar<STRICT>;
    if (OP == OP_SCALE) return A::mul(v, k);          // *self *= k           ndarray.rs:15
    if (OP == OP_SMUL_TO) return A::mul(k, u);        // *t = k * s           ndarray.rs:19
    if (OP == OP_AXPY) return A::axpy(v, k, u);       // *y = *y + (k * *x)   ndarray.rs:23
    if (OP == OP_ADD) return A::add(v, u);            // *self += other       ndarray.rs:27
    return A::sub(v, u);                              // *self -= y           ndarray.rs:31
}

// v (in/out) and u (in). For OP_SMUL_TO `v` is the target and is write-only.
template <int OP, bool STRICT>
__global__ void __launch_bounds__(LC_THREADS) lc_binary_kernel(double* __restrict__ v, const double* __restrict__ u, double k, int64_t n) {
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double2* v2 = reinterpret_cast<double2*>(v);
    const double2* u2 = reinterpret_cast<const double2*>(u);
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + stride < n2; i += 2 * stride) {
        double2 a0 = {0, 0}, a1 = {0, 0}, b0 = {0, 0}, b1 = {0, 0};
        if (OP != OP_SMUL_TO) a0 = v2[i], a1 = v2[i + stride];
        if (OP != OP_SCALE) b0 = u2[i], b1 = u2[i + stride];
        a0.x = lc_apply<OP, STRICT>(a0.x, b0.x, k), a0.y = lc_apply<OP, STRICT>(a0.y, b0.y, k);
        a1.x = lc_apply<OP, STRICT>(a1.x, b1.x, k), a1.y = lc_apply<OP, STRICT>(a1.y, b1.y, k);
        v2[i] = a0, v2[i + stride] = a1;
    }
    for (; i < n2; i += stride) {
        double2 a0 = {0, 0}, b0 = {0, 0};
        if (OP != OP_SMUL_TO) a0 = v2[i];
        if (OP != OP_SCALE) b0 = u2[i];
        a0.x = lc_apply<OP, STRICT>(a0.x, b0.x, k), a0.y = lc_apply<OP, STRICT>(a0.y, b0.y, k);
        v2[i] = a0;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t e = n - 1;
        v[e] = lc_apply<OP, STRICT>(OP != OP_SMUL_TO ? v[e] : 0.0, OP != OP_SCALE ? u[e] : 0.0, k);
    }
}

// v = k0*v0; v = v + (kj*vj) ...; optionally v = v*dt; v = v + x0   (lc.rs:20-54, rk.rs:121-124)
template <int NT, bool STRICT, bool STAGE>
__global__ void __launch_bounds__(LC_THREADS) lc_lincomb_kernel(double* __restrict__ v, const __grid_constant__ LcTerms tm, double dt,
                                                                const double* __restrict__ x0, int64_t n) {
    using A = Ar<STRICT>;
    const int nt = NT > 0 ? NT : tm.n;
    const int64_t n2 = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 t[NT > 0 ? NT : 1];
        double2 acc;
        if (NT > 0) {
#pragma unroll
            for (int j = 0; j < NT; ++j) t[j] = reinterpret_cast<const double2*>(tm.v[j])[i];
            acc.x = A::mul(tm.k[0], t[0].x), acc.y = A::mul(tm.k[0], t[0].y);
#pragma unroll
            for (int j = 1; j < NT; ++j) acc.x = A::axpy(acc.x, tm.k[j], t[j].x), acc.y = A::axpy(acc.y, tm.k[j], t[j].y);
        } else {
            double2 a = reinterpret_cast<const double2*>(tm.v[0])[i];
            acc.x = A::mul(tm.k[0], a.x), acc.y = A::mul(tm.k[0], a.y);
            for (int j = 1; j < nt; ++j) {
                a = reinterpret_cast<const double2*>(tm.v[j])[i];
                acc.x = A::axpy(acc.x, tm.k[j], a.x), acc.y = A::axpy(acc.y, tm.k[j], a.y);
            }
        }
        if (STAGE) {
            const double2 b = reinterpret_cast<const double2*>(x0)[i];
            acc.x = A::add(A::mul(acc.x, dt), b.x), acc.y = A::add(A::mul(acc.y, dt), b.y);
        }
        reinterpret_cast<double2*>(v)[i] = acc;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t e = n - 1;
        double acc = A::mul(tm.k[0], tm.v[0][e]);
        for (int j = 1; j < nt; ++j) acc = A::axpy(acc, tm.k[j], tm.v[j][e]);
        if (STAGE) acc = A::add(A::mul(acc, dt), x0[e]);
        v[e] = acc;
    }
}

int lc_grid(vo_ctx c, int64_t n) {
    const int64_t want = ceil_div(n / 2 + 1, (int64_t)LC_THREADS * 2);
    const int64_t cap = (int64_t)c->sm_count * 8;  // 8 resident 256-thread CTAs per SM
    return (int)std::max<int64_t>(1, std::min(want, cap));
}

template <int OP> int32_t launch_binary(vo_ens v, vo_ens u, double k) {
    vo_ctx c = v->ctx;
    DeviceGuard g(c->device);
    const int64_t n = v->elems();
    const int grid = lc_grid(c, n);
    if (c->arith == VO_ARITH_STRICT)
        lc_binary_kernel<OP, true><<<grid, LC_THREADS, 0, c->stream>>>(v->p, u ? u->p : nullptr, k, n);
    else
        lc_binary_kernel<OP, false><<<grid, LC_THREADS, 0, c->stream>>>(v->p, u ? u->p : nullptr, k, n);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

bool same_shape(vo_ens a, vo_ens b) { return a && b && a->d == b->d && a->n == b->n && a->ctx == b->ctx; }

template <bool STRICT, bool STAGE> void launch_lincomb(vo_ctx c, int grid, double* v, const LcTerms& tm, double dt, const double* x0, int64_t n) {
#define VO_LC_CASE(NT) \
    case NT: lc_lincomb_kernel<NT, STRICT, STAGE><<<grid, LC_THREADS, 0, c->stream>>>(v, tm, dt, x0, n); break;
    switch (tm.n) {
        VO_LC_CASE(1) VO_LC_CASE(2) VO_LC_CASE(3) VO_LC_CASE(4) VO_LC_CASE(5) VO_LC_CASE(6) VO_LC_CASE(7) VO_LC_CASE(8)
        default: lc_lincomb_kernel<0, STRICT, STAGE><<<grid, LC_THREADS, 0, c->stream>>>(v, tm, dt, x0, n);
    }
#undef VO_LC_CASE
}

int32_t lincomb_impl(vo_ens v, const vo_ens* v_arr, const double* k_arr, int32_t n, bool stage, double dt, vo_ens x0) {
    if (!v) return VO_ERR_BAD_ARG;
    vo_ctx c = v->ctx;
    if (!v_arr || !k_arr || n <= 0) return vo_fail(c, VO_ERR_BAD_ARG, "linear_combination: slices cannot be empty");  // lc.rs:21-23
    if (n > VO_MAX_TERMS) return vo_fail(c, VO_ERR_UNSUPPORTED, "linear_combination: more than VO_MAX_TERMS terms");
    LcTerms tm;
    tm.n = n;
    for (int j = 0; j < n; ++j) {
        if (!same_shape(v, v_arr[j])) return vo_fail(c, VO_ERR_SHAPE, "linear_combination: operand shape mismatch");
        if (v_arr[j]->p == v->p) return vo_fail(c, VO_ERR_BAD_ARG, "linear_combination: target aliases an operand");
        tm.v[j] = v_arr[j]->p, tm.k[j] = k_arr[j];
    }
    if (stage && !same_shape(v, x0)) return vo_fail(c, VO_ERR_SHAPE, "stage_combine: x0 shape mismatch");
    DeviceGuard g(c->device);
    const int64_t ne = v->elems();
    const int grid = lc_grid(c, ne);
    const bool strict = c->arith == VO_ARITH_STRICT;
    if (strict && stage) launch_lincomb<true, true>(c, grid, v->p, tm, dt, x0->p, ne);
    else if (strict) launch_lincomb<true, false>(c, grid, v->p, tm, dt, nullptr, ne);
    else if (stage) launch_lincomb<false, true>(c, grid, v->p, tm, dt, x0->p, ne);
    else launch_lincomb<false, false>(c, grid, v->p, tm, dt, nullptr, ne);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

// ---- complex scalars: LinearCombination<Complex<f64>, V> -------------------------------------------------------------------
// src/impls/ndarray.rs:8-33 is generic over the element type A, and with A = num_complex::Complex<f64> the scalar k is complex too
// (the states and operators of the exponential integrators are such vectors). Elements are stored interleaved (re, im), one
// 16-byte vector per element. Arithmetic as num-complex 0.4 writes it: Mul is (a.re b.re - a.im b.im, a.re b.im + a.im b.re),
// MulAssign is re = re k.re - im k.im, im = im k.re + re k.im (the same two sums; IEEE addition commutes), Add is componentwise.
struct LcTermsZ {
    const double2* v[VO_MAX_TERMS];
    double2 k[VO_MAX_TERMS];
    int n;
};

enum LcOpZ { ZOP_SCALE, ZOP_SMUL_TO, ZOP_AXPY };

template <bool STRICT> __device__ __forceinline__ double2 zmul(double2 a, double2 b) {  // a * b
    using A = Ar<STRICT>;
    return make_double2(A::sub(A::mul(a.x, b.x), A::mul(a.y, b.y)), A::add(A::mul(a.x, b.y), A::mul(a.y, b.x)));
}
template <int OP, bool STRICT> __device__ __forceinline__ double2 lcz_apply(double2 v, double2 u, double2 k) {
    using A = Ar<STRICT>;
    if (OP == ZOP_SCALE) return zmul<STRICT>(v, k);  // *self *= k           ndarray.rs:15
    const double2 p = zmul<STRICT>(k, u);            // k * s                ndarray.rs:19
    if (OP == ZOP_SMUL_TO) return p;
    return make_double2(A::add(v.x, p.x), A::add(v.y, p.y));  // *y = *y + (k * *x)   ndarray.rs:23
}

// nz complex elements; v in/out (write-only for ZOP_SMUL_TO), u in. Two 16-byte elements in flight per thread and iteration.
template <int OP, bool STRICT>
__global__ void __launch_bounds__(LC_THREADS) lcz_binary_kernel(double2* __restrict__ v, const double2* __restrict__ u, double2 k, int64_t nz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + stride < nz; i += 2 * stride) {
        double2 a0 = {0, 0}, a1 = {0, 0}, b0 = {0, 0}, b1 = {0, 0};
        if (OP != ZOP_SMUL_TO) a0 = v[i], a1 = v[i + stride];
        if (OP != ZOP_SCALE) b0 = u[i], b1 = u[i + stride];
        v[i] = lcz_apply<OP, STRICT>(a0, b0, k), v[i + stride] = lcz_apply<OP, STRICT>(a1, b1, k);
    }
    for (; i < nz; i += stride) {
        double2 a0 = {0, 0}, b0 = {0, 0};
        if (OP != ZOP_SMUL_TO) a0 = v[i];
        if (OP != ZOP_SCALE) b0 = u[i];
        v[i] = lcz_apply<OP, STRICT>(a0, b0, k);
    }
}

// v = k0*v0; v = v + (kj*vj) left to right (lc.rs:20-54), one pass
template <bool STRICT> __global__ void __launch_bounds__(LC_THREADS) lcz_lincomb_kernel(double2* __restrict__ v, const __grid_constant__ LcTermsZ tm, int64_t nz) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nz; i += stride) {
        double2 acc = zmul<STRICT>(tm.k[0], tm.v[0][i]);
        for (int j = 1; j < tm.n; ++j) acc = lcz_apply<ZOP_AXPY, STRICT>(acc, tm.v[j][i], tm.k[j]);
        v[i] = acc;
    }
}

// a complex vector is any ensemble whose rows hold whole (re, im) pairs
bool complex_ok(vo_ens e) { return e && e->n % 2 == 0; }

template <int OP> int32_t launch_binary_z(vo_ens v, vo_ens u, double kr, double ki) {
    vo_ctx c = v->ctx;
    if (!complex_ok(v)) return vo_fail(c, VO_ERR_SHAPE, "complex LinearCombination: rows must hold whole (re, im) pairs (n even)");
    DeviceGuard g(c->device);
    const int64_t nz = v->elems() / 2;
    const int grid = lc_grid(c, 2 * nz);
    const double2 k = make_double2(kr, ki);
    if (c->arith == VO_ARITH_STRICT)
        lcz_binary_kernel<OP, true><<<grid, LC_THREADS, 0, c->stream>>>(reinterpret_cast<double2*>(v->p), u ? reinterpret_cast<const double2*>(u->p) : nullptr, k, nz);
    else
        lcz_binary_kernel<OP, false><<<grid, LC_THREADS, 0, c->stream>>>(reinterpret_cast<double2*>(v->p), u ? reinterpret_cast<const double2*>(u->p) : nullptr, k, nz);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

// ---- norms -------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm_term(double e, int kind) { return kind == VO_NORM_L2 ? e * e : fabs(e); }
__device__ __forceinline__ double norm_join(double a, double b, int kind) { return kind == VO_NORM_LINF ? fmax(a, b) : a + b; }

// d <= 64: one thread per trajectory, left-to-right over the components (coalesced over i).
template <bool STRICT> __global__ void norm_small_kernel(const double* __restrict__ x, int64_t d, int64_t n, int kind, double* __restrict__ out, bool finish) {
    using A = Ar<STRICT>;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    if (kind == VO_NORM_HYPOT) {
        if (d == 2) {
            out[i] = hypot(x[i], x[n + i]);
            return;
        }
        for (int64_t c = 0; c + 1 < d; c += 2) {
            const double m = hypot(x[c * n + i], x[(c + 1) * n + i]);
            acc = A::add(acc, A::mul(m, m));
        }
        out[i] = sqrt(acc);
        return;
    }
    for (int64_t c = 0; c < d; ++c) {
        const double e = x[c * n + i];
        if (kind == VO_NORM_L2) acc = A::add(acc, A::mul(e, e));
        else if (kind == VO_NORM_LINF) acc = fmax(acc, fabs(e));
        else acc = A::add(acc, fabs(e));
    }
    out[i] = (kind == VO_NORM_L2 && finish) ? sqrt(acc) : acc;
}

// large d: grid (chunks, trajectories); warp-shuffle tree inside each block; fixed order -> reproducible.
__global__ void __launch_bounds__(256) norm_partial_kernel(const double* __restrict__ x, int64_t d, int64_t n, int kind, double* __restrict__ partial) {
    const int64_t traj = blockIdx.y;
    const int64_t chunk = ceil((double)d / gridDim.x);
    const int64_t lo = blockIdx.x * chunk, hi = min(d, lo + chunk);
    double acc = 0.0;
    for (int64_t c = lo + threadIdx.x; c < hi; c += blockDim.x) acc = norm_join(acc, norm_term(x[c * n + traj], kind), kind);
    for (int off = 16; off > 0; off >>= 1) acc = norm_join(acc, __shfl_down_sync(0xffffffffu, acc, off), kind);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        for (int off = 4; off > 0; off >>= 1) acc = norm_join(acc, __shfl_down_sync(0xffffffffu, acc, off), kind);
        if (threadIdx.x == 0) partial[traj * gridDim.x + blockIdx.x] = acc;
    }
}
__global__ void norm_final_kernel(const double* __restrict__ partial, int chunks, int kind, double* __restrict__ out, bool finish) {
    const int64_t traj = blockIdx.x;
    double acc = 0.0;
    for (int c = threadIdx.x; c < chunks; c += 32) acc = norm_join(acc, partial[traj * chunks + c], kind);
    for (int off = 16; off > 0; off >>= 1) acc = norm_join(acc, __shfl_down_sync(0xffffffffu, acc, off), kind);
    if (threadIdx.x == 0) out[traj] = (kind == VO_NORM_L2 && finish) ? sqrt(acc) : acc;
}

}  // namespace

// Device-side norm into a device buffer (used by the stage-path controller as well). finish = false leaves the reduction
// accumulator (sum of squares for L2) so that several partial states can be combined before the square root (vo_adaptive_try).
int32_t vo_norm_device(vo_ens e, int32_t kind, double* out_dev, double* partial_dev, int partial_cap, bool finish) {
    vo_ctx c = e->ctx;
    if (e->d <= 64) {
        const int grid = (int)ceil_div(e->n, 256);
        if (c->arith == VO_ARITH_STRICT) norm_small_kernel<true><<<grid, 256, 0, c->stream>>>(e->p, e->d, e->n, kind, out_dev, finish);
        else norm_small_kernel<false><<<grid, 256, 0, c->stream>>>(e->p, e->d, e->n, kind, out_dev, finish);
        VO_CHECK_LAUNCH(c);
        return VO_OK;
    }
    if (kind == VO_NORM_HYPOT) return vo_fail(c, VO_ERR_UNSUPPORTED, "vo_norm: HYPOT needs d <= 64");
    int chunks = (int)std::min<int64_t>(std::max<int64_t>(1, e->d / 4096), std::max<int64_t>(1, (int64_t)c->sm_count * 4 / std::max<int64_t>(1, e->n)));
    chunks = std::max(1, std::min(chunks, partial_cap / (int)std::max<int64_t>(1, e->n)));
    if ((int64_t)chunks * e->n > partial_cap) return vo_fail(c, VO_ERR_UNSUPPORTED, "vo_norm: too many trajectories for the large-d path");
    dim3 grid(chunks, (unsigned)e->n);
    norm_partial_kernel<<<grid, 256, 0, c->stream>>>(e->p, e->d, e->n, kind, partial_dev);
    VO_CHECK_LAUNCH(c);
    norm_final_kernel<<<(unsigned)e->n, 32, 0, c->stream>>>(partial_dev, chunks, kind, out_dev, finish);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

extern "C" {

int32_t vo_lc_scale(vo_ens v, double k) {
    if (!v) return VO_ERR_BAD_ARG;
    return launch_binary<OP_SCALE>(v, nullptr, k);
}
int32_t vo_lc_scalar_multiply_to(vo_ens v, double k, vo_ens target) {
    if (!same_shape(v, target)) return vo_fail(v ? v->ctx : nullptr, VO_ERR_SHAPE, "scalar_multiply_to: shape mismatch");
    return launch_binary<OP_SMUL_TO>(target, v, k);
}
int32_t vo_lc_add_scalar_mul(vo_ens v, double k, vo_ens u) {
    if (!same_shape(v, u)) return vo_fail(v ? v->ctx : nullptr, VO_ERR_SHAPE, "add_scalar_mul: shape mismatch");
    return launch_binary<OP_AXPY>(v, u, k);
}
int32_t vo_lc_add_assign_ref(vo_ens v, vo_ens u) {
    if (!same_shape(v, u)) return vo_fail(v ? v->ctx : nullptr, VO_ERR_SHAPE, "add_assign_ref: shape mismatch");
    return launch_binary<OP_ADD>(v, u, 0.0);
}
int32_t vo_lc_delta(vo_ens v, vo_ens y) {
    if (!same_shape(v, y)) return vo_fail(v ? v->ctx : nullptr, VO_ERR_SHAPE, "delta: shape mismatch");
    return launch_binary<OP_DELTA>(v, y, 0.0);
}
int32_t vo_lc_linear_combination(vo_ens v, const vo_ens* v_arr, const double* k_arr, int32_t n) {
    return lincomb_impl(v, v_arr, k_arr, n, false, 0.0, nullptr);
}
int32_t vo_lc_stage_combine(vo_ens v, const vo_ens* v_arr, const double* k_arr, int32_t n, double dt, vo_ens x0) {
    return lincomb_impl(v, v_arr, k_arr, n, true, dt, x0);
}

int32_t vo_lc_scale_z(vo_ens v, double k_re, double k_im) {
    if (!v) return VO_ERR_BAD_ARG;
    return launch_binary_z<ZOP_SCALE>(v, nullptr, k_re, k_im);
}
int32_t vo_lc_scalar_multiply_to_z(vo_ens v, double k_re, double k_im, vo_ens target) {
    if (!same_shape(v, target)) return vo_fail(v ? v->ctx : nullptr, VO_ERR_SHAPE, "scalar_multiply_to: shape mismatch");
    return launch_binary_z<ZOP_SMUL_TO>(target, v, k_re, k_im);
}
int32_t vo_lc_add_scalar_mul_z(vo_ens v, double k_re, double k_im, vo_ens u) {
    if (!same_shape(v, u)) return vo_fail(v ? v->ctx : nullptr, VO_ERR_SHAPE, "add_scalar_mul: shape mismatch");
    return launch_binary_z<ZOP_AXPY>(v, u, k_re, k_im);
}
int32_t vo_lc_linear_combination_z(vo_ens v, const vo_ens* v_arr, const double* k_arr, int32_t n) {
    if (!v) return VO_ERR_BAD_ARG;
    vo_ctx c = v->ctx;
    if (!v_arr || !k_arr || n <= 0) return vo_fail(c, VO_ERR_BAD_ARG, "linear_combination: slices cannot be empty");  // lc.rs:21-23
    if (n > VO_MAX_TERMS) return vo_fail(c, VO_ERR_UNSUPPORTED, "linear_combination: more than VO_MAX_TERMS terms");
    if (!complex_ok(v)) return vo_fail(c, VO_ERR_SHAPE, "complex LinearCombination: rows must hold whole (re, im) pairs (n even)");
    LcTermsZ tm;
    tm.n = n;
    for (int j = 0; j < n; ++j) {
        if (!same_shape(v, v_arr[j])) return vo_fail(c, VO_ERR_SHAPE, "linear_combination: operand shape mismatch");
        if (v_arr[j]->p == v->p) return vo_fail(c, VO_ERR_BAD_ARG, "linear_combination: target aliases an operand");
        tm.v[j] = reinterpret_cast<const double2*>(v_arr[j]->p), tm.k[j] = make_double2(k_arr[2 * j], k_arr[2 * j + 1]);
    }
    DeviceGuard g(c->device);
    const int64_t nz = v->elems() / 2;
    const int grid = lc_grid(c, 2 * nz);
    if (c->arith == VO_ARITH_STRICT) lcz_lincomb_kernel<true><<<grid, LC_THREADS, 0, c->stream>>>(reinterpret_cast<double2*>(v->p), tm, nz);
    else lcz_lincomb_kernel<false><<<grid, LC_THREADS, 0, c->stream>>>(reinterpret_cast<double2*>(v->p), tm, nz);
    VO_CHECK_LAUNCH(c);
    return VO_OK;
}

int32_t vo_norm(vo_ens e, int32_t kind, double* out_host) {
    if (!e || !out_host || kind < 0 || kind > VO_NORM_HYPOT) return vo_fail(e ? e->ctx : nullptr, VO_ERR_BAD_ARG, "vo_norm: bad argument");
    vo_ctx c = e->ctx;
    DeviceGuard g(c->device);
    double *out_dev = nullptr, *partial = nullptr;
    const int cap = 1 << 16;
    VO_CUDA(c, cudaMallocAsync(&out_dev, sizeof(double) * e->n, c->stream));
    VO_CUDA(c, cudaMallocAsync(&partial, sizeof(double) * cap, c->stream));
    int32_t r = vo_norm_device(e, kind, out_dev, partial, cap, true);
    if (r == VO_OK) {
        cudaError_t ce = cudaMemcpyAsync(out_host, out_dev, sizeof(double) * e->n, cudaMemcpyDeviceToHost, c->stream);
        if (ce != cudaSuccess) r = vo_fail(c, VO_ERR_CUDA, cudaGetErrorString(ce));
    }
    cudaFreeAsync(out_dev, c->stream), cudaFreeAsync(partial, c->stream);
    VO_CUDA(c, cudaStreamSynchronize(c->stream));
    return r;
}

}  // extern "C"
