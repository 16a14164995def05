// vecode_oracle.cpp — CPU restatement of hmunozb/vec-ode's time-stepping path.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product (`vec-ode_b200/`) may include, link or call
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs load the shared object built from it.
//
// PARITY UNPINNED: the reference is a Rust crate, there is no Rust toolchain in this image, and the
// crate's only tests (src/impls/nalgebra.rs:52-107) print results and assert nothing. This file is a
// source-faithful restatement (same operations, same order, one vector pass per LinearCombination
// call, no FMA contraction: build with -O2 -ffp-contract=off, no fast-math). It is cross-checked
// bit-for-bit against an independent pure-Python restatement (oracle/vecode_oracle.py) and frozen
// golden vectors (tests/golden/), but never against the compiled crate itself.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <utility>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Scalars. RK45RealSolver uses S = T = f64; RK45ComplexSolver uses S = Complex<T> (src/base/rk.rs:217-218)
// and converts every real coefficient with `.into()` before multiplying (src/base/rk.rs:102, lc.rs:45-50),
// so products are full complex multiplications with a zero imaginary part.
// ---------------------------------------------------------------------------------------------
struct cplx {
    double re, im;
};
// num-complex 0.4 `impl Mul for Complex`: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re)
static inline cplx operator*(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
static inline cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }

template <class E> struct scalar_of;
template <> struct scalar_of<double> {
    static double from_real(double x) { return x; }
};
template <> struct scalar_of<cplx> {
    static cplx from_real(double x) { return {x, 0.0}; }  // From<T> for Complex<T>
};

// ---------------------------------------------------------------------------------------------
// LinearCombination (src/lc.rs:7-55) with the element-wise arithmetic of the ndarray impl
// (src/impls/ndarray.rs:14-32) / RK45SolverDefaultLC (src/base/rk.rs:183-201).
// Each call is ONE pass over the vector, exactly like the reference.
// ---------------------------------------------------------------------------------------------
template <class E> struct LC {
    using V = std::vector<E>;
    static void scale(V& v, E k) {  // lc.rs:10 ; ndarray.rs:15 `*self *= k`
        for (auto& e : v) e = e * k;
    }
    static void scalar_multiply_to(const V& v, E k, V& target) {  // lc.rs:12 ; ndarray.rs:19 `*t = k * s`
        for (size_t i = 0; i < v.size(); ++i) target[i] = k * v[i];
    }
    static void add_scalar_mul(V& v, E k, const V& u) {  // lc.rs:14 ; ndarray.rs:23 `*y = *y + (k * *x)`
        for (size_t i = 0; i < v.size(); ++i) v[i] = v[i] + (k * u[i]);
    }
    static void add_assign_ref(V& v, const V& u) {  // lc.rs:16 ; ndarray.rs:27
        for (size_t i = 0; i < v.size(); ++i) v[i] = v[i] + u[i];
    }
    static void delta(V& v, const V& y) {  // lc.rs:18 ; ndarray.rs:31
        for (size_t i = 0; i < v.size(); ++i) v[i] = v[i] - y[i];
    }
    // lc.rs:20-35 and lc.rs:37-54: first term by scalar_multiply_to, the rest left to right by
    // add_scalar_mul; zero coefficients are NOT skipped. Returns false where the reference
    // panics / returns Err (empty input).
    static bool linear_combination(V& v, const V* const* v_arr, const double* k_arr, int n) {
        if (n <= 0) return false;
        scalar_multiply_to(*v_arr[0], scalar_of<E>::from_real(k_arr[0]), v);
        for (int j = 1; j < n; ++j) add_scalar_mul(v, scalar_of<E>::from_real(k_arr[j]), *v_arr[j]);
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// Coefficient tables (src/dat/mod.rs). Written as f64 division EXPRESSIONS like the reference so
// that the rounded constants are identical; includes the literal `-3544./2526.` of dat/mod.rs:19.
// ---------------------------------------------------------------------------------------------
static const double RK45_AC[36] = {  // dat/mod.rs:9-20  (c_i sits on the diagonal)
    0., 0., 0., 0., 0., 0.,
    1. / 4., 1. / 4., 0., 0., 0., 0.,
    3.0 / 32., 9.0 / 32., 3. / 8., 0., 0., 0.,
    1932. / 2197., -7200. / 2197., 7296. / 2197., 12. / 13., 0., 0.,
    439. / 216., -8., 3680. / 513., -845. / 4104., 1.0, 0.,
    -8. / 27., 2., -3544. / 2526., 1859. / 4104., -11. / 40., 1.0 / 2.0};
static const double RK45_B[6] = {16. / 135., 0., 6656. / 12825., 28561. / 56430., -9. / 50., 2. / 55.};  // :22-23
static const double RK45_BERR[6] = {25. / 216., 0., 1408. / 2565., 2197. / 4104., -1. / 5., 0.};        // :25-27
static const double C_GAUSS_LEGENDRE_4[2] = {0.21132486540518711775, 0.78867513459481288225};           // :4
static const double CFM_R2_J1_GL[2] = {0.5, 0.5};                                                       // :67-69
static const double CFM_R4_J2_GL[4] = {0.53867513459481288225, -0.038675134594812882255,
                                       -0.038675134594812882255, 0.53867513459481288225};              // :71-74

// ButcherTableu (src/base/rk.rs:22-78): s x s row-major `ac`, `b`, optional `b_err`.
struct Tableau {
    int s = 0;
    std::vector<double> ac, b, b_err;
    bool has_err = false;
};

// ---------------------------------------------------------------------------------------------
// Right-hand sides. The reference takes a user closure f(t, &x, &mut dx) (src/base/rk.rs:97); the
// only closures it ships are the linear-decay ones in its tests (src/impls/nalgebra.rs:54-58,
// 74-78, 93-96). The other families are this project's synthetic workloads (BASELINE.json configs);
// their operation order is DEFINED here and mirrored by the device functors.
// ---------------------------------------------------------------------------------------------
enum RhsKind { RHS_DIAG_LINEAR = 0, RHS_HARMONIC2D = 1, RHS_LORENZ63 = 2, RHS_VDP = 3, RHS_HEAT1D = 4 };

struct Rhs {
    int kind;
    const double* p;  // per-trajectory parameter vector (already gathered), see n_params()
    static int n_params(int kind, int d) {
        switch (kind) {
            case RHS_DIAG_LINEAR: return d;  // lambda_c : dx_c = lambda_c * x_c
            case RHS_HARMONIC2D: return 1;   // k       : dx = v ; dv = -(k*x)
            case RHS_LORENZ63: return 3;     // sigma, rho, beta
            case RHS_VDP: return 1;          // mu
            case RHS_HEAT1D: return 1;       // kappa   : du_j = kappa*((u_{j-1} + u_{j+1}) - 2*u_j), periodic
        }
        return 0;
    }
    void operator()(double /*t*/, const std::vector<double>& x, std::vector<double>& dx) const {
        switch (kind) {
            case RHS_DIAG_LINEAR:
                for (size_t c = 0; c < x.size(); ++c) dx[c] = p[c] * x[c];
                break;
            case RHS_HARMONIC2D:
                dx[0] = x[1];
                dx[1] = -(p[0] * x[0]);
                break;
            case RHS_LORENZ63:
                dx[0] = p[0] * (x[1] - x[0]);
                dx[1] = x[0] * (p[1] - x[2]) - x[1];
                dx[2] = x[0] * x[1] - p[2] * x[2];
                break;
            case RHS_VDP:
                dx[0] = x[1];
                dx[1] = (p[0] * (1.0 - x[0] * x[0])) * x[1] - x[0];
                break;
            case RHS_HEAT1D: {
                const size_t n = x.size();
                for (size_t j = 0; j < n; ++j) {
                    const double l = x[j == 0 ? n - 1 : j - 1], r = x[j + 1 == n ? 0 : j + 1];
                    dx[j] = p[0] * ((l + r) - 2.0 * x[j]);
                }
            } break;
        }
    }
    // complex flavour: only the diagonal-linear family (src/impls/nalgebra.rs:54-58: `y[1] = x[1] * -2.0`,
    // complex times real scales both parts).
    void operator()(double /*t*/, const std::vector<cplx>& x, std::vector<cplx>& dx) const {
        for (size_t c = 0; c < x.size(); ++c) dx[c] = {x[c].re * p[c], x[c].im * p[c]};
    }
};

// ---------------------------------------------------------------------------------------------
// rk_step (src/base/rk.rs:90-155) — one explicit RK step with optional embedded pair.
// ---------------------------------------------------------------------------------------------
template <class E>
static void rk_step(const Rhs& f, double t, const std::vector<E>& x0, std::vector<E>& xf, std::vector<E>* x_err,
                    double dt, const Tableau& tabl, std::vector<std::vector<E>>& K) {
    using L = LC<E>;
    using V = std::vector<E>;
    const E _dt = scalar_of<E>::from_real(dt);  // rk.rs:102
    const int s = (int)K.size() - 1;            // rk.rs:104-108 (K[s] is the unused work slot)
    f(t, x0, K[0]);                             // rk.rs:111
    const V* kp[32];
    for (int j = 0; j < s; ++j) kp[j] = &K[j];
    for (int i = 1; i < s; ++i) {               // rk.rs:118
        const double* ac = &tabl.ac[(size_t)i * s];
        const double ti = t + ac[i] * dt;       // rk.rs:119
        L::linear_combination(xf, kp, ac, i);   // rk.rs:121-122 (first i entries of row i)
        L::scale(xf, _dt);                      // rk.rs:123
        L::add_assign_ref(xf, x0);              // rk.rs:124
        f(ti, xf, K[i]);                        // rk.rs:127
    }
    L::linear_combination(xf, kp, tabl.b.data(), s);  // rk.rs:131
    L::scale(xf, _dt);                                // rk.rs:132
    L::add_assign_ref(xf, x0);                        // rk.rs:133
    if (tabl.has_err && x_err) {                      // rk.rs:136-151
        std::swap(*x_err, xf);                        // xe := X_b            rk.rs:142
        L::linear_combination(xf, kp, tabl.b_err.data(), s);  // rk.rs:143
        L::scale(xf, _dt);                            // rk.rs:145
        L::add_assign_ref(xf, x0);                    // rk.rs:146   xf := X_berr (propagated)
        L::delta(*x_err, xf);                         // rk.rs:147   x_err = X_b - X_berr
    }
}

// ---------------------------------------------------------------------------------------------
// Stepping state machine (src/base/ode.rs).
// ---------------------------------------------------------------------------------------------
enum StepKind { EV_STEP = 0, EV_CHKPT = 1, EV_REJECT = 2, EV_END = 3, EV_ERR = 4 };  // ODEStep, ode.rs:42-48
enum StateKind { ST_OK = 0, ST_DONE = 1, ST_ERR = 2 };                                // ODEState, ode.rs:34-38

// approx 0.5 `RelativeEq::relative_eq` for f64 (dependency, not under /root/reference; semantics restated):
// equal -> true; either infinite -> false; |a-b| <= epsilon -> true; else |a-b| <= max(|a|,|b|) * max_relative.
static inline bool relative_eq(double a, double b, double eps, double max_rel) {
    if (a == b) return true;
    if (std::isinf(a) || std::isinf(b)) return false;
    const double ad = std::fabs(a - b);
    if (ad <= eps) return true;
    const double aa = std::fabs(a), ab = std::fabs(b);
    const double largest = ab > aa ? ab : aa;
    return ad <= largest * max_rel;
}

// check_step (ode.rs:389-399): returns false for None.
static inline bool check_step(double t0, double tf, double dt, double* out) {
    const double rem = tf - t0;
    const double e = std::numeric_limits<double>::epsilon();
    if (relative_eq(rem, 0.0, e, e)) return false;
    *out = (rem < dt) ? rem : dt;
    return true;
}

// Rust f64::max/min (used through num_traits::real::Real, ode.rs:321-324) return the non-NaN operand;
// C fmax/fmin have the same NaN rule.
static inline double rmax(double a, double b) { return std::fmax(a, b); }
static inline double rmin(double a, double b) { return std::fmin(a, b); }

template <class E> struct ODEData {  // ode.rs:79-95, 139-206
    double t0, tf, t;
    std::vector<E> x0, x, next_x;
    std::vector<double> t_list;
    size_t tgt_t;
    double next_dt, h, prev_h;
    ODEData(double t0_, double tf_, const std::vector<E>& x0_, double h_)  // ode.rs:141-150
        : t0(t0_), tf(tf_), t(t0_), x0(x0_), x(x0_), next_x(x0_), t_list{t0_, tf_}, tgt_t(0), next_dt(h_), h(h_),
          prev_h(h_) {}
    int step_size_of(double dt_max, double* dt) const {  // ode.rs:165-176
        if (tgt_t >= t_list.size()) return EV_END;
        if (check_step(t, t_list[tgt_t], dt_max, dt)) return EV_STEP;
        if (tgt_t >= t_list.size() - 1) return EV_END;
        return EV_CHKPT;
    }
    void advance() {  // ode.rs:184-188
        std::swap(x, next_x);
        t += next_dt;
    }
    void checkpoint_update() {  // ode.rs:192-195
        tgt_t += 1;
        h = prev_h;
    }
    void reset_step_size(double h_) { h = h_, prev_h = h_; }        // ode.rs:197-200
    void update_step_size(double h_) { prev_h = h, h = h_; }        // ode.rs:202-205
};

struct ODEAdaptiveData {  // ode.rs:98-137 (the unused `dx: V` member is omitted)
    double atol = 1.0e-6, rtol = 1.0e-4, dx_norm = 0.0, alpha = 0.9, min_dt = 1.0e-6, max_dt = 1.0, pow_ = 1.0 / 3.0;
    explicit ODEAdaptiveData(double order) : pow_(1.0 / order) {}   // `order.recip()`, ode.rs:120
    double step_size_mul(double f) const { return alpha * std::pow(f, pow_); }  // ode.rs:133-135
};

enum NormKind { NORM_L2 = 0, NORM_LINF = 1, NORM_L1 = 2, NORM_HYPOT = 3 };
// Normed (ode.rs:9-11) is user-supplied for array states; scalars get |x| and complex scalars hypot
// (rk.rs:204-214). L2 = sqrt of the left-to-right sum of squares; for d = 1 it is sqrt(x*x) = |x| exactly.
static double norm_of(const std::vector<double>& v, int kind) {
    double acc = 0.0;
    switch (kind) {
        case NORM_L2:
            for (double e : v) acc = acc + e * e;
            return std::sqrt(acc);
        case NORM_LINF:
            for (double e : v) acc = rmax(acc, std::fabs(e));
            return acc;
        case NORM_L1:
            for (double e : v) acc = acc + std::fabs(e);
            return acc;
        case NORM_HYPOT:  // state interpreted as interleaved (re, im) pairs, 2-norm over the hypot moduli
            if (v.size() == 2) return std::hypot(v[0], v[1]);
            for (size_t i = 0; i + 1 < v.size(); i += 2) {
                const double m = std::hypot(v[i], v[i + 1]);
                acc = acc + m * m;
            }
            return std::sqrt(acc);
    }
    return 0.0;
}
static double norm_of(const std::vector<cplx>& v, int kind) {
    if (v.size() == 1) return std::hypot(v[0].re, v[0].im);  // rk.rs:209-214
    double acc = 0.0;
    for (auto e : v) {
        const double m = std::hypot(e.re, e.im);
        if (kind == NORM_LINF) acc = rmax(acc, m);
        else acc = acc + m * m;
    }
    return kind == NORM_LINF ? acc : std::sqrt(acc);
}

// RK45Solver (src/base/rk.rs:158-320) generalised to a runtime tableau (the reference hard-wires
// RK45_AC/B/BERR at rk.rs:250-254; everything else is tableau-agnostic).
template <class E> struct RKSolver {
    Rhs f;
    ODEData<E> dat;
    ODEAdaptiveData adat;
    bool has_x_err;  // Option<V> x_err: Some at construction (rk.rs:249), None after no_adaptive (rk.rs:233-237)
    std::vector<E> x_err;
    Tableau tabl;
    std::vector<std::vector<E>> K;
    int norm_kind = NORM_L2;
    uint64_t n_accept = 0, n_reject = 0, n_calls = 0;

    RKSolver(const Rhs& f_, const Tableau& tb, double t0, double tf, const std::vector<E>& x0, double h)
        : f(f_), dat(t0, tf, x0, h), adat(3.0), has_x_err(true), x_err(x0), tabl(tb),
          K((size_t)tb.s + 1, x0) {}  // rk.rs:248-262: order 3.0, alpha 0.9, K.resize(s+1, x0)

    void try_step(double dt) {  // rk.rs:287-293
        rk_step<E>(f, dat.t, dat.x, dat.next_x, has_x_err ? &x_err : nullptr, dt, tabl, K);
    }
    int handle_try_step(int ev, double dt) {  // ode.rs:242-246
        if (ev == EV_STEP) {
            dat.next_dt = dt;
            try_step(dt);
        }
        return ev;
    }
    int apply_step(int ev, bool adaptive) {  // ode.rs:402-428
        switch (ev) {
            case EV_STEP: dat.advance(); ++n_accept; return ST_OK;
            case EV_CHKPT: dat.checkpoint_update(); return ST_OK;
            case EV_REJECT: ++n_reject; return adaptive ? ST_OK : ST_ERR;
            case EV_END: dat.checkpoint_update(); return ST_DONE;
        }
        return ST_ERR;
    }
    int step(int* ev_out = nullptr) {  // ode.rs:249-253
        ++n_calls;
        double dt = 0.0;
        int ev = dat.step_size_of(dat.h, &dt);
        ev = handle_try_step(ev, dt);
        if (ev_out) *ev_out = ev;
        return apply_step(ev, false);
    }
    // ode.rs:311-341. Returns ST_ERR (-> the reference panics, ode.rs:312) when x_err is None.
    int step_adaptive(int* ev_out = nullptr) {
        ++n_calls;
        if (!has_x_err) return ST_ERR;
        double dt = 0.0;
        int ev = dat.step_size_of(dat.h, &dt);
        const double h = dat.h;                       // ode.rs:314 (nominal h, not the clipped dt)
        ev = handle_try_step(ev, dt);
        if (ev == EV_STEP) {
            const double dx_norm = norm_of(x_err, norm_kind);  // ode.rs:317, rk.rs:312-315
            adat.dx_norm = dx_norm;
            const double f_ = adat.rtol / adat.dx_norm;        // ode.rs:320 (atol is never used)
            const double fp_lim = rmin(rmax(adat.step_size_mul(f_), 0.3), 2.0);     // ode.rs:321-323
            const double new_h = rmin(rmax(fp_lim * h, adat.min_dt), adat.max_dt);  // ode.rs:324
            dat.update_step_size(new_h);                        // ode.rs:326
            if (f_ <= 1.0) ev = EV_REJECT;                      // ode.rs:328-330
        }
        if (ev_out) *ev_out = ev;
        return apply_step(ev, true);
    }
};

static Tableau make_tableau(const double* ac, const double* b, const double* b_err, int s) {
    Tableau t;
    t.s = s;
    t.ac.assign(ac, ac + (size_t)s * s);
    t.b.assign(b, b + s);
    t.has_err = b_err != nullptr;
    if (b_err) t.b_err.assign(b_err, b_err + s);
    return t;
}

struct SolveCfg {
    int32_t rhs_kind, d, s, has_berr;
    int32_t adaptive, no_adaptive, norm_kind, n_tlist;
    double t0, tf, h0;
    double rtol, atol, min_dt, max_dt, order, alpha;
    int64_t max_calls;
};

template <class E>
static void configure(RKSolver<E>& sv, const SolveCfg& c, const double* t_list) {
    if (c.no_adaptive) sv.has_x_err = false;       // rk.rs:233-237
    sv.norm_kind = c.norm_kind;
    sv.adat.rtol = c.rtol, sv.adat.atol = c.atol;  // with_tolerance, ode.rs:298-306
    sv.adat.min_dt = c.min_dt, sv.adat.max_dt = c.max_dt;
    sv.adat.pow_ = 1.0 / c.order, sv.adat.alpha = c.alpha;
    if (c.n_tlist > 0) sv.dat.t_list.assign(t_list, t_list + c.n_tlist);  // pub field, ode.rs:89
}

struct SolveOut {
    double t, h, prev_h, dx_norm;
    int64_t n_accept, n_reject, n_calls;
    int32_t state;
};

template <class E>
static void drive(RKSolver<E>& sv, const SolveCfg& c, SolveOut* out, double* trace, int64_t trace_cap) {
    int st = ST_OK;
    int64_t calls = 0;
    while (st == ST_OK && (c.max_calls <= 0 || calls < c.max_calls)) {  // `while let ODEState::Ok(_) = solver.step()`
        int ev = 0;
        st = c.adaptive ? sv.step_adaptive(&ev) : sv.step(&ev);
        if (trace && calls < trace_cap) {
            trace[calls * 4 + 0] = (double)ev;
            trace[calls * 4 + 1] = sv.dat.t;
            trace[calls * 4 + 2] = sv.dat.h;
            trace[calls * 4 + 3] = sv.adat.dx_norm;
        }
        ++calls;
    }
    out->t = sv.dat.t, out->h = sv.dat.h, out->prev_h = sv.dat.prev_h, out->dx_norm = sv.adat.dx_norm;
    out->n_accept = (int64_t)sv.n_accept, out->n_reject = (int64_t)sv.n_reject, out->n_calls = calls;
    out->state = st;
}

// ---------------------------------------------------------------------------------------------
// Exponential integrators (src/exp). The reference supplies the SCHEMES (nodes, weights, composition
// order); exp / map_exp / commutator / norm are user trait methods with no implementation in the crate
// (src/exp/mod.rs:11-54). The user side restated here is this project's "shared-basis dense split":
//   L = sum_m coef[m] * B_m   with M shared complex n x n basis matrices B_m,
// so an operator L is its coefficient vector (M complex numbers), LinearCombination on L acts on
// coefficients, exp(L) is lazy (U = L) and map_exp(U, x) evaluates the scaled Taylor series
//   x <- (sum_{k<=m} (L/sq)^k / k!)^sq x       (degree/sq chosen from theta = ||L||_1 bound).
// ---------------------------------------------------------------------------------------------
struct BasisSplit {
    int n = 0, M = 0;
    const cplx* B = nullptr;            // [M][n][n] row-major
    std::vector<double> b_norm1;        // induced 1-norm of each basis matrix
    std::vector<cplx> comm;             // structure tensor: [B_a, B_b] expanded on the basis, or empty
    int taylor_deg = 0;                 // 0 = choose automatically
    mutable std::vector<cplx> w, acc, term;

    void init(int n_, int M_, const cplx* B_) {
        n = n_, M = M_, B = B_;
        b_norm1.assign(M, 0.0);
        for (int m = 0; m < M; ++m) {
            double best = 0.0;
            for (int c = 0; c < n; ++c) {
                double col = 0.0;
                for (int r = 0; r < n; ++r) col += std::hypot(B[((size_t)m * n + r) * n + c].re, B[((size_t)m * n + r) * n + c].im);
                best = std::max(best, col);
            }
            b_norm1[m] = best;
        }
        w.resize(n), acc.resize(n), term.resize(n);
    }
    // theta bound: sum_m |coef_m| * ||B_m||_1
    double theta(const cplx* coef) const {
        double th = 0.0;
        for (int m = 0; m < M; ++m) th += std::hypot(coef[m].re, coef[m].im) * b_norm1[m];
        return th;
    }
    static void plan(double theta, int* sq, int* deg) {
        // sub-steps so that theta/sq <= 1, then the smallest degree with (theta/sq)^k/k! <= 2^-53 relative to 1
        int s = theta > 1.0 ? (int)std::ceil(theta) : 1;
        const double th = theta / s;
        double term = 1.0;
        int k = 0;
        while (k < 60) {
            ++k;
            term = term * th / k;
            if (term <= 1.1102230246251565e-16) break;
        }
        *sq = s, *deg = k;
    }
    // y = (sum_m coef[m] B_m) x
    void apply_L(const cplx* coef, const cplx* x, cplx* y) const {
        for (int r = 0; r < n; ++r) y[r] = {0.0, 0.0};
        for (int m = 0; m < M; ++m) {
            const cplx cm = coef[m];
            const cplx* Bm = B + (size_t)m * n * n;
            for (int r = 0; r < n; ++r) {
                cplx a = {0.0, 0.0};
                for (int c = 0; c < n; ++c) a = a + Bm[(size_t)r * n + c] * x[c];
                y[r] = y[r] + cm * a;
            }
        }
    }
    // map_exp(exp(L), x) (exp/mod.rs:23-25)
    void map_exp(const cplx* coef, const cplx* x, cplx* out) const {
        int sq, deg;
        plan(theta(coef), &sq, &deg);
        if (taylor_deg > 0) deg = taylor_deg;
        std::vector<cplx> cs(coef, coef + M);
        const double inv = 1.0 / sq;
        for (auto& c : cs) c = {c.re * inv, c.im * inv};
        std::vector<cplx> cur(x, x + n);
        for (int rep = 0; rep < sq; ++rep) {
            acc = cur;
            term = cur;
            for (int k = 1; k <= deg; ++k) {
                apply_L(cs.data(), term.data(), w.data());
                const double ik = 1.0 / k;
                for (int r = 0; r < n; ++r) {
                    term[r] = {w[r].re * ik, w[r].im * ik};
                    acc[r] = acc[r] + term[r];
                }
            }
            cur = acc;
        }
        for (int r = 0; r < n; ++r) out[r] = cur[r];
    }
};

// cfm_exp (src/exp/cfm.rs:20-40): k = a . m ; k *= dt ; x1 = map_exp(exp(k), x0). Operators are coefficient vectors.
static void cfm_exp(const BasisSplit& sp, const cplx* x0, cplx* x1, double dt, const std::vector<std::vector<cplx>>& m,
                    std::vector<cplx>& k, const double* a, int na) {
    const cplx a0 = {a[0], 0.0};
    for (size_t q = 0; q < k.size(); ++q) k[q] = a0 * m[0][q];                    // scalar_multiply_to, cfm.rs:31
    for (int i = 1; i < na; ++i) {                                                 // cfm.rs:33-36
        const cplx ai = {a[i], 0.0};
        for (size_t q = 0; q < k.size(); ++q) k[q] = k[q] + (ai * m[i][q]);
    }
    const cplx cdt = {dt, 0.0};
    for (auto& e : k) e = e * cdt;                                                // scale, cfm.rs:37
    sp.map_exp(k.data(), x0, x1);                                                 // cfm.rs:38-39
}

}  // namespace orc

// =============================================================================================
// C entry points (ctypes). All arrays are caller-owned.
// =============================================================================================
using namespace orc;

extern "C" {

void orc_builtin_tableau(int32_t which, double* ac, double* b, double* b_err, int32_t* s, int32_t* has_err) {
    // 0: RKF45_REF (dat/mod.rs:9-27, literal incl. the 2526 typo); 1: classical RK4; 2: Dormand-Prince 5(4)
    // in the reference layout with the 4th-order weights in `b` and the 5th-order ones in `b_err`, so the
    // "propagate X_berr" rule of rk.rs:142-146 yields local extrapolation.
    if (which == 0) {
        std::memcpy(ac, RK45_AC, sizeof RK45_AC), std::memcpy(b, RK45_B, sizeof RK45_B);
        std::memcpy(b_err, RK45_BERR, sizeof RK45_BERR);
        *s = 6, *has_err = 1;
    } else if (which == 1) {
        const double a[16] = {0., 0., 0., 0., 1. / 2., 1. / 2., 0., 0., 0., 1. / 2., 1. / 2., 0., 0., 0., 1., 1.};
        const double bb[4] = {1. / 6., 1. / 3., 1. / 3., 1. / 6.};
        std::memcpy(ac, a, sizeof a), std::memcpy(b, bb, sizeof bb);
        *s = 4, *has_err = 0;
    } else {
        const double a[49] = {
            0., 0., 0., 0., 0., 0., 0.,
            1. / 5., 1. / 5., 0., 0., 0., 0., 0.,
            3. / 40., 9. / 40., 3. / 10., 0., 0., 0., 0.,
            44. / 45., -56. / 15., 32. / 9., 4. / 5., 0., 0., 0.,
            19372. / 6561., -25360. / 2187., 64448. / 6561., -212. / 729., 8. / 9., 0., 0.,
            9017. / 3168., -355. / 33., 46732. / 5247., 49. / 176., -5103. / 18656., 1., 0.,
            35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 1.};
        const double b4[7] = {5179. / 57600., 0., 7571. / 16695., 393. / 640., -92097. / 339200., 187. / 2100., 1. / 40.};
        const double b5[7] = {35. / 384., 0., 500. / 1113., 125. / 192., -2187. / 6784., 11. / 84., 0.};
        std::memcpy(ac, a, sizeof a), std::memcpy(b, b4, sizeof b4), std::memcpy(b_err, b5, sizeof b5);
        *s = 7, *has_err = 1;
    }
}

int32_t orc_rhs_n_params(int32_t kind, int32_t d) { return Rhs::n_params(kind, d); }

// Solve ONE trajectory with the reference's driver loop. x: in = x0, out = current x. trace (optional):
// [calls][4] = (event, t, h, dx_norm) after each call.
void orc_rk_solve(const SolveCfg* cfg, const double* ac, const double* b, const double* b_err, const double* params,
                  const double* t_list, double* x, SolveOut* out, double* trace, int64_t trace_cap) {
    Tableau tb = make_tableau(ac, b, cfg->has_berr ? b_err : nullptr, cfg->s);
    Rhs f{cfg->rhs_kind, params};
    std::vector<double> x0(x, x + cfg->d);
    RKSolver<double> sv(f, tb, cfg->t0, cfg->tf, x0, cfg->h0);
    configure(sv, *cfg, t_list);
    drive(sv, *cfg, out, trace, trace_cap);
    std::copy(sv.dat.x.begin(), sv.dat.x.end(), x);
}

// Complex-state flavour (RK45ComplexSolver, rk.rs:218): x holds d interleaved (re, im) pairs.
void orc_rk_solve_c64(const SolveCfg* cfg, const double* ac, const double* b, const double* b_err, const double* params,
                      const double* t_list, double* x, SolveOut* out) {
    Tableau tb = make_tableau(ac, b, cfg->has_berr ? b_err : nullptr, cfg->s);
    Rhs f{RHS_DIAG_LINEAR, params};
    std::vector<cplx> x0(cfg->d);
    for (int c = 0; c < cfg->d; ++c) x0[c] = {x[2 * c], x[2 * c + 1]};
    RKSolver<cplx> sv(f, tb, cfg->t0, cfg->tf, x0, cfg->h0);
    configure(sv, *cfg, t_list);
    drive(sv, *cfg, out, nullptr, 0);
    for (int c = 0; c < cfg->d; ++c) x[2 * c] = sv.dat.x[c].re, x[2 * c + 1] = sv.dat.x[c].im;
}

// Ensemble of N independent solvers in the reference's own shape: one heap-allocated solver object per
// trajectory, un-fused LinearCombination passes, contiguous trajectory ranges per worker thread.
// x: AoS [N][d] in/out. params: AoS [N][n_params]. Per-trajectory outputs are optional (nullable).
// h0_arr (nullable) overrides cfg->h0 per trajectory.
void orc_rk_ensemble(const SolveCfg* cfg, const double* ac, const double* b, const double* b_err, const double* params,
                     int32_t n_params, const double* t_list, int64_t N, double* x, const double* h0_arr, double* t_out,
                     double* h_out, int64_t* acc_out, int64_t* rej_out, int32_t* state_out, int32_t n_threads) {
    Tableau tb = make_tableau(ac, b, cfg->has_berr ? b_err : nullptr, cfg->s);
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            Rhs f{cfg->rhs_kind, params + (size_t)i * n_params};
            std::vector<double> x0(x + (size_t)i * cfg->d, x + (size_t)(i + 1) * cfg->d);
            RKSolver<double> sv(f, tb, cfg->t0, cfg->tf, x0, h0_arr ? h0_arr[i] : cfg->h0);
            configure(sv, *cfg, t_list);
            SolveOut o;
            drive(sv, *cfg, &o, nullptr, 0);
            std::copy(sv.dat.x.begin(), sv.dat.x.end(), x + (size_t)i * cfg->d);
            if (t_out) t_out[i] = o.t;
            if (h_out) h_out[i] = o.h;
            if (acc_out) acc_out[i] = o.n_accept;
            if (rej_out) rej_out[i] = o.n_reject;
            if (state_out) state_out[i] = o.state;
        }
    };
    if (n_threads == 1) {
        work(0, N);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t chunk = (N + n_threads - 1) / n_threads;
    for (int w = 0; w < n_threads; ++w) {
        const int64_t lo = w * chunk, hi = std::min<int64_t>(N, lo + chunk);
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
}

// One bare rk_step (for kernel-level parity tests): x0 -> xf, x_err (nullable), K stages out (nullable, [s][d]).
void orc_rk_step(int32_t rhs_kind, int32_t d, const double* params, const double* ac, const double* b, const double* b_err,
                 int32_t s, double t, double dt, const double* x0, double* xf, double* x_err, double* K_out) {
    Tableau tb = make_tableau(ac, b, b_err, s);
    Rhs f{rhs_kind, params};
    std::vector<double> v0(x0, x0 + d), vf(v0), ve(v0);
    std::vector<std::vector<double>> K((size_t)s + 1, v0);
    rk_step<double>(f, t, v0, vf, x_err ? &ve : nullptr, dt, tb, K);
    std::copy(vf.begin(), vf.end(), xf);
    if (x_err) std::copy(ve.begin(), ve.end(), x_err);
    if (K_out)
        for (int j = 0; j < s; ++j) std::copy(K[j].begin(), K[j].end(), K_out + (size_t)j * d);
}

// LinearCombination primitives on flat vectors (lc.rs:10-18), for LC-kernel parity tests.
void orc_lc_scale(double* v, double k, int64_t n) { for (int64_t i = 0; i < n; ++i) v[i] = v[i] * k; }
void orc_lc_scalar_multiply_to(const double* v, double k, double* t, int64_t n) { for (int64_t i = 0; i < n; ++i) t[i] = k * v[i]; }
void orc_lc_add_scalar_mul(double* v, double k, const double* u, int64_t n) { for (int64_t i = 0; i < n; ++i) v[i] = v[i] + (k * u[i]); }
void orc_lc_add_assign_ref(double* v, const double* u, int64_t n) { for (int64_t i = 0; i < n; ++i) v[i] = v[i] + u[i]; }
void orc_lc_delta(double* v, const double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) v[i] = v[i] - y[i]; }
void orc_lc_linear_combination(double* v, const double* const* v_arr, const double* k_arr, int32_t m, int64_t n) {
    orc_lc_scalar_multiply_to(v_arr[0], k_arr[0], v, n);
    for (int j = 1; j < m; ++j) orc_lc_add_scalar_mul(v, k_arr[j], v_arr[j], n);
}

// The same primitives with A = num_complex::Complex<f64> (src/impls/ndarray.rs:8-33 is generic over the element type), vectors of nz
// interleaved (re, im) pairs. num-complex 0.4 (third-party, not under /root/reference; restated from its published source):
// Mul: (a.re b.re - a.im b.im, a.re b.im + a.im b.re); MulAssign: re = re k.re - im k.im, im = im k.re + re k.im; Add componentwise.
static inline void zmul(double ar, double ai, double br, double bi, double* re, double* im) { *re = ar * br - ai * bi, *im = ar * bi + ai * br; }
void orc_lcz_scale(double* v, double kr, double ki, int64_t nz) {
    for (int64_t i = 0; i < nz; ++i) {  // *self *= k (MulAssign)
        const double a = v[2 * i], re = v[2 * i] * kr - v[2 * i + 1] * ki, im = v[2 * i + 1] * kr + a * ki;
        v[2 * i] = re, v[2 * i + 1] = im;
    }
}
void orc_lcz_scalar_multiply_to(const double* v, double kr, double ki, double* t, int64_t nz) {
    for (int64_t i = 0; i < nz; ++i) zmul(kr, ki, v[2 * i], v[2 * i + 1], &t[2 * i], &t[2 * i + 1]);  // *t = k * s
}
void orc_lcz_add_scalar_mul(double* v, double kr, double ki, const double* u, int64_t nz) {
    for (int64_t i = 0; i < nz; ++i) {  // *y = *y + (k * *x)
        double pr, pi;
        zmul(kr, ki, u[2 * i], u[2 * i + 1], &pr, &pi);
        v[2 * i] = v[2 * i] + pr, v[2 * i + 1] = v[2 * i + 1] + pi;
    }
}
void orc_lcz_linear_combination(double* v, const double* const* v_arr, const double* k_arr, int32_t m, int64_t nz) {
    orc_lcz_scalar_multiply_to(v_arr[0], k_arr[0], k_arr[1], v, nz);
    for (int j = 1; j < m; ++j) orc_lcz_add_scalar_mul(v, k_arr[2 * j], k_arr[2 * j + 1], v_arr[j], nz);
}

// ---- exponential integrators on the shared-basis dense split ---------------------------------------
// Generator family: L_i(t) = -i * (H_0 + sum_{m>=1} g_m(t; p_i) H_m), g_m(t) = amp_{i,m} * cos(omega_{i,m} t + phase_{i,m}).
// Basis handed in as B_m = -i * H_m so coefficients are real; gp: [N][M-1][3] = (amp, omega, phase).
struct ExpCfg {
    int32_t n, M, scheme /*0 midpoint, 1 cfm4, 2 magnus42, 3 cfm_general with the tables below*/, adaptive, no_adaptive, taylor_deg;
    double t0, tf, h0, rtol, min_dt, max_dt, order, alpha;
    int64_t max_calls;
    // scheme 3: cfm_general (cfm.rs:43-100) with caller-supplied nodes c[n_nodes], alpha[n_rows][n_nodes] and the optional
    // lower-order alph_err[n_rows_err][n_nodes]
    int32_t n_nodes, n_rows, n_rows_err, literal_norm /* magnus.rs:274-276 as written: norm of adaptive_dat.dx = x0, never updated */;
    const double* c;
    const double* alpha_tab;
    const double* alpha_err_tab;
};

static void gen_coef(const ExpCfg& c, const double* gp_i, double t, std::vector<cplx>& coef) {
    coef.assign((size_t)c.M, cplx{0.0, 0.0});
    coef[0] = {1.0, 0.0};
    for (int m = 1; m < c.M && m < c.M; ++m) {
        const double* g = gp_i + (size_t)(m - 1) * 3;
        coef[m] = {g[0] * std::cos(g[1] * t + g[2]), 0.0};
    }
}

// psi: AoS [N][n] interleaved complex, in/out. comm_map: for magnus, commutator structure: M_total includes the
// commutator basis matrices; [B_a,B_b] = sum_c cs[a][b][c] B_c given as real tensor [M][M][M] (nullable for non-magnus).
void orc_exp_ensemble(const ExpCfg* cfg, const double* basis /*[M][n][n] complex interleaved*/, const double* gp,
                      int32_t M_gen /*number of basis matrices the generator uses (rest are commutator slots)*/,
                      const double* cs, int64_t N, double* psi, double* t_out, double* h_out, int64_t* acc_out,
                      int64_t* rej_out, int32_t n_threads) {
    const int n = cfg->n, M = cfg->M;
    auto work = [&](int64_t lo, int64_t hi) {
        BasisSplit sp;
        sp.init(n, M, reinterpret_cast<const cplx*>(basis));
        sp.taylor_deg = cfg->taylor_deg;
        ExpCfg gc = *cfg;
        gc.M = M_gen;
        for (int64_t i = lo; i < hi; ++i) {
            const double* gpi = gp + (size_t)i * (M_gen - 1) * 3;
            std::vector<cplx> x0(n);
            for (int r = 0; r < n; ++r) x0[r] = {psi[((size_t)i * n + r) * 2], psi[((size_t)i * n + r) * 2 + 1]};
            ODEData<cplx> dat(cfg->t0, cfg->tf, x0, cfg->h0);
            ODEAdaptiveData ad(cfg->order);
            ad.alpha = cfg->alpha, ad.rtol = cfg->rtol, ad.min_dt = cfg->min_dt, ad.max_dt = cfg->max_dt;
            std::vector<cplx> dx(x0), k((size_t)M), tmp(n), tmp2(n);
            std::vector<std::vector<cplx>> va;
            int64_t acc = 0, rej = 0, calls = 0;
            const bool want_err = !cfg->no_adaptive;
            auto pad = [&](std::vector<cplx>& v) { v.resize((size_t)M, cplx{0.0, 0.0}); };
            auto try_step = [&](double dt) {
                const double t = dat.t;
                if (cfg->scheme == 0) {  // midpoint, magnus.rs:10-26
                    const double t_mid = t + dt * 0.5;
                    std::vector<cplx> l;
                    gen_coef(gc, gpi, t_mid, l), pad(l);
                    const cplx cdt = {dt, 0.0};
                    for (auto& e : l) e = e * cdt;
                    sp.map_exp(l.data(), dat.x.data(), dat.next_x.data());
                } else if (cfg->scheme == 1) {  // cfm_general with CFM4 tables, cfm.rs:43-100, 131-154
                    va.resize(2);
                    for (int q = 0; q < 2; ++q) gen_coef(gc, gpi, t + C_GAUSS_LEGENDRE_4[q] * dt, va[q]), pad(va[q]);
                    cfm_exp(sp, dat.x.data(), tmp.data(), dt, va, k, &CFM_R4_J2_GL[0], 2);        // row 0
                    cfm_exp(sp, tmp.data(), dat.next_x.data(), dt, va, k, &CFM_R4_J2_GL[2], 2);   // row 1
                    if (want_err) {                                                                // cfm.rs:83-97
                        cfm_exp(sp, dat.x.data(), dx.data(), dt, va, k, CFM_R2_J1_GL, 2);
                        for (int r = 0; r < n; ++r) dx[r] = dx[r] - dat.next_x[r];
                    }
                } else if (cfg->scheme == 3) {  // cfm_general with runtime tables, cfm.rs:43-100
                    const int kn = cfg->n_nodes;
                    va.resize((size_t)kn);
                    for (int q = 0; q < kn; ++q) gen_coef(gc, gpi, t + cfg->c[q] * dt, va[q]), pad(va[q]);   // :70-72
                    cfm_exp(sp, dat.x.data(), tmp.data(), dt, va, k, cfg->alpha_tab, kn);                      // :74-75
                    for (int i = 1; i < cfg->n_rows; ++i) {                                                    // :76-80
                        cfm_exp(sp, tmp.data(), tmp2.data(), dt, va, k, cfg->alpha_tab + (size_t)i * kn, kn);
                        std::swap(tmp, tmp2);
                    }
                    dat.next_x = tmp;                                                                          // :81
                    if (want_err && cfg->alpha_err_tab) {                                                      // :83-97
                        cfm_exp(sp, dat.x.data(), tmp.data(), dt, va, k, cfg->alpha_err_tab, kn);
                        for (int i = 1; i < cfg->n_rows_err; ++i) {
                            cfm_exp(sp, tmp.data(), tmp2.data(), dt, va, k, cfg->alpha_err_tab + (size_t)i * kn, kn);
                            std::swap(tmp, tmp2);
                        }
                        for (int r = 0; r < n; ++r) dx[r] = tmp[r] - dat.next_x[r];
                    }
                } else {  // magnus_42, magnus.rs:28-83
                    const double c_mid = 0.288675134594812882254574390251;
                    const double b1 = dt * 0.5;
                    const double b2 = dt * dt * -0.144337567297406441127287195125;
                    const double mid_t = t + b1;
                    std::vector<cplx> l0, l1, w2((size_t)M, cplx{0.0, 0.0}), w1, w;
                    gen_coef(gc, gpi, mid_t - c_mid * dt, l0), pad(l0);
                    gen_coef(gc, gpi, mid_t + c_mid * dt, l1), pad(l1);
                    // commutator(l0, l1) expanded on the basis through the structure tensor (magnus.rs:55)
                    for (int a = 0; a < M; ++a)
                        for (int bq = 0; bq < M; ++bq) {
                            const cplx ab = l0[a] * l1[bq];
                            for (int cq = 0; cq < M; ++cq) {
                                const double sc = cs[((size_t)a * M + bq) * M + cq];
                                if (sc != 0.0) w2[cq] = w2[cq] + cplx{ab.re * sc, ab.im * sc};
                            }
                        }
                    const cplx cb2 = {b2, 0.0}, cb1 = {b1, 0.0};
                    for (auto& e : w2) e = e * cb2;                                   // magnus.rs:56
                    w1 = l0;                                                           // magnus.rs:59
                    for (int q = 0; q < M; ++q) w1[q] = w1[q] + l1[q];                 // :60
                    for (auto& e : w1) e = e * cb1;                                    // :61
                    w = w1;
                    for (int q = 0; q < M; ++q) w[q] = w[q] + w2[q];                   // :65-66
                    sp.map_exp(w.data(), dat.x.data(), dat.next_x.data());            // :72,75
                    if (want_err) {                                                    // :76-79
                        sp.map_exp(w1.data(), dat.x.data(), dx.data());
                        for (int r = 0; r < n; ++r) dx[r] = dx[r] - dat.next_x[r];
                    }
                }
            };
            int st = ST_OK;
            while (st == ST_OK && (cfg->max_calls <= 0 || calls < cfg->max_calls)) {
                ++calls;
                double dt = 0.0;
                int ev = dat.step_size_of(dat.h, &dt);
                const double h = dat.h;
                if (ev == EV_STEP) {
                    dat.next_dt = dt;
                    try_step(dt);
                    if (cfg->adaptive) {  // ode.rs:311-334 with norm = 2-norm of dx
                        const std::vector<cplx>& nv = (cfg->literal_norm && cfg->scheme == 2) ? x0 : dx;  // magnus.rs:274-276 reads adaptive_dat.dx (= x0)
                        double nn = 0.0;
                        for (int r = 0; r < n; ++r) nn = nn + (nv[r].re * nv[r].re + nv[r].im * nv[r].im);
                        ad.dx_norm = std::sqrt(nn);
                        const double f_ = ad.rtol / ad.dx_norm;
                        const double fp = rmin(rmax(ad.step_size_mul(f_), 0.3), 2.0);
                        dat.update_step_size(rmin(rmax(fp * h, ad.min_dt), ad.max_dt));
                        if (f_ <= 1.0) ev = EV_REJECT;
                    }
                }
                switch (ev) {
                    case EV_STEP: dat.advance(), ++acc; break;
                    case EV_CHKPT: dat.checkpoint_update(); break;
                    case EV_REJECT: ++rej; break;
                    case EV_END: dat.checkpoint_update(), st = ST_DONE; break;
                }
            }
            for (int r = 0; r < n; ++r) psi[((size_t)i * n + r) * 2] = dat.x[r].re, psi[((size_t)i * n + r) * 2 + 1] = dat.x[r].im;
            if (t_out) t_out[i] = dat.t;
            if (h_out) h_out[i] = dat.h;
            if (acc_out) acc_out[i] = acc;
            if (rej_out) rej_out[i] = rej;
        }
    };
    if (n_threads <= 1) {
        work(0, N);
        return;
    }
    std::vector<std::thread> pool;
    const int64_t chunk = (N + n_threads - 1) / n_threads;
    for (int w = 0; w < n_threads; ++w) {
        const int64_t lo = w * chunk, hi = std::min<int64_t>(N, lo + chunk);
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    for (auto& th : pool) th.join();
}

// Taylor plan used by map_exp, exposed so tests can pin degree/sub-step choices.
void orc_exp_plan(double theta, int32_t* sq, int32_t* deg) {
    int a, b;
    BasisSplit::plan(theta, &a, &b);
    *sq = a, *deg = b;
}

int32_t orc_hardware_threads(void) { return (int32_t)std::thread::hardware_concurrency(); }

}  // extern "C"
