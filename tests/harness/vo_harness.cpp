// vo_harness.cpp — a plain C++ caller of the C ABI (include/vecode_b200.h), the "C++ harness" of SURVEY.md §8(b): what a compiled
// host (the reference is a compiled Rust crate) does with libvecode_b200.so, with no Python and no torch anywhere.
// Test infrastructure. It restates the reference's own three tests (src/impls/nalgebra.rs:52-107: `while let ODEState::Ok(_) =
// solver.step() {}` on y' = (-y0, -2 y1) and the adaptive scalar y' = -y with tolerance 1e-10) through the ABI, and a fixed-step
// RK4 sweep over a Lorenz-63 ensemble that is compared BIT FOR BIT with a scalar restatement of rk_step (src/base/rk.rs:90-155,
// LinearCombination order of src/lc.rs:20-54, zero coefficients kept) written out below.
//   exit 0: everything matched;  exit 3: no CUDA device (the library has no CPU fallback and says so);  exit 1: mismatch.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vecode_b200.h"

#define CHECK(ctx, call)                                                                      \
    do {                                                                                      \
        const int32_t rc_ = (call);                                                           \
        if (rc_ != 0) {                                                                       \
            std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vo_last_error(ctx));     \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

// rk_step for RK4 on Lorenz-63 in the reference's operation order (tableau layout of rk.rs:31-42: c_i on the diagonal).
static void lorenz(const double* x, double* dx, double s, double r, double b) {
    dx[0] = s * (x[1] - x[0]);
    dx[1] = x[0] * (r - x[2]) - x[1];
    dx[2] = x[0] * x[1] - b * x[2];
}
static void rk4_step_ref(double* x, double dt, double s, double r, double b) {
    static const double ac[16] = {0, 0, 0, 0, 0.5, 0.5, 0, 0, 0, 0.5, 0.5, 0, 0, 0, 1, 1};
    static const double bw[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    double K[4][3], xs[3], v[3];
    lorenz(x, K[0], s, r, b);                                   // rk.rs:111
    for (int i = 1; i < 4; ++i) {
        for (int c = 0; c < 3; ++c) v[c] = ac[i * 4] * K[0][c];  // lc.rs:27: v = k0 * v0
        for (int j = 1; j < i; ++j)
            for (int c = 0; c < 3; ++c) v[c] = v[c] + ac[i * 4 + j] * K[j][c];  // lc.rs:29-31 (zeros kept)
        for (int c = 0; c < 3; ++c) xs[c] = v[c] * dt + x[c];    // rk.rs:123-124
        lorenz(xs, K[i], s, r, b);                               // rk.rs:127
    }
    for (int c = 0; c < 3; ++c) v[c] = bw[0] * K[0][c];
    for (int j = 1; j < 4; ++j)
        for (int c = 0; c < 3; ++c) v[c] = v[c] + bw[j] * K[j][c];
    for (int c = 0; c < 3; ++c) x[c] = v[c] * dt + x[c];          // rk.rs:131-133
}

int main() {
    vo_ctx ctx = nullptr;
    if (vo_ctx_create(0, nullptr, &ctx) != 0) {
        std::fprintf(stderr, "vo_ctx_create: %s\n", vo_last_error(nullptr));
        return 3;
    }
    CHECK(ctx, vo_ctx_set_arith(ctx, VO_ARITH_STRICT));

    // ---- test_rk45_2 (nalgebra.rs:72-88): y' = (-y0, -2 y1), RK45Solver::new(g, 0, 2, x0, 1e-4), step() until Done -----------
    {
        vo_rhs rhs = nullptr;
        vo_ens x0 = nullptr;
        vo_solver s = nullptr;
        CHECK(ctx, vo_rhs_create(ctx, VO_RHS_DIAG_LINEAR, 2, &rhs));
        CHECK(ctx, vo_rhs_set_param(rhs, 0, -1.0));
        CHECK(ctx, vo_rhs_set_param(rhs, 1, -2.0));
        CHECK(ctx, vo_ens_create(ctx, 2, 1, &x0));
        const double one[2] = {1.0, 1.0};
        CHECK(ctx, vo_ens_upload(x0, one, VO_LAYOUT_AOS));
        CHECK(ctx, vo_rk45_create(ctx, rhs, 0.0, 2.0, x0, 1.0e-4, &s));
        vo_step_result res;
        int64_t calls = 0, steps = 0;
        do {  // while let ODEState::Ok(_) = solver.step() {}
            CHECK(ctx, vo_step(s, &res));
            ++calls, steps += res.n_step;
        } while (res.state == VO_STATE_OK);
        double t0 = 0, t1 = 0, xf[2];
        vo_ens cur = nullptr;
        CHECK(ctx, vo_current(s, &t0, &t1, &cur));
        CHECK(ctx, vo_ens_download(cur, xf, VO_LAYOUT_AOS));
        std::printf("test_rk45_2: %lld calls, %lld steps, tf = %.17g, xf = (%.17g, %.17g)\n", (long long)calls, (long long)steps, t1, xf[0], xf[1]);
        if (res.state != VO_STATE_DONE || std::fabs(t1 - 2.0) > 1e-12 || std::fabs(xf[0] - std::exp(-2.0)) > 1e-9 || std::fabs(xf[1] - std::exp(-4.0)) > 1e-9) return 1;
        vo_solver_destroy(s), vo_ens_destroy(x0), vo_rhs_destroy(rhs);
    }
    // ---- test_rk45_f64 (nalgebra.rs:90-105): y' = -y, with_tolerance(1e-10, 1e-10), step_adaptive() until Done ------------------
    {
        vo_rhs rhs = nullptr;
        vo_ens x0 = nullptr;
        vo_solver s = nullptr;
        CHECK(ctx, vo_rhs_create(ctx, VO_RHS_DIAG_LINEAR, 1, &rhs));
        CHECK(ctx, vo_rhs_set_param(rhs, 0, -1.0));
        CHECK(ctx, vo_ens_create(ctx, 1, 1, &x0));
        const double one = 1.0;
        CHECK(ctx, vo_ens_upload(x0, &one, VO_LAYOUT_AOS));
        CHECK(ctx, vo_rk45_create(ctx, rhs, 0.0, 2.0, x0, 1.0e-4, &s));
        CHECK(ctx, vo_solver_with_tolerance(s, 1.0e-10, 1.0e-10));
        vo_step_result res;
        int64_t acc = 0, rej = 0;
        do {
            CHECK(ctx, vo_step_adaptive(s, &res));
            acc += res.n_step, rej += res.n_reject;
        } while (res.state == VO_STATE_OK);
        double xf = 0, t1 = 0;
        vo_ens cur = nullptr;
        CHECK(ctx, vo_current(s, nullptr, &t1, &cur));
        CHECK(ctx, vo_ens_download(cur, &xf, VO_LAYOUT_AOS));
        std::printf("test_rk45_f64: accepted %lld, rejected %lld, tf = %.17g, xf = %.17g (e^-2 = %.17g)\n", (long long)acc, (long long)rej, t1, xf, std::exp(-2.0));
        if (res.state != VO_STATE_DONE || std::fabs(xf - std::exp(-2.0)) > 1e-8 || acc < 100) return 1;
        vo_solver_destroy(s), vo_ens_destroy(x0), vo_rhs_destroy(rhs);
    }
    // ---- config 2, small: fixed-step RK4 over a Lorenz-63 ensemble, bit for bit against the scalar restatement above -----------
    {
        const int64_t N = 3000;  // not a multiple of any tile
        const int n_steps = 40;
        const double dt = 1.0e-3, sg = 10.0, rho = 28.0, beta = 8.0 / 3.0;
        std::vector<double> x((size_t)N * 3), ref, got((size_t)N * 3);
        for (int64_t i = 0; i < N; ++i) x[3 * i] = 1.0 + 1e-3 * std::sin(0.1 * i), x[3 * i + 1] = 1.0 - 1e-3 * std::cos(0.3 * i), x[3 * i + 2] = 1.0 + 1e-6 * i;
        ref = x;
        vo_rhs rhs = nullptr;
        vo_tableau tab = nullptr;
        vo_ens x0 = nullptr;
        vo_solver s = nullptr;
        CHECK(ctx, vo_rhs_create(ctx, VO_RHS_LORENZ63, 3, &rhs));
        CHECK(ctx, vo_rhs_set_param(rhs, 0, sg));
        CHECK(ctx, vo_rhs_set_param(rhs, 1, rho));
        CHECK(ctx, vo_rhs_set_param(rhs, 2, beta));
        CHECK(ctx, vo_tableau_builtin(VO_TABLEAU_RK4, &tab));
        CHECK(ctx, vo_ens_create(ctx, 3, N, &x0));
        CHECK(ctx, vo_ens_upload(x0, x.data(), VO_LAYOUT_AOS));
        CHECK(ctx, vo_rk_create(ctx, tab, rhs, 0.0, 1.0e9, x0, dt, &s));
        vo_step_result res;
        CHECK(ctx, vo_run(s, 0, n_steps + 1, &res));  // the first call is the Chkpt at t0 (ode.rs:144-145)
        vo_ens cur = nullptr;
        CHECK(ctx, vo_current(s, nullptr, nullptr, &cur));
        CHECK(ctx, vo_ens_download(cur, got.data(), VO_LAYOUT_AOS));
        for (int64_t i = 0; i < N; ++i)
            for (int k = 0; k < n_steps; ++k) rk4_step_ref(&ref[3 * i], dt, sg, rho, beta);
        const bool same = std::memcmp(ref.data(), got.data(), sizeof(double) * ref.size()) == 0;
        std::printf("lorenz rk4: %lld trajectories x %d steps, %lld step events, kernels launched so far %lld, bit-exact vs restatement: %s\n", (long long)N, n_steps,
                    (long long)res.n_step, (long long)vo_ctx_launch_count(ctx), same ? "yes" : "NO");
        if (!same || res.n_step != N * n_steps) return 1;
        vo_solver_destroy(s), vo_ens_destroy(x0), vo_tableau_destroy(tab), vo_rhs_destroy(rhs);
    }
    // ---- the multi-GPU entry points from one process (vo_group_create_local over the GPUs of the box, at least this one): scatter
    // ---- from a host array, `while let Ok(_) = step()` on every shard, gather back to the host; must equal the single-ctx solve
    {
        int n_gpu = 1;
        if (std::getenv("VO_HARNESS_GPUS")) n_gpu = std::atoi(std::getenv("VO_HARNESS_GPUS"));
        std::vector<vo_ctx> ctxs((size_t)n_gpu);
        ctxs[0] = ctx;
        for (int g = 1; g < n_gpu; ++g) {
            CHECK(nullptr, vo_ctx_create(g, nullptr, &ctxs[(size_t)g]));
            CHECK(ctxs[(size_t)g], vo_ctx_set_arith(ctxs[(size_t)g], VO_ARITH_STRICT));
        }
        vo_group grp = nullptr;
        CHECK(ctx, vo_group_create_local(ctxs.data(), n_gpu, &grp));
        const int64_t N = 5001;
        const double dt = 1.0e-3;
        std::vector<double> x((size_t)N * 3), whole((size_t)N * 3), got((size_t)N * 3);
        for (int64_t i = 0; i < N; ++i) x[3 * i] = 1.0 + 1e-4 * i, x[3 * i + 1] = 1.0, x[3 * i + 2] = 1.0 - 1e-5 * i;
        vo_tableau tab = nullptr;
        CHECK(ctx, vo_tableau_builtin(VO_TABLEAU_RK4, &tab));
        std::vector<vo_ens> shard((size_t)n_gpu), fin((size_t)n_gpu);
        std::vector<vo_rhs> rhs((size_t)n_gpu);
        std::vector<vo_solver> sol((size_t)n_gpu);
        for (int g = 0; g < n_gpu; ++g) {
            int64_t lo = 0, hi = 0;
            CHECK(ctx, vo_group_shard_range(N, g, n_gpu, &lo, &hi));
            CHECK(ctxs[(size_t)g], vo_ens_create(ctxs[(size_t)g], 3, hi - lo, &shard[(size_t)g]));
        }
        CHECK(ctx, vo_group_scatter(grp, x.data(), VO_LAYOUT_AOS, 3, N, 0, shard.data()));
        for (int g = 0; g < n_gpu; ++g) {
            vo_ctx c = ctxs[(size_t)g];
            CHECK(c, vo_rhs_create(c, VO_RHS_LORENZ63, 3, &rhs[(size_t)g]));
            CHECK(c, vo_rhs_set_param(rhs[(size_t)g], 0, 10.0));
            CHECK(c, vo_rhs_set_param(rhs[(size_t)g], 1, 28.0));
            CHECK(c, vo_rhs_set_param(rhs[(size_t)g], 2, 8.0 / 3.0));
            CHECK(c, vo_rk_create(c, tab, rhs[(size_t)g], 0.0, 0.02, shard[(size_t)g], dt, &sol[(size_t)g]));
        }
        vo_group_stats tot;
        CHECK(ctx, vo_group_run(grp, sol.data(), 0, 0, &tot));
        for (int g = 0; g < n_gpu; ++g) CHECK(ctxs[(size_t)g], vo_current(sol[(size_t)g], nullptr, nullptr, &fin[(size_t)g]));
        CHECK(ctx, vo_group_gather(grp, fin.data(), N, 0, got.data(), VO_LAYOUT_AOS));
        // the same ensemble on one ctx
        vo_ens x0 = nullptr;
        vo_solver one = nullptr;
        vo_rhs r1 = nullptr;
        CHECK(ctx, vo_rhs_create(ctx, VO_RHS_LORENZ63, 3, &r1));
        CHECK(ctx, vo_rhs_set_param(r1, 0, 10.0));
        CHECK(ctx, vo_rhs_set_param(r1, 1, 28.0));
        CHECK(ctx, vo_rhs_set_param(r1, 2, 8.0 / 3.0));
        CHECK(ctx, vo_ens_create(ctx, 3, N, &x0));
        CHECK(ctx, vo_ens_upload(x0, x.data(), VO_LAYOUT_AOS));
        CHECK(ctx, vo_rk_create(ctx, tab, r1, 0.0, 0.02, x0, dt, &one));
        vo_step_result res;
        CHECK(ctx, vo_run(one, 0, 0, &res));
        vo_ens cur = nullptr;
        CHECK(ctx, vo_current(one, nullptr, nullptr, &cur));
        CHECK(ctx, vo_ens_download(cur, whole.data(), VO_LAYOUT_AOS));
        const bool same = std::memcmp(whole.data(), got.data(), sizeof(double) * whole.size()) == 0;
        std::printf("group of %d GPU(s): %lld trajectories done, accepted %lld, gathered ensemble == single-ctx solve: %s\n", vo_group_world(grp), (long long)tot.n_done,
                    (long long)tot.accepted, same ? "yes" : "NO");
        if (!same || tot.n_done != N || tot.t_max != 0.02) return 1;
        for (int g = 0; g < n_gpu; ++g) vo_solver_destroy(sol[(size_t)g]), vo_rhs_destroy(rhs[(size_t)g]), vo_ens_destroy(shard[(size_t)g]);
        vo_solver_destroy(one), vo_ens_destroy(x0), vo_rhs_destroy(r1), vo_tableau_destroy(tab), vo_group_destroy(grp);
        for (int g = 1; g < n_gpu; ++g) vo_ctx_destroy(ctxs[(size_t)g]);
    }
    vo_ctx_destroy(ctx);
    std::printf("harness ok\n");
    return 0;
}
