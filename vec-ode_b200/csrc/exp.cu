// exp.cu — exponential integrators (src/exp). PLACEHOLDER entry points: replaced by the batched complex kernels.
#include "common.cuh"
extern "C" {
#define VO_EXP_STUB(ctx) return vo_fail(ctx, VO_ERR_UNSUPPORTED, "exponential integrators: not built yet")
int32_t vo_split_basis_create(vo_ctx c, int32_t, int32_t, const double*, vo_split*) { VO_EXP_STUB(c); }
int32_t vo_split_destroy(vo_split) { return VO_OK; }
int32_t vo_split_set_commutator(vo_split, const double*) { VO_EXP_STUB(nullptr); }
int32_t vo_split_set_taylor_degree(vo_split, int32_t) { VO_EXP_STUB(nullptr); }
int32_t vo_map_exp(vo_split, const double*, int64_t, void*, void*) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_create(vo_ctx c, vo_split, int32_t, int32_t, const double*, int64_t, double, double, const double*, double, vo_expsolver*) { VO_EXP_STUB(c); }
int32_t vo_exp_destroy(vo_expsolver) { return VO_OK; }
int32_t vo_exp_no_adaptive(vo_expsolver) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_with_tolerance(vo_expsolver, double, double) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_with_step_range(vo_expsolver, double, double) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_step(vo_expsolver, vo_step_result*) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_step_adaptive(vo_expsolver, vo_step_result*) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_run(vo_expsolver, int32_t, int64_t, vo_step_result*) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_current(vo_expsolver, double*, double*, double*) { VO_EXP_STUB(nullptr); }
int32_t vo_exp_stats(vo_expsolver, int64_t*, int64_t*, double*, double*, double*) { VO_EXP_STUB(nullptr); }
void* vo_exp_state_device_ptr(vo_expsolver) { return nullptr; }
}
