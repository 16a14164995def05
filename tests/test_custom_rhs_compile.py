"""User RHS source -> sm_100a cubin with NVRTC (vo_rhs_custom_check): needs no GPU, so the run-time compilation path of
vo_rhs_create_custom (the closure replacement, src/base/rk.rs:97) is checked on CPU too. No kernel is launched here."""
import pytest

BODY = "dx[0] = p[0] * x[0] - p[1] * x[0] * x[1];\ndx[1] = p[3] * x[0] * x[1] - p[2] * x[1];"


@pytest.mark.parametrize("stages,arith", [(-1, "strict"), (7, "fast"), (5, "strict")])
def test_user_rhs_compiles_into_the_kernel_templates(vo, stages, arith):
    assert vo.Rhs.check_source(BODY, 2, 4, stages, arith) > 10_000  # bytes of cubin


def test_user_rhs_compile_error_carries_the_log(vo):
    with pytest.raises(vo.VecOdeError) as ei:
        vo.Rhs.check_source("dx[0] = nope;", 1, 0)
    assert "rhs_body(1)" in str(ei.value) and "nope" in str(ei.value)
    with pytest.raises(vo.VecOdeError):
        vo.Rhs.check_source("dx[0] = x[0];", 33, 0)


def test_user_generator_compiles_into_the_tensor_core_kernel(vo):
    """vo_exp_generator_check: the generator closure of the exponential integrators -> exp_step_kernel, no GPU needed."""
    assert vo.ExpCFMSolver.check_generator("g[1] = p[0] * cos(p[1] * t + p[2]);", 64, 2) > 10_000
    with pytest.raises(vo.VecOdeError) as ei:
        vo.ExpCFMSolver.check_generator("g[1] = nope;", 16, 2)
    assert "generator_body(1)" in str(ei.value) and "nope" in str(ei.value)


WRMS_MAP, WRMS_FINISH = "m = (e * e + im * im) / (1.0 + i);", "r = sqrt(acc / n);"


def test_user_norm_compiles_alone_and_into_the_solver_kernels(vo):
    """vo_normfn_check / vo_normfn_check_kernels: the user's `Normed` impl (ode.rs:9-11) and NormFn closure (cfm.rs:105) as source —
    the reduction kernels, the register-resident control kernels of a compiled-in family (VdP, DoPri5, both arithmetic modes) and
    the DMMA kernel of the exponential integrators, all without a GPU."""
    assert vo.NormFn.check_source(WRMS_MAP, "sum", WRMS_FINISH) > 5_000
    assert vo.NormFn.check_source("m = fabs(e);", "max") > 5_000
    for arith in ("strict", "fast"):
        assert vo.NormFn.check_kernels(WRMS_MAP, "sum", WRMS_FINISH, rhs_kind=3, d=2, stages=7, arith=arith) > 50_000
    assert vo.NormFn.check_kernels(WRMS_MAP, "sum", WRMS_FINISH, rhs_kind=2, d=3, stages=5) > 50_000   # Lorenz, generic stage count
    assert vo.NormFn.check_kernels("m = fabs(e) + fabs(im);", "max", "", exp_n=16, exp_M=2) > 10_000
    assert vo.NormFn.check_kernels(WRMS_MAP, "sum", WRMS_FINISH, exp_n=64, exp_M=2) > 10_000


def test_user_norm_compile_error_carries_the_log(vo):
    with pytest.raises(vo.VecOdeError) as ei:
        vo.NormFn.check_source("m = nope;", "sum")
    assert "norm_map_body(1)" in str(ei.value) and "nope" in str(ei.value)
    with pytest.raises(vo.VecOdeError) as ei:
        vo.NormFn.check_source("m = e * e;", "sum", "r = sqrt(oops);")
    assert "norm_finish_body(1)" in str(ei.value)


HEAT_STENCIL = "du = p[0] * ((u[0] + u[2]) - 2.0 * u[1]);"


def test_user_stencil_and_wide_pointwise_rhs_compile(vo):
    """vo_rhs_custom_stencil_check (a grid stencil into the fused per-stage kernel of rk_stage_stencil.cuh) and a 24-component
    pointwise right-hand side into the stage-path kernel: both without a GPU."""
    for arith in ("strict", "fast"):
        assert vo.Rhs.check_stencil_source(HEAT_STENCIL, 1, 1, arith) > 5_000
    assert vo.Rhs.check_stencil_source("du = (-u[0] + 16.0 * u[1] - 30.0 * u[2] + 16.0 * u[3] - u[4]) * (p[0] / 12.0) + p[1] * sin(t) * (j == 0);", 2, 2) > 5_000
    with pytest.raises(vo.VecOdeError) as ei:
        vo.Rhs.check_stencil_source("du = u[7];\nnope();", 1, 0)
    assert "rhs_body(2)" in str(ei.value)
    ring = "\n".join(f"dx[{c}] = p[0] * (x[{(c + 1) % 24}] - x[{c}]) - x[{c}] * x[{c}] * x[{c}];" for c in range(24))
    assert vo.Rhs.check_source(ring, 24, 1, -1, "strict") > 10_000
    with pytest.raises(vo.VecOdeError):
        vo.Rhs.check_source(ring, 24, 1, 7, "strict")  # the register-resident kernels stop at 8 components
