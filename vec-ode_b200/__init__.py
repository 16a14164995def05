"""vecode_b200 — B200-native batched ODE time-stepping engine behind hmunozb/vec-ode's integrator API.

The directory is `vec-ode_b200/` (not an importable name); `import vecode_b200` works through the shim module
`vecode_b200.py` at the repository root.
"""
from . import _cabi, domain, group, pipeline, workloads
from ._cabi import SO_PATH, StepResult, VecOdeError, build
from .exp import DenseBasisSplit, DenseSplit, ExpCFMGeneralSolver, ExpCFMSolver, MagnusExpLinearSolver, MidpointExpLinearSolver, cfm_table, with_commutator_slot
from .split_exp import (CommutativeExpSplit, DirectSumL, ExpSplitCFMSolver, ExpSplitMidpointSolver, RKNR4ExpSplit, SemiComplexO4ExpSplit, StrangSplit,
                        TripleJumpExpSplit)
from .base import (ButcherTableu, ComplexLinearCombination, Context, Ensemble, LinearCombination, NormFn, ODEError, ODEState, ODEStep, RK45Solver, Rhs, check_step, step_many)

__all__ = ["ButcherTableu", "ComplexLinearCombination", "Context", "Ensemble", "LinearCombination", "NormFn", "ODEError", "ODEState", "ODEStep", "check_step", "RK45Solver", "Rhs", "step_many", "DenseBasisSplit", "DenseSplit", "ExpCFMSolver", "ExpCFMGeneralSolver", "ExpSplitCFMSolver", "cfm_table", "MagnusExpLinearSolver", "MidpointExpLinearSolver",
           "with_commutator_slot", "CommutativeExpSplit", "DirectSumL", "ExpSplitMidpointSolver", "RKNR4ExpSplit",
           "SemiComplexO4ExpSplit", "StrangSplit", "TripleJumpExpSplit",
           "StepResult", "VecOdeError", "build", "workloads", "group", "domain", "pipeline", "SO_PATH"]
