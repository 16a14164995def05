"""Host-side mirror of vec-ode's composite exponential splits (src/exp/split_exp.rs) over the C ABI.

The reference composes two user-supplied splits SpA, SpB (each an `ExponentialSplit`) into CommutativeExpSplit
(:24-203), StrangSplit (:205-275), SemiComplexO4ExpSplit (:281-396), TripleJumpExpSplit (:296-446) and RKNR4ExpSplit
(:449-517); their operator type is `DirectSumL { a, b }` (:48-141) and their `map_exp` is a fixed sequence of
`sp_a.map_exp` / `sp_b.map_exp` calls. With the engine's `DenseBasisSplit`, A and B are two disjoint index sets of ONE
shared basis, an operator of either is its coefficient vector, and every composite reduces to K coefficient sets applied
one after the other — which `vo_map_exp_seq` runs in a single launch with the state resident in registers.
Coefficient tables are the literals of src/dat/mod.rs:30-62.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from ._cabi import check, lib
from .base import _np_ptr
from .exp import DenseBasisSplit, _ExpSolver

_vp = C.c_void_p

# src/dat/mod.rs:34-40 (Blanes & Moan 2002, BAB convention)
RKN_O4_A = [0.209515106613362, -0.143851773179818, 0.434336666566456]
RKN_O4_B = [0.0792036964311957, 0.353172906049774, -0.0420650803577195, 0.21937695575349958]
# src/dat/mod.rs:46-62
TJ_O4_A = [complex(0.32439640402017118298, 0.13458627249080669679), complex(0.35120719195965763405, -0.26917254498161339358)]
TJ_O4_B = [complex(0.16219820201008559149, 0.06729313624540334839), complex(0.33780179798991440851, -0.06729313624540334839)]
SEMI_COMPLEX_O4_B = [complex(0.1, -1.0 / 30.0), complex(4.0 / 15.0, 2.0 / 15.0), complex(4.0 / 15.0, -1.0 / 5.0)]


@dataclass
class DirectSumL:
    """split_exp.rs:48-141: the pair (a, b) of operators of the two splits; here coefficient arrays [N][Ma], [N][Mb].
    DirectSumLinearCombination acts component-wise."""
    a: np.ndarray
    b: np.ndarray

    def scaled(self, k) -> "DirectSumL":
        return DirectSumL(self.a * k, self.b * k)

    def __add__(self, o: "DirectSumL") -> "DirectSumL":
        return DirectSumL(self.a + o.a, self.b + o.b)


class _PairSplit:
    """Two disjoint index sets of one DenseBasisSplit acting as the reference's (sp_a, sp_b)."""

    def __init__(self, sp: DenseBasisSplit, a_idx: Sequence[int], b_idx: Sequence[int]):
        assert not set(a_idx) & set(b_idx) and max(list(a_idx) + list(b_idx)) < sp.M
        self.sp, self.a_idx, self.b_idx = sp, list(a_idx), list(b_idx)

    def lin_zero(self, n_systems: int) -> DirectSumL:
        return DirectSumL(np.zeros((n_systems, len(self.a_idx)), complex), np.zeros((n_systems, len(self.b_idx)), complex))

    def _A(self, la, k=1.0):
        c = np.zeros((la.shape[0], self.sp.M), complex)
        c[:, self.a_idx] = np.asarray(la) * k
        return c

    def _B(self, lb, k=1.0):
        c = np.zeros((lb.shape[0], self.sp.M), complex)
        c[:, self.b_idx] = np.asarray(lb) * k
        return c

    def sequence(self, l: DirectSumL):  # -> list of [N][M] coefficient arrays in APPLICATION order
        raise NotImplementedError

    def exp(self, l: DirectSumL) -> np.ndarray:
        """`U`: the K lazy exponentials of the composition, stacked [K][N][M] in application order."""
        return np.ascontiguousarray(np.stack(self.sequence(l)), dtype=np.complex128)

    def multi_exp(self, l: DirectSumL, k_arr):  # exp/mod.rs:28-34 and the overrides of split_exp.rs
        return [self.exp(l.scaled(k)) for k in k_arr]

    def map_exp(self, u: np.ndarray, psi_dev_in: int, psi_dev_out: int):
        u = np.ascontiguousarray(u, dtype=np.complex128)
        check(lib().vo_map_exp_seq(self.sp._h, _np_ptr(u.view(np.float64)), u.shape[0], u.shape[1], _vp(psi_dev_in), _vp(psi_dev_out)),
              self.sp.ctx._h)

    def norm(self, psi_dev: int, n_systems: int):  # NormedExponentialSplit: sp_a.norm (split_exp.rs:187, 393)
        return self.sp.norm(psi_dev, n_systems)


class CommutativeExpSplit(_PairSplit):
    """split_exp.rs:143-203: exp(A + B) = exp(B) exp(A) for commuting A, B; map_exp applies A then B (:165-167)."""

    def sequence(self, l):
        return [self._A(l.a), self._B(l.b)]

    def commutator(self, l1: DirectSumL, l2: DirectSumL) -> DirectSumL:  # :191-202, component-wise
        full = self.sp.commutator(self._A(l1.a) + self._B(l1.b) * 0, self._A(l2.a))  # [A1, A2]
        fullb = self.sp.commutator(self._B(l1.b), self._B(l2.b))
        return DirectSumL(full[:, self.a_idx], fullb[:, self.b_idx])


class StrangSplit(_PairSplit):
    """split_exp.rs:228-275: exp scales l.b by 1/2 (:246-248); map_exp is B, A, B (:254-257)."""

    def sequence(self, l):
        return [self._B(l.b, 0.5), self._A(l.a), self._B(l.b, 0.5)]


class SemiComplexO4ExpSplit(_PairSplit):
    """split_exp.rs:333-383: u_a = exp(l.a / 4), u_b[k] = exp(SEMI_COMPLEX_O4_B[k] l.b); B0 A B1 A B2 A B1 A B0 (:364-382)."""

    def sequence(self, l):
        a, b = self._A(l.a, 0.25), [self._B(l.b, k) for k in SEMI_COMPLEX_O4_B]
        return [b[0], a, b[1], a, b[2], a, b[1], a, b[0]]


class TripleJumpExpSplit(_PairSplit):
    """split_exp.rs:410-446: u_a[k] = exp(TJ_O4_A[k] l.a), u_b[k] = exp(TJ_O4_B[k] l.b); B0 A0 B1 A1 B1 A0 B0 (:440-445)."""

    def sequence(self, l):
        a, b = [self._A(l.a, k) for k in TJ_O4_A], [self._B(l.b, k) for k in TJ_O4_B]
        return [b[0], a[0], b[1], a[1], b[1], a[0], b[0]]


class RKNR4ExpSplit(_PairSplit):
    """split_exp.rs:449-517 with RKN_O4_A/B: B0 A0 B1 A1 B2 A2 B3 A2 B2 A1 B1 A0 B0 (:506-515)."""

    def sequence(self, l):
        a, b = [self._A(l.a, k) for k in RKN_O4_A], [self._B(l.b, k) for k in RKN_O4_B]
        return [b[0], a[0], b[1], a[1], b[2], a[2], b[3], a[2], b[2], a[1], b[1], a[0], b[0]]


class ExpSplitMidpointSolver(_ExpSolver):
    """split_exp.rs:613-685 over split_exp_midpoint (:520-562), literally: the generator is sampled at t, BOTH splits are
    scaled by dt/2, and the step is A B A. `a_idx` = the basis matrices that form split A."""
    SCHEME = "split_midpoint"

    def __init__(self, sp: DenseBasisSplit, a_idx: Sequence[int], gp, t0, tf, psi0, h, M_gen=None):
        super().__init__(sp, gp, t0, tf, psi0, h, M_gen)
        mask = 0
        for m in a_idx:
            mask |= 1 << m
        check(lib().vo_exp_set_split_mask(self._h, mask), self.ctx._h)
        check(lib().vo_exp_no_adaptive(self._h), self.ctx._h)  # it only implements ODESolver


class ExpSplitCFMSolver(_ExpSolver):
    """split_cfm (split_exp.rs:568-609) as a solver: per step B(sigma_0) A(rho_0) B(sigma_1) ... A(rho_{s-1}) B(sigma_s), every
    exponent a commutator-free combination of the generator at the nodes `c` (cfm_exp, exp/cfm.rs:20-40). The reference declares
    the struct (`ExpSplitCFMSolver`, :688-705) without an impl; this is the solver its fields and `split_cfm` describe.
    `a_idx` = the basis matrices that form split A; `rho` [s][k], `sigma` [s + 1][k]."""
    SCHEME = "split_cfm"

    def __init__(self, sp: DenseBasisSplit, a_idx: Sequence[int], gp, t0, tf, psi0, h, c, rho, sigma, M_gen=None):
        super().__init__(sp, gp, t0, tf, psi0, h, M_gen)
        mask = 0
        for m in a_idx:
            mask |= 1 << m
        check(lib().vo_exp_set_split_mask(self._h, mask), self.ctx._h)
        c = np.ascontiguousarray(c, dtype=np.float64).ravel()
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        if rho.ndim != 2 or sigma.ndim != 2 or rho.shape[1] != c.size or sigma.shape[1] != c.size or sigma.shape[0] != rho.shape[0] + 1:
            from ._cabi import VO_ERR_SHAPE, VecOdeError
            raise VecOdeError(VO_ERR_SHAPE, "split_cfm: Incompatible array dimensions")  # split_exp.rs:587-592
        check(lib().vo_exp_set_split_cfm_tables(self._h, _np_ptr(c), c.size, _np_ptr(rho), _np_ptr(sigma), rho.shape[0]), self.ctx._h)
        check(lib().vo_exp_no_adaptive(self._h), self.ctx._h)
