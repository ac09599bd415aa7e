"""A very small affine/convex expression layer so that the reference's loss / constraint
CALLBACKS keep working without cvxpy.

The reference hands cvxpy Variables to user callbacks (`tzddpc/tzddpc.py:213,222`;
examples: `examples/1.double_integrator_sim.py:22-34`, `examples/2.pulley_sim.py:17-28`,
`examples/3.5dimsystem_sim.py:14-26`) which combine them with `cp.norm(.., p=1|2)`, `** 2`,
sums, slices and `<=`/`>=`.  cvxpy is absent here and banned from the GPU path, so the
callbacks receive `Variable`s of this module instead:

    from tzddpc_b200 import cvx as cp      # instead of `import cvxpy as cp`

Only what reduces to the structured `StageCost` / `BoxConstraint` of program.py is
accepted (|affine scalar|, squared 2-norms / quadratic forms of affine vectors, box
constraints); anything else raises NotImplementedError.  Host-side, build time only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from .program import BoxConstraint, StageCost


class Expression:
    """Array of affine forms: coef[..., 0] constant, coef[..., 1:] over the flat unknown vector."""
    __array_priority__ = 100

    def __init__(self, coef: np.ndarray):
        self.coef = np.asarray(coef, dtype=np.float64)

    @property
    def shape(self) -> Tuple[int, ...]:
        return self.coef.shape[:-1]

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    def is_dcp(self) -> bool:
        return True

    def __getitem__(self, idx) -> "Expression":
        if not isinstance(idx, tuple):
            idx = (idx,)
        return Expression(self.coef[idx + (slice(None),)])

    @staticmethod
    def _lift(other, nunk: int, shape) -> "Expression":
        if isinstance(other, Expression):
            return other
        a = np.asarray(other, dtype=np.float64)
        c = np.zeros(a.shape + (1 + nunk,))
        c[..., 0] = a
        return Expression(c)

    def __add__(self, other):
        if isinstance(other, ConvexSum):
            return other + self
        o = self._lift(other, self.coef.shape[-1] - 1, self.shape)
        return Expression(self.coef + o.coef)

    __radd__ = __add__

    def __neg__(self):
        return Expression(-self.coef)

    def __sub__(self, other):
        o = self._lift(other, self.coef.shape[-1] - 1, self.shape)
        return Expression(self.coef - o.coef)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        a = np.asarray(other, dtype=np.float64)
        return Expression(self.coef * a[..., None])

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self * (1.0 / float(other))

    def __rmatmul__(self, M):
        return Expression(np.tensordot(np.asarray(M, dtype=np.float64), self.coef, axes=(-1, 0)))

    def __matmul__(self, M):
        M = np.asarray(M, dtype=np.float64)
        return Expression(np.moveaxis(np.tensordot(self.coef, M, axes=(-2, 0)), -2, -1))

    @property
    def T(self):
        return Expression(np.swapaxes(self.coef, 0, 1)) if len(self.shape) == 2 else self

    def __le__(self, other):
        return Constraint(self - other)                    # self - other <= 0

    def __ge__(self, other):
        return Constraint((-self) + other)                 # other - self <= 0

    def __eq__(self, other):                               # noqa: PLW1641
        raise NotImplementedError("equality constraints in user callbacks are not supported")

    def __pow__(self, p):
        if p == 2 and self.size == 1:
            return sum_squares(self)
        raise NotImplementedError("only squares are supported")


class Variable(Expression):
    def __init__(self, shape, offset: int = 0, nunk: Optional[int] = None):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        size = int(np.prod(shape))
        nunk = size if nunk is None else nunk
        c = np.zeros((size, 1 + nunk))
        c[np.arange(size), 1 + offset + np.arange(size)] = 1.0
        super().__init__(c.reshape(shape + (1 + nunk,)))


class Constraint:
    """expr <= 0 elementwise."""

    def __init__(self, expr: Expression):
        self.expr = expr

    def is_dcp(self) -> bool:
        return True


class _Abs:
    def __init__(self, form: np.ndarray):
        self.form = form            # (1+nunk,)


class _Quad:
    def __init__(self, forms: np.ndarray, Q: Optional[np.ndarray] = None):
        self.forms, self.Q = forms, Q      # (k, 1+nunk), optional k x k weight


class _Norm2:
    """||affine vector||_2, only meaningful once squared."""

    def __init__(self, forms: np.ndarray):
        self.forms = forms

    def __pow__(self, p):
        if p != 2:
            raise NotImplementedError("only norm(.)**2 is supported for vector 2-norms")
        return ConvexSum([(1.0, _Quad(self.forms))])

    def is_dcp(self) -> bool:
        return True

    def _as_sum(self):
        raise NotImplementedError("an un-squared 2-norm of a vector is a second-order cone term: not representable "
                                  "in the QP the GPU path solves")

    __add__ = __radd__ = __mul__ = __rmul__ = lambda self, other: self._as_sum()


class ConvexSum(Expression):
    """Non-negative combination of |affine| and quadratic atoms (+ an affine part)."""

    def __init__(self, terms: Sequence[Tuple[float, Union[_Abs, _Quad]]], affine: Optional[np.ndarray] = None):
        self.terms = list(terms)
        self.affine = affine
        self.coef = np.zeros((1,))

    def is_dcp(self) -> bool:
        return all(w >= 0 for w, _ in self.terms)

    def __add__(self, other):
        if isinstance(other, ConvexSum):
            return ConvexSum(self.terms + other.terms)
        if isinstance(other, _Norm2):
            other._as_sum()
        if isinstance(other, Expression):
            if np.any(other.coef[..., 1:] != 0.0):
                raise NotImplementedError("affine cost terms are not supported")
            return self
        if np.isscalar(other):
            return self                                   # a constant offset (callbacks start from `cost = 0`)
        return NotImplemented

    __radd__ = __add__

    def __mul__(self, other):
        w = float(other)
        return ConvexSum([(w * a, t) for a, t in self.terms])

    __rmul__ = __mul__

    def __pow__(self, p):
        if p == 2 and len(self.terms) == 1 and isinstance(self.terms[0][1], _Abs):
            w, t = self.terms[0]
            return ConvexSum([(w * w, _Quad(t.form[None, :]))])
        raise NotImplementedError("only squares of a single |.| / norm are supported")


def _flat_forms(e: Expression) -> np.ndarray:
    return e.coef.reshape(-1, e.coef.shape[-1])


def norm(e, p=2):
    if not isinstance(e, Expression):
        raise NotImplementedError("norm of a constant")
    f = _flat_forms(e)
    if p == 1:
        return ConvexSum([(1.0, _Abs(row)) for row in f])
    if p == 2:
        return ConvexSum([(1.0, _Abs(f[0]))]) if f.shape[0] == 1 else _Norm2(f)
    raise NotImplementedError(f"norm p={p!r} is not supported (p=1, p=2)")


def abs(e):                                                # noqa: A001  (mirrors cp.abs)
    return ConvexSum([(1.0, _Abs(row)) for row in _flat_forms(e)])


def sum_squares(e):
    return ConvexSum([(1.0, _Quad(_flat_forms(e)))])


def quad_form(e, Q):
    return ConvexSum([(1.0, _Quad(_flat_forms(e), np.asarray(Q, dtype=np.float64)))])


def sum(e):                                                # noqa: A001
    if isinstance(e, ConvexSum):
        return e
    raise NotImplementedError("sum of affine expressions as a cost")


def _rows(N: int, n: int, m: int, simplified: bool):
    """(u, x) variables as the reference builds them: u (N x m); x (N+1 x n) [build_problem, :222]
    or (N x n) [build_problem_simplified, :336].  Unknown vector = [u.flat, x.flat]."""
    xr = N if simplified else N + 1
    nunk = N * m + xr * n
    return Variable((N, m), 0, nunk), Variable((xr, n), N * m, nunk), xr, nunk


def extract_stage_cost(build_loss, N: int, n: int, m: int, simplified: bool) -> StageCost:
    u, x, xr, nunk = _rows(N, n, m, simplified)
    loss = build_loss(u, x)
    if loss is None or not isinstance(loss, (ConvexSum, _Norm2)) or not loss.is_dcp():
        raise Exception('Loss function is not defined or is not convex!')          # tzddpc/tzddpc.py:224-225
    if isinstance(loss, _Norm2):
        loss._as_sum()
    # per-row accumulators
    Qx = np.zeros((xr, n, n)); bx = np.zeros((xr, n)); wx = np.zeros((xr, n)); rx = np.full((xr, n), np.nan)
    Qu = np.zeros((N, m, m)); bu = np.zeros((N, m)); wu = np.zeros((N, m)); ru = np.full((N, m), np.nan)

    def locate(form: np.ndarray):
        nzc = np.flatnonzero(form[1:])
        ku = nzc[nzc < N * m]
        kx = nzc[nzc >= N * m] - N * m
        if len(ku) and len(kx):
            raise NotImplementedError("cost terms mixing inputs and states")
        if len(ku):
            rows = np.unique(ku // m)
            if len(rows) > 1:
                raise NotImplementedError("cost terms coupling several stages")
            return "u", int(rows[0]), form[1 + rows[0] * m: 1 + (rows[0] + 1) * m]
        if len(kx):
            rows = np.unique(kx // n)
            if len(rows) > 1:
                raise NotImplementedError("cost terms coupling several stages")
            return "x", int(rows[0]), form[1 + N * m + rows[0] * n: 1 + N * m + (rows[0] + 1) * n]
        return None, 0, None

    for w, t in loss.terms:
        if isinstance(t, _Abs):
            kind, row, a = locate(t.form)
            if kind is None:
                continue
            j = np.flatnonzero(a)
            if len(j) != 1:
                raise NotImplementedError("|.| of a combination of several components")
            j = int(j[0])
            wgt, ref = w * np.abs(a[j]), -t.form[0] / a[j]
            W_, R_ = (wx, rx) if kind == "x" else (wu, ru)
            if W_[row, j] != 0.0 and not np.isclose(R_[row, j], ref):
                raise NotImplementedError("two |.| terms with different references on one component")
            W_[row, j] += wgt
            R_[row, j] = ref
        else:
            F = t.forms
            Qw = np.eye(F.shape[0]) if t.Q is None else t.Q
            kinds = [locate(f) for f in F]
            kind = next((k for k, _, _ in kinds if k), None)
            if kind is None:
                continue
            rows = {r for k, r, _ in kinds if k}
            if len(rows) > 1 or any(k not in (None, kind) for k, _, _ in kinds):
                raise NotImplementedError("quadratic terms coupling several stages")
            row = rows.pop()
            dim = n if kind == "x" else m
            M = np.stack([a if a is not None else np.zeros(dim) for _, _, a in kinds])
            c = F[:, 0]
            Q_, B_ = (Qx, bx) if kind == "x" else (Qu, bu)
            Q_[row] += w * M.T @ Qw @ M
            B_[row] += w * M.T @ Qw @ c               # (M z + c)'Q(M z + c) = z'M'QMz + 2 c'QM z + const

    def stage(Q_, B_, W_, R_, nrows, dim, what):
        used = [r for r in range(nrows) if np.any(Q_[r]) or np.any(W_[r])]
        if not used:
            return None, None, None
        r0 = used[0]
        for r in used[1:]:
            if not (np.allclose(Q_[r], Q_[r0]) and np.allclose(B_[r], B_[r0]) and np.allclose(W_[r], W_[r0])
                    and np.allclose(np.nan_to_num(R_[r]), np.nan_to_num(R_[r0]))):
                raise NotImplementedError(f"stage cost on {what} differs between stages")
        expect = list(range(N))
        if used != expect:
            raise NotImplementedError(f"stage cost on {what} must cover stages 0..N-1 exactly (got rows {used})")
        ref = np.nan_to_num(R_[r0])
        Q = Q_[r0] if np.any(Q_[r0]) else None
        if Q is not None:
            qref = -np.linalg.lstsq(Q, B_[r0], rcond=None)[0]
            if not np.allclose(Q @ qref, -B_[r0]):
                raise NotImplementedError("quadratic cost with an affine part outside its range")
            if np.any(W_[r0]) and not np.allclose(qref[W_[r0] > 0], ref[W_[r0] > 0]):
                raise NotImplementedError("quadratic and |.| terms with different references")
            ref = np.where(W_[r0] > 0, ref, qref)
        return Q, ref, (W_[r0] if np.any(W_[r0]) else None)

    Q, x_ref, w_abs = stage(Qx, bx, wx, rx, xr, n, "x")
    R, u_ref, r_abs = stage(Qu, bu, wu, ru, N, m, "u")
    return StageCost(Q=Q, x_ref=x_ref, w_abs=w_abs, R=R, u_ref=u_ref, r_abs=r_abs)


def extract_box_constraints(build_constraints, N: int, n: int, m: int, simplified: bool) -> BoxConstraint:
    u, x, xr, nunk = _rows(N, n, m, simplified)
    cons = build_constraints(u, x)
    x_lo = np.full((xr, n), -np.inf); x_hi = np.full((xr, n), np.inf)
    v_lo = np.full((N, m), -np.inf); v_hi = np.full((N, m), np.inf)
    for idx, c in enumerate(cons if cons is not None else (None, None)):
        if c is None or not isinstance(c, Constraint) or not c.is_dcp():
            raise Exception(f'Constraint {idx} is not defined or is not convex.')   # tzddpc/tzddpc.py:216-217
        for f in _flat_forms(c.expr):
            nzc = np.flatnonzero(f[1:])
            if len(nzc) == 0:
                if f[0] > 0:
                    raise Exception(f'Constraint {idx} is infeasible')
                continue
            if len(nzc) != 1:
                raise NotImplementedError("only box constraints on single components are supported")
            k, a = int(nzc[0]), f[1 + nzc[0]]
            bound = -f[0] / a                                 # a z + c <= 0
            if k < N * m:
                tgt_lo, tgt_hi, r, j = v_lo, v_hi, k // m, k % m
            else:
                tgt_lo, tgt_hi, r, j = x_lo, x_hi, (k - N * m) // n, (k - N * m) % n
            if a > 0:
                tgt_hi[r, j] = min(tgt_hi[r, j], bound)
            else:
                tgt_lo[r, j] = max(tgt_lo[r, j], bound)

    def collapse(a, what):
        if not np.all(a == a[0]):
            raise NotImplementedError(f"box constraints on {what} must be the same for every stage")
        return a[0] if np.any(np.isfinite(a[0])) else None

    return BoxConstraint(x_lo=collapse(x_lo, "x"), x_hi=collapse(x_hi, "x"),
                         v_lo=collapse(v_lo, "v"), v_hi=collapse(v_hi, "v"))
