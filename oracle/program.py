"""The TZDDPC controller -- ORACLE restatement (test infrastructure, parity unpinned).

Follows the reference class `TZDDPC` statement by statement:

  update_identification_data   tzddpc/tzddpc.py:45-65
  build_zonotopes              tzddpc/tzddpc.py:67-85
  build_zonotopes_theta        tzddpc/tzddpc.py:95-130   (K is an input: gain synthesis,
                                                          tzddpc/utils.py:60-103, is out of scope)
  build_problem                tzddpc/tzddpc.py:132-241  (quirks Q4-Q8 of SURVEY.md 3.5 kept)
  build_problem_simplified     tzddpc/tzddpc.py:243-355
  solve                        tzddpc/tzddpc.py:357-377
  closed loop                  examples/2.pulley_sim.py:81-96

The reference builds the program symbolically with cvxpy.  cvxpy is absent, so
the "symbolic" objects here are affine-coefficient tensors: every entry of a
`CVXZonotope.Z` is an affine function of  w = [1, v (N*m), xbar0 (n), e0 (n)]
(the nominal states xbar_k are eliminated through the equality constraints
`:166-170`), stored as Zc[row, column, coefficient].  The loss/constraint
callbacks of the reference are cvxpy expressions; here they are the structured
`StageCost` / `BoxConstraint` records that describe the same three example
losses (`examples/1.double_integrator_sim.py:22-28`, `examples/2.pulley_sim.py:17-22`,
`examples/3.5dimsystem_sim.py:14-26`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, NamedTuple, Optional, Tuple

import numpy as np

from .qp import solve_qp_ipm
from .zono import (Conventions, DEFAULT, MatrixZonotope, Zonotope, compute_LTI_matrix_zonotope,
                   concatenate_zonotope, matzono_times_Z)


class Data(NamedTuple):          # tzddpc/objects.py:33-40
    u: np.ndarray
    x: np.ndarray


class DataDrivenDataset(NamedTuple):   # tzddpc/objects.py:43-52
    Xp: np.ndarray
    Xm: np.ndarray
    Um: np.ndarray
    original_data: Data


class SystemZonotopes(NamedTuple):     # tzddpc/objects.py:55-67
    X0: Zonotope
    U: Zonotope
    X: Zonotope
    W: Zonotope


class Theta(NamedTuple):               # tzddpc/objects.py:69-72
    K: np.ndarray
    deltaA: np.ndarray
    deltaB: np.ndarray


@dataclass
class StageCost:
    """sum_i [ (x_i-x_ref)'Q(x_i-x_ref) + sum_j w_abs[j]|x_i[j]-x_ref[j]|
              + (u_i-u_ref)'R(u_i-u_ref) + sum_j r_abs[j]|u_i[j]-u_ref[j]| ].
    In `build_problem` the callback receives the FREE variable `u`
    (tzddpc/tzddpc.py:160,222 -- quirk Q7) so the u-terms sit at their floor 0 and
    the x-rows are xbar_0..xbar_{N-1}; in `build_problem_simplified` it receives
    (v, xbar[1:]) (`:336`)."""
    Q: Optional[np.ndarray] = None
    x_ref: Optional[np.ndarray] = None
    w_abs: Optional[np.ndarray] = None
    R: Optional[np.ndarray] = None
    u_ref: Optional[np.ndarray] = None
    r_abs: Optional[np.ndarray] = None


@dataclass
class BoxConstraint:
    """User constraints of the form lo <= xbar[:, j] <= hi / lo <= v[:, j] <= hi
    (examples/3.5dimsystem_sim.py:23-26).  Applied to ALL rows the callback is
    handed: xbar_0..xbar_N in build_problem (`:213`), xbar_1..xbar_N in the simplified one (`:327`)."""
    x_lo: Optional[np.ndarray] = None
    x_hi: Optional[np.ndarray] = None
    v_lo: Optional[np.ndarray] = None
    v_hi: Optional[np.ndarray] = None


class SolveResult(NamedTuple):
    status: int                 # 0 optimal, 2 infeasible, 1 solver did not converge
    cost: float
    v: np.ndarray               # N x m
    xbar: np.ndarray            # (N+1) x n
    Ze1: np.ndarray             # n x (1+g1): Ze[1].Z.value, column 0 the centre
    Ze_all: List[np.ndarray]    # numeric Ze[k].Z for k = 0..N (diagnostics / parity of longer horizons)


class _Tube:
    """Constraint `center_r + sum_j |gen_rj| <= hi_r`, `center_r - sum_j|gen_rj| >= lo_r`
    with every entry affine in w (coefficient tensors)."""

    def __init__(self, center: np.ndarray, gens: np.ndarray, lo: np.ndarray, hi: np.ndarray):
        self.center, self.gens, self.lo, self.hi = center, gens, lo, hi     # (d,nw), (d,g,nw), (d,), (d,)


class OracleTZDDPC:
    def __init__(self, data: Data, conv: Conventions = DEFAULT):
        self.conv = conv
        self.update_identification_data(data)

    # tzddpc/tzddpc.py:30-43
    @property
    def num_samples(self) -> int:
        return self.dataset.Um.shape[0] + 1

    @property
    def dim_u(self) -> int:
        return self.dataset.Um.shape[1]

    @property
    def dim_x(self) -> int:
        return self.dataset.Xp.shape[1]

    # tzddpc/tzddpc.py:45-65
    def update_identification_data(self, data: Data):
        assert len(data.u.shape) == 2 and len(data.x.shape) == 2
        assert data.x.shape[0] == data.u.shape[0], "Input/state data must have the same length"
        self.dataset = DataDrivenDataset(data.x[1:], data.x[:-1], data.u[:-1], data)

    # tzddpc/tzddpc.py:67-85
    def build_zonotopes(self, zonotopes: SystemZonotopes) -> MatrixZonotope:
        X0, W, X = zonotopes.X0, zonotopes.W, zonotopes.X
        assert X0.dimension == W.dimension == self.dim_x == X.dimension, \
            "The zonotopes do not have the correct dimension"
        self.zonotopes = zonotopes
        Mw = concatenate_zonotope(W, self.num_samples - 1, self.conv)
        self.Mdata = compute_LTI_matrix_zonotope(self.dataset.Xm, self.dataset.Xp, self.dataset.Um, Mw)
        return self.Mdata

    # tzddpc/tzddpc.py:95-130 with K supplied
    def build_zonotopes_theta(self, zonotopes: SystemZonotopes, K: np.ndarray) -> Tuple[Theta, MatrixZonotope]:
        self.build_zonotopes(zonotopes)
        n, m = self.dim_x, self.dim_u
        K = np.asarray(K, dtype=np.float64).reshape(m, n)
        self.theta = Theta(K, np.zeros((n, n)), np.zeros((n, m)))
        self.MdataK = self.Mdata * np.vstack([np.eye(n), K])                # :119 (unreduced Mdata)
        self.Mdelta = self.Mdata + (-1.0 * self.Mdata.center)               # :122-123
        self.Mdata = self.Mdata.reduce(1)                                   # :126-128
        self.MdataK = self.MdataK.reduce(1)
        self.Mdelta = self.Mdelta.reduce(1)
        return self.theta, self.Mdata

    # ------------------------------------------------------------------------------
    # build_problem: tzddpc/tzddpc.py:132-241 ; simplified: :243-355 (k0 is not None)
    # ------------------------------------------------------------------------------
    def build_problem(self, horizon: int, cost: StageCost, box: Optional[BoxConstraint] = None,
                      k0: Optional[int] = None):
        n, m, N = self.dim_x, self.dim_u, int(horizon)
        simplified = k0 is not None
        nw = 1 + N * m + 2 * n
        iv = lambda k, j: 1 + k * m + j          # noqa: E731  coefficient slot of v[k, j]
        ix0, ie0 = 1 + N * m, 1 + N * m + n
        A, B = self.Mdata.center[:, :n], self.Mdata.center[:, n:]           # :163
        K = self.theta.K
        W = self.zonotopes.W

        # nominal states, eliminated through :166-170
        xbar = np.zeros((N + 1, n, nw))
        xbar[0, np.arange(n), ix0 + np.arange(n)] = 1.0
        vaff = np.zeros((N, m, nw))
        for k in range(N):
            for j in range(m):
                vaff[k, j, iv(k, j)] = 1.0
            xbar[k + 1] = A @ xbar[k] + B @ vaff[k]

        def const_zono(Z: np.ndarray) -> np.ndarray:
            out = np.zeros(Z.shape + (nw,))
            out[..., 0] = Z
            return out

        def mz_times(M: MatrixZonotope, Zc: np.ndarray) -> np.ndarray:
            return matzono_times_Z(M.center, M.generators, Zc)

        def plus_zono(Zc: np.ndarray, Zother: np.ndarray) -> np.ndarray:
            """Minkowski sum [c1+c2, G1, G2] (App. A.2)."""
            out = np.concatenate([Zc, Zother[:, 1:]], axis=1)
            out[:, 0] = Zc[:, 0] + Zother[:, 0]
            return out

        Wc = const_zono(W.Z)
        e0c = np.zeros((n, 2, nw))
        e0c[np.arange(n), 0, ie0 + np.arange(n)] = 1.0                       # :172  <e0, zeros(n,1)>
        Ze = [e0c]
        XU = []
        for k in range(N):                                                   # :174
            z = np.zeros((n + m, 2, nw))
            z[:n, 0] = xbar[k]
            z[n:, 0] = vaff[k]
            XU.append(z)
        T1 = [mz_times(self.MdataK, Ze[0])]                                  # :175
        Zn = [plus_zono(mz_times(self.Mdelta, XU[k]), Wc) for k in range(N)]  # :176
        T2 = []
        for k in range(N):
            if not simplified:                                               # :180-186
                T1.append(mz_times(self.MdataK, T1[-1]))
                noise = Zn[0]
                for j in range(1, k):
                    noise = plus_zono(mz_times(self.MdataK, noise), Zn[j])
            else:                                                            # :290-302
                T1.append(T1[-1] if k > k0 else mz_times(self.MdataK, T1[-1]))
                start = max(0, k - k0)
                noise = Zn[start]
                for j in range(1, min(k, k0)):
                    noise = plus_zono(mz_times(self.MdataK, noise), Zn[start + j])
            T2.append(noise)

        Xi, Ui = self.zonotopes.X.interval, self.zonotopes.U.interval
        tubes: List[_Tube] = []
        self.num_generators_log = []
        for k in range(N):                                                   # :189-209
            Zk = Ze[-1]
            tubes.append(_Tube(Zk[:, 0] + xbar[k], Zk[:, 1:], Xi.left_limit, Xi.right_limit))        # :191,194-195
            KZ = np.tensordot(K, Zk, axes=(1, 0))
            tubes.append(_Tube(KZ[:, 0] + vaff[k], KZ[:, 1:], Ui.left_limit, Ui.right_limit))        # :192,196-197
            Ze_new = plus_zono(T1[k], T2[k])                                 # :205
            self.num_generators_log.append(Ze_new.shape[1] - 1)              # :206
            Ze.append(Ze_new)

        self._N, self._nw, self._slots = N, nw, (iv, ix0, ie0)
        self._xbar_aff, self._v_aff, self._Ze_aff, self._tubes = xbar, vaff, Ze, tubes
        self._cost, self._box, self._simplified = cost, (box or BoxConstraint()), simplified
        return self

    # ------------------------------------------------------------------------------
    def _assemble(self, xbar0: np.ndarray, e0: np.ndarray):
        """Substitute the parameters and put the program in  min .5 y'Py+q'y+c0  s.t. Gy<=h,
        y = [v ; one epigraph variable per decision-dependent |.| atom]."""
        n, m, N = self.dim_x, self.dim_u, self._N
        nv = N * m
        iv, ix0, ie0 = self._slots
        par = np.r_[xbar0, e0]

        def split(aff: np.ndarray):
            """affine tensor (..., nw) -> (constant (...), v-coefficients (..., nv))"""
            return aff[..., 0] + aff[..., 1 + nv:] @ par, aff[..., 1:1 + nv]

        rows_G: List[np.ndarray] = []       # each row over [v ; t...] is built lazily
        rows_h: List[float] = []
        atoms: List[Tuple[float, np.ndarray]] = []     # epigraph variables t_i >= |c_i + a_i v|
        param_ok = True

        def add_row(av: np.ndarray, t_idx: List[int], t_w: List[float], rhs: float):
            rows_G.append((av, t_idx, t_w))
            rows_h.append(rhs)

        def abs_sum(gconst: np.ndarray, gv: np.ndarray):
            """sum_j |gconst_j + gv_j v| -> (numeric constant, [epigraph indices])"""
            cst, idx = 0.0, []
            for j in range(gconst.shape[0]):
                if np.any(gv[j] != 0.0):
                    atoms.append((gconst[j], gv[j]))
                    idx.append(len(atoms) - 1)
                else:
                    cst += abs(gconst[j])
            return cst, idx

        for tb in self._tubes:
            cc, cv = split(tb.center)
            gc, gv = split(tb.gens)
            for r in range(cc.shape[0]):
                cst, idx = abs_sum(gc[r], gv[r])
                ones = [1.0] * len(idx)
                if np.isfinite(tb.hi[r]):
                    if not idx and not np.any(cv[r]):
                        param_ok &= bool(cc[r] + cst <= tb.hi[r] + 1e-9 * max(1.0, abs(tb.hi[r])))
                    else:
                        add_row(cv[r], idx, ones, tb.hi[r] - cc[r] - cst)                 # c + sum|g| <= hi
                if np.isfinite(tb.lo[r]):
                    if not idx and not np.any(cv[r]):
                        param_ok &= bool(cc[r] - cst >= tb.lo[r] - 1e-9 * max(1.0, abs(tb.lo[r])))
                    else:
                        add_row(-cv[r], idx, ones, -tb.lo[r] + cc[r] - cst)               # -(c - sum|g|) <= -lo

        # user constraints (examples/3.5dimsystem_sim.py:23-26)
        bx = self._box
        xrows = range(1, N + 1) if self._simplified else range(N + 1)
        for k in xrows:
            cc, cv = split(self._xbar_aff[k])
            for j in range(n):
                for lim, sgn in ((bx.x_hi, 1.0), (bx.x_lo, -1.0)):
                    if lim is not None and np.isfinite(lim[j]):
                        if not np.any(cv[j]):
                            param_ok &= bool(sgn * cc[j] <= sgn * lim[j] + 1e-9 * max(1.0, abs(lim[j])))
                        else:
                            add_row(sgn * cv[j], [], [], sgn * (lim[j] - cc[j]))
        for k in range(N):
            cc, cv = split(self._v_aff[k])
            for j in range(m):
                for lim, sgn in ((bx.v_hi, 1.0), (bx.v_lo, -1.0)):
                    if lim is not None and np.isfinite(lim[j]):
                        add_row(sgn * cv[j], [], [], sgn * (lim[j] - cc[j]))

        # objective
        c = self._cost
        P = np.zeros((nv, nv))
        q = np.zeros(nv)
        c0 = 0.0
        cost_atoms: List[Tuple[float, int]] = []
        if self._simplified:
            xcost_rows = [self._xbar_aff[k] for k in range(1, N + 1)]
            ucost_rows = [self._v_aff[k] for k in range(N)]
        else:
            xcost_rows = [self._xbar_aff[k] for k in range(N)]      # Q7: xbar_0..xbar_{N-1}
            ucost_rows = []                                         # Q7: free u sits at the floor
        x_ref = np.zeros(n) if c.x_ref is None else np.asarray(c.x_ref, dtype=np.float64)
        u_ref = np.zeros(m) if c.u_ref is None else np.asarray(c.u_ref, dtype=np.float64)
        for rows, Qm, ref, wabs in ((xcost_rows, c.Q, x_ref, c.w_abs), (ucost_rows, c.R, u_ref, c.r_abs)):
            for aff in rows:
                cc, cv = split(aff)
                d = cc - ref
                if Qm is not None:
                    Qm_ = np.asarray(Qm, dtype=np.float64)
                    P += 2.0 * cv.T @ Qm_ @ cv
                    q += 2.0 * cv.T @ Qm_ @ d
                    c0 += float(d @ Qm_ @ d)
                if wabs is not None:
                    for j, wj in enumerate(np.asarray(wabs, dtype=np.float64)):
                        if wj == 0.0:
                            continue
                        if np.any(cv[j]):
                            atoms.append((d[j], cv[j]))
                            cost_atoms.append((wj, len(atoms) - 1))
                        else:
                            c0 += wj * abs(d[j])

        nt = len(atoms)
        ny = nv + nt
        Gm = np.zeros((len(rows_G) + 2 * nt, ny))
        hm = np.zeros(len(rows_G) + 2 * nt)
        for i, (av, idx, tw) in enumerate(rows_G):
            Gm[i, :nv] = av
            for ti, w in zip(idx, tw):
                Gm[i, nv + ti] += w
            hm[i] = rows_h[i]
        r0 = len(rows_G)
        for i, (cst, av) in enumerate(atoms):           #  c + a v <= t ; -(c + a v) <= t
            Gm[r0 + 2 * i, :nv], Gm[r0 + 2 * i, nv + i], hm[r0 + 2 * i] = av, -1.0, -cst
            Gm[r0 + 2 * i + 1, :nv], Gm[r0 + 2 * i + 1, nv + i], hm[r0 + 2 * i + 1] = -av, -1.0, cst
        Pf = np.zeros((ny, ny))
        Pf[:nv, :nv] = P
        qf = np.r_[q, np.zeros(nt)]
        for wj, ti in cost_atoms:
            qf[nv + ti] += wj
        return Pf, qf, c0, Gm, hm, param_ok

    def evaluate_tube(self, xbar0: np.ndarray, e0: np.ndarray, v: np.ndarray, k: int = 1) -> np.ndarray:
        """Numeric Ze[k].Z at (xbar0, e0, v) -- what `Ze[1].Z.value` returns (examples/2.pulley_sim.py:96)."""
        w = np.r_[1.0, np.asarray(v, dtype=np.float64).reshape(-1), xbar0, e0]
        return self._Ze_aff[k] @ w

    # tzddpc/tzddpc.py:357-377
    def solve_status(self, xbar0: np.ndarray, e0: np.ndarray, check_feasibility: bool = True) -> SolveResult:
        n, m, N = self.dim_x, self.dim_u, self._N
        xbar0 = np.asarray(xbar0, dtype=np.float64).reshape(n)
        e0 = np.asarray(e0, dtype=np.float64).reshape(n)
        P, q, c0, G, h, param_ok = self._assemble(xbar0, e0)
        nv = N * m
        bad = SolveResult(2, np.inf, np.full((N, m), np.nan), np.full((N + 1, n), np.nan), np.zeros((n, 0)), [])
        if not param_ok:
            return bad
        if check_feasibility and G.shape[0]:
            from scipy.optimize import linprog
            fr = linprog(np.zeros(G.shape[1]), A_ub=G, b_ub=h + 1e-9 * np.maximum(1.0, np.abs(h)),
                         bounds=[(None, None)] * G.shape[1], method="highs")
            if fr.status == 2:
                return bad
        y, _, info = solve_qp_ipm(P, q, G, h)
        status = 0 if info["status"] == "optimal" else 1
        v = y[:nv].reshape(N, m)
        w = np.r_[1.0, y[:nv], xbar0, e0]
        xbar = self._xbar_aff @ w
        cost = float(0.5 * y @ P @ y + q @ y + c0)
        Ze_all = [Zc @ w for Zc in self._Ze_aff]
        return SolveResult(status, cost, v, xbar, Ze_all[1], Ze_all)

    def solve(self, xbar0: np.ndarray, e0: np.ndarray):
        r = self.solve_status(xbar0, e0)
        if np.isinf(r.cost):
            raise Exception("Problem is unbounded")                          # :374-375
        return r.cost, r.v, r.xbar, r.Ze1

    # examples/2.pulley_sim.py:81-96 ; examples/3.5dimsystem_sim.py:73-89
    def closed_loop(self, A_true: np.ndarray, B_true: np.ndarray, x0: np.ndarray, noise: np.ndarray,
                    keep_tubes: bool = False):
        """noise: (steps, n), the w_t realisations (an INPUT: the RNG stream of the absent
        pyzonotope is not reproducible).  Returns dict of trajectories."""
        n, m = self.dim_x, self.dim_u
        K = self.theta.K
        steps = noise.shape[0]
        x = np.zeros((steps + 1, n))
        xbar = np.zeros((steps + 1, n))
        e = np.zeros((steps + 1, n))
        u = np.zeros((steps, m))
        v0 = np.zeros((steps, m))
        cost = np.zeros(steps)
        status = np.zeros(steps, dtype=np.int32)
        tubes = []
        x[0] = x0
        xbar[0] = x0                                                         # :71-72
        for t in range(steps):
            r = self.solve_status(xbar[t], e[t])                             # :82-86
            status[t] = r.status
            if r.status == 2:
                x[t + 1:], xbar[t + 1:], e[t + 1:] = np.nan, np.nan, np.nan
                break
            cost[t] = r.cost
            v0[t] = r.v[0]
            xbar[t + 1] = r.xbar[1]                                          # :90
            u[t] = K @ e[t] + r.v[0]                                         # :91
            x[t + 1] = A_true @ x[t] + B_true @ u[t] + noise[t]              # :92
            e[t + 1] = x[t + 1] - xbar[t + 1]                                # :94
            if keep_tubes:
                tubes.append(r.Ze1)                                          # :96
        return {"x": x, "xbar": xbar, "e": e, "u": u, "v0": v0, "cost": cost, "status": status, "tubes": tubes}


def generate_trajectories(A: np.ndarray, B: np.ndarray, X0: Zonotope, U: Zonotope, W: Zonotope,
                          num_trajectories: int, num_steps: int, rng) -> Data:
    """examples/utils.py:6-45, including quirk Q9 (the returned first row is the origin: Y[j,0]=0)
    and the data noise being a uniformly random VERTEX of W (`:37`)."""
    n, m = B.shape
    total = num_steps * num_trajectories
    u = U.sample(total, rng).reshape(num_trajectories, num_steps, m)
    Wv = W.compute_vertices()
    X = np.zeros((num_trajectories, num_steps, n))
    Y = np.zeros((num_trajectories, num_steps, n))
    for j in range(num_trajectories):
        X[j, 0] = X0.sample(1, rng)[0]
        for i in range(1, num_steps):
            X[j, i] = A @ X[j, i - 1] + B @ u[j, i - 1] + Wv[rng.integers(len(Wv))]
            Y[j, i] = X[j, i]
    return Data(u.reshape(total, m), Y.reshape(total, n))
