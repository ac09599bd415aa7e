"""Host-side canonicalisation of the TZDDPC per-step program (runs once per build_problem).

The reference builds its convex program symbolically with cvxpy and lets cvxpy
canonicalise it once; every closed-loop step then only updates two parameters
(`tzddpc/tzddpc.py:132-241` build, `:357-377` solve).  This module is the
B200-side counterpart of that one-off canonicalisation: it turns the model
(M_Sigma centre, boxed M_K / M_Delta, K, W, X, U), the horizon and the stage cost
into a small *parametric* QP in OSQP form whose matrices depend on the model
only,

    minimise   0.5 z'P z + (q0 + Qp p)'z + c0(p)
    subject to l0 + r(p) <= A z <= u0 + r(p),     r(p) = R [1; p; alpha(p)],
               alpha_j(p) = |Bt_j p + gam_j|,

with p = [xbar0; e0] the two per-step parameters of the reference (`:155-157`),
z = [v; t] (nominal inputs and one epigraph variable per *distinct*
decision-dependent |.| atom).  The CUDA kernels (csrc/closed_loop.cu,
csrc/qp_admm.cu) evaluate r(p), run ADMM on the fixed (P, A) and materialise
Ze[1].Z from the sparse term table built here.  No numeric per-step quantity is
computed on the host.

Quirks of the reference that are kept on purpose: Q4 (Ze[N] is
never constrained), Q7 (`build_loss(u, xbar)` gets a FREE u, and xbar_0..xbar_{N-1}),
Q8 (noise recursion `range(1, k)`), and `build_problem_simplified`'s k0 window.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import itertools

import numpy as np

INF = float("inf")


@dataclass
class StageCost:
    """Structured stage cost (the reference takes a cvxpy callback, `tzddpc/tzddpc.py:135,222`):
    sum_i (x_i-x_ref)'Q(x_i-x_ref) + w_abs.|x_i-x_ref| + (u_i-u_ref)'R(u_i-u_ref) + r_abs.|u_i-u_ref|."""
    Q: Optional[np.ndarray] = None
    x_ref: Optional[np.ndarray] = None
    w_abs: Optional[np.ndarray] = None
    R: Optional[np.ndarray] = None
    u_ref: Optional[np.ndarray] = None
    r_abs: Optional[np.ndarray] = None


@dataclass
class BoxConstraint:
    """User constraints lo <= xbar[:, j] <= hi, lo <= v[:, j] <= hi (examples/3.5dimsystem_sim.py:23-26)."""
    x_lo: Optional[np.ndarray] = None
    x_hi: Optional[np.ndarray] = None
    v_lo: Optional[np.ndarray] = None
    v_hi: Optional[np.ndarray] = None


@dataclass
class TubeModel:
    """Everything `build_problem` reads from the controller (`tzddpc/tzddpc.py:163,175-176,191-197`)."""
    AB: np.ndarray          # n x (n+m)  centre of Mdata  = [A_hat B_hat]
    Acl: np.ndarray         # n x n      centre of MdataK = A_hat + B_hat K
    GK: np.ndarray          # NK x n x n       generators of MdataK
    GD: np.ndarray          # ND x n x (n+m)   generators of Mdelta (centre is zero)
    K: np.ndarray           # m x n
    WZ: np.ndarray          # n x (1+gW)  [c_W, G_W]
    X_lo: np.ndarray
    X_hi: np.ndarray
    U_lo: np.ndarray
    U_hi: np.ndarray

    @property
    def n(self) -> int:
        return self.AB.shape[0]

    @property
    def m(self) -> int:
        return self.AB.shape[1] - self.AB.shape[0]


@dataclass
class CompiledProgram:
    n: int
    m: int
    N: int
    nv: int
    nt: int
    nz: int
    nc: int
    npar: int
    na: int
    g1: int
    P: np.ndarray
    q0: np.ndarray
    Qp: np.ndarray
    A: np.ndarray
    l0: np.ndarray
    u0: np.ndarray
    kink0: np.ndarray       # nc: |.|-cost rows carry  wabs_i |(A z)_i - (kink0_i + r_i(p))|
    wabs: np.ndarray
    R: np.ndarray           # nc x (1+npar+na)
    Bt: np.ndarray          # na x npar
    gam: np.ndarray         # na
    Rchk: np.ndarray        # nchk x (1+npar+na): feasible iff Rchk [1;p;alpha] <= 0
    cc: np.ndarray          # 1+npar+na : linear part of the cost constant
    CC2: np.ndarray         # npar x npar: quadratic part of the cost constant
    XB: np.ndarray          # (N+1)n x (1+nv+npar): xbar trajectory map
    ze1_ptr: np.ndarray     # n(1+g1)+1  CSR over the entries of Ze[1].Z, row-major (r, j)
    ze1_idx: np.ndarray     # term -> index into w = [1, v, p]
    ze1_val: np.ndarray
    gens_per_step: List[int] = field(default_factory=list)     # what `tzddpc/tzddpc.py:206` prints
    # Ruiz equilibration (model-only): zbar = z / D, Abar = E A D, Pbar = c D P D
    D: np.ndarray = None
    E: np.ndarray = None
    c: float = 1.0
    wmax: float = 1.0       # largest |cost weight|: scale of the cost tolerance


# ------------------------------------------------------------------------------------------------------------------
# Every numeric tensor below carries a leading DATA-SET axis (size D): `compile_program_batch` canonicalises D models of
# one structure (same dimensions, horizon, cost, constraints) at once -- the structural decisions (which generators are
# zero, which atoms are proportional, which rows coincide) are taken once, over the whole batch, and the arithmetic is
# element-wise numpy in a fixed order, so slice d of the batch equals the canonicalisation of model d alone BIT FOR BIT
# (tests/test_program_host.py).  `compile_program` is the batch of one.
# ------------------------------------------------------------------------------------------------------------------
def _contract(M: np.ndarray, Zc: np.ndarray) -> np.ndarray:
    """out[d, r, j, c] = sum_k M[d, r, k] Zc[d, k, j, c], accumulated k = 0, 1, ... (explicit order: einsum / BLAS pick their
    summation order by shape, which would make the result depend on the batch size)."""
    out = M[:, :, 0, None, None] * Zc[:, None, 0]
    for k in range(1, M.shape[2]):
        out = out + M[:, :, k, None, None] * Zc[:, None, k]
    return out


def _mz_times(C: np.ndarray, G: np.ndarray, Zc: np.ndarray) -> np.ndarray:
    """MatrixZonotope x Zonotope on coefficient tensors Zc[d, row, column, coef]:
    [C Z, G_1 Z, ..., G_N Z] (pyzonotope column order; call sites
    tzddpc/tzddpc.py:175-176,181,185).  Zero columns are retained.  C: (D, n, p), G: (D, NG, n, p)."""
    D, n = C.shape[0], C.shape[1]
    w = Zc.shape[2]
    out = np.empty((D, n, w * (1 + G.shape[1]), Zc.shape[3]))
    out[:, :, :w] = _contract(C, Zc)
    for i in range(G.shape[1]):
        out[:, :, (i + 1) * w:(i + 2) * w] = _contract(G[:, i], Zc)
    return out


def _plus(Zc: np.ndarray, Zo: np.ndarray) -> np.ndarray:
    """Minkowski sum on coefficient tensors: [c1 + c2, G1, G2] (tzddpc/tzddpc.py:176,185,205)."""
    out = np.concatenate([Zc, Zo[:, :, 1:]], axis=2)
    out[:, :, 0] = Zc[:, :, 0] + Zo[:, :, 0]
    return out


def _dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """sum over the last axis, accumulated left to right (batch-size independent)."""
    out = a[..., 0] * b[..., 0]
    for k in range(1, a.shape[-1]):
        out = out + a[..., k] * b[..., k]
    return out


class _Forms:
    """Registry of unit-normalised affine forms (D x nw each); forms that are proportional in EVERY data set share one entry."""

    def __init__(self, nw: int):
        self.nw = nw
        self.forms: List[np.ndarray] = []
        self._index = {}

    def add(self, f: np.ndarray) -> Tuple[int, np.ndarray]:
        """f: (D, nw) -> (index, weight (D,)) with f = +-weight * forms[index]."""
        nrm = np.sqrt(_dot(f, f))
        absmax = np.max(np.abs(f), axis=0)                         # over the batch: the structure is the union
        big = np.flatnonzero(absmax > 1e-9 * np.max(absmax))
        sgn = np.where(f[:, big[0]] > 0, 1.0, -1.0)
        u = sgn[:, None] * f / np.where(nrm > 0, nrm, 1.0)[:, None]
        key = tuple(np.round(u[0], 9) + 0.0)
        i = self._index.get(key)
        if i is not None and not np.allclose(self.forms[i], u, rtol=0.0, atol=1e-9):
            i = None                                               # proportional in data set 0 only: keep them apart
        if i is None:
            i = len(self.forms)
            self._index.setdefault(key, i)
            self.forms.append(u)
        return i, nrm


@dataclass
class TubeModelBatch:
    """D models of one structure (the fields of TubeModel with a leading data-set axis); the sets are shared."""
    AB: np.ndarray          # D x n x (n+m)
    Acl: np.ndarray         # D x n x n
    GK: np.ndarray          # D x NK x n x n
    GD: np.ndarray          # D x ND x n x (n+m)
    K: np.ndarray           # D x m x n
    WZ: np.ndarray          # n x (1+gW)
    X_lo: np.ndarray
    X_hi: np.ndarray
    U_lo: np.ndarray
    U_hi: np.ndarray

    @classmethod
    def of(cls, models: List[TubeModel]) -> "TubeModelBatch":
        m0 = models[0]
        st = lambda name: np.stack([np.asarray(getattr(m_, name), dtype=np.float64) for m_ in models])    # noqa: E731
        return cls(st("AB"), st("Acl"), st("GK"), st("GD"), st("K").reshape(len(models), -1, m0.n), np.asarray(m0.WZ, dtype=np.float64),
                   m0.X_lo, m0.X_hi, m0.U_lo, m0.U_hi)

    @classmethod
    def boxed(cls, AB: np.ndarray, dAB: np.ndarray, dK: np.ndarray, K: np.ndarray, WZ, X_lo, X_hi, U_lo, U_hi) -> "TubeModelBatch":
        """From the outputs of tz_identify (order-1 reduced model, tzddpc/tzddpc.py:119-128): centre AB, boxes dAB of M_Delta and
        dK of M_K; the single-entry generators d[r,c] E_rc in Girard's diag order (zonotope.boxed_generators)."""
        D, n, d = AB.shape
        K = np.asarray(K, dtype=np.float64).reshape(D, -1, n)
        GK = np.zeros((D, n * n, n, n))
        GD = np.zeros((D, n * d, n, d))
        for r in range(n):
            for c in range(n):
                GK[:, r * n + c, r, c] = dK[:, r, c]
            for c in range(d):
                GD[:, r * d + c, r, c] = dAB[:, r, c]
        Acl = AB[:, :, :n].copy()
        for k in range(K.shape[1]):                                 # A_hat + B_hat K, accumulated in order (as MatrixZonotope * [I; K])
            Acl = Acl + AB[:, :, n + k, None] * K[:, None, k, :]
        return cls(AB, Acl, GK, GD, K, np.asarray(WZ, dtype=np.float64), X_lo, X_hi, U_lo, U_hi)

    @property
    def n(self) -> int:
        return self.AB.shape[1]

    @property
    def m(self) -> int:
        return self.AB.shape[2] - self.AB.shape[1]

    @property
    def D(self) -> int:
        return self.AB.shape[0]


@dataclass
class CompiledProgramBatch:
    """D compiled programs of one structure: the numeric fields of CompiledProgram with a leading data-set axis."""
    n: int
    m: int
    N: int
    nv: int
    nt: int
    nz: int
    nc: int
    npar: int
    na: int
    g1: int
    P: np.ndarray
    q0: np.ndarray
    Qp: np.ndarray
    A: np.ndarray
    l0: np.ndarray
    u0: np.ndarray
    kink0: np.ndarray
    wabs: np.ndarray
    R: np.ndarray
    Bt: np.ndarray
    gam: np.ndarray
    Rchk: np.ndarray
    cc: np.ndarray
    CC2: np.ndarray
    XB: np.ndarray
    ze1_ptr: np.ndarray
    ze1_idx: np.ndarray
    ze1_val: np.ndarray
    gens_per_step: List[int]
    D_: np.ndarray
    E: np.ndarray
    c: np.ndarray
    wmax: float

    @property
    def num(self) -> int:
        return self.P.shape[0]

    def program(self, d: int) -> CompiledProgram:
        prog = CompiledProgram(n=self.n, m=self.m, N=self.N, nv=self.nv, nt=self.nt, nz=self.nz, nc=self.nc, npar=self.npar, na=self.na,
                               g1=self.g1, P=self.P[d], q0=self.q0[d], Qp=self.Qp[d], A=self.A[d], l0=self.l0[d], u0=self.u0[d],
                               kink0=self.kink0[d], wabs=self.wabs[d], R=self.R[d], Bt=self.Bt[d], gam=self.gam[d], Rchk=self.Rchk[d],
                               cc=self.cc[d], CC2=self.CC2[d], XB=self.XB[d], ze1_ptr=self.ze1_ptr, ze1_idx=self.ze1_idx,
                               ze1_val=self.ze1_val[d], gens_per_step=list(self.gens_per_step), wmax=self.wmax)
        prog.D, prog.E, prog.c = self.D_[d], self.E[d], float(self.c[d])
        return prog


def compile_program(model: TubeModel, horizon: int, cost: StageCost, box: Optional[BoxConstraint] = None,
                    k0: Optional[int] = None, zero_tol: float = 0.0, expand_max_atoms: int = 1) -> CompiledProgram:
    return compile_program_batch(TubeModelBatch.of([model]), horizon, cost, box, k0, zero_tol, expand_max_atoms).program(0)


def compile_program_batch(model: TubeModelBatch, horizon: int, cost: StageCost, box: Optional[BoxConstraint] = None,
                          k0: Optional[int] = None, zero_tol: float = 0.0, expand_max_atoms: int = 1) -> CompiledProgramBatch:
    n, m, N, D = model.n, model.m, int(horizon), model.D
    assert N >= 1
    box = box or BoxConstraint()
    simplified = k0 is not None
    nv, npar = N * m, 2 * n
    nw = 1 + nv + npar
    ix0, ie0 = 1 + nv, 1 + nv + n
    A_hat, B_hat = model.AB[:, :, :n], model.AB[:, :, n:]

    def anyb(x: np.ndarray) -> np.ndarray:
        """structural test over the batch: non-zero in ANY data set"""
        return np.any(x != 0.0, axis=0)

    # ---- nominal trajectory, eliminated through the equality constraints (:166-170) ----
    xbar = np.zeros((D, N + 1, n, nw))
    xbar[:, 0, np.arange(n), ix0 + np.arange(n)] = 1.0
    vaff = np.zeros((D, N, m, nw))
    for k in range(N):
        vaff[:, k, np.arange(m), 1 + k * m + np.arange(m)] = 1.0
        xbar[:, k + 1] = _contract(A_hat, xbar[:, k, :, None, :])[:, :, 0] + _contract(B_hat, vaff[:, k, :, None, :])[:, :, 0]

    # ---- error tubes (:172-186, simplified :283-302) ----
    Wc = np.zeros((D,) + model.WZ.shape + (nw,))
    Wc[..., 0] = model.WZ
    ze0 = np.zeros((D, n, 2, nw))
    ze0[:, np.arange(n), 0, ie0 + np.arange(n)] = 1.0
    Ze = [ze0]
    XU = []
    for k in range(N):
        z = np.zeros((D, n + m, 2, nw))
        z[:, :n, 0], z[:, n:, 0] = xbar[:, k], vaff[:, k]
        XU.append(z)
    zeroC = np.zeros((D, n, n + m))
    T1 = [_mz_times(model.Acl, model.GK, Ze[0])]
    Zn = [_plus(_mz_times(zeroC, model.GD, XU[k]), Wc) for k in range(N)]
    T2 = []
    for k in range(N):
        if not simplified:
            # T1[k+1] is only read when Ze[k+2] is built, i.e. for k + 2 < N (Ze[N] is never built, Q4)
            T1.append(_mz_times(model.Acl, model.GK, T1[-1]) if k + 2 < N else None)
            noise = Zn[0]
            for j in range(1, k):                                                      # Q8
                noise = _plus(_mz_times(model.Acl, model.GK, noise), Zn[j])
        else:
            T1.append(T1[-1] if k > k0 else _mz_times(model.Acl, model.GK, T1[-1]))
            start = max(0, k - k0)
            noise = Zn[start]
            for j in range(1, min(k, k0)):
                noise = _plus(_mz_times(model.Acl, model.GK, noise), Zn[start + j])
        T2.append(noise)

    # ---- abs-sum rows: centre_r +- sum_j |gen_rj| within [lo_r, hi_r] (:191-199) ----
    tubes = []      # (centre (D,d,nw), gens (D,d,g,nw), lo (d), hi (d))
    gens_per_step = []
    width = lambda Z: Z.shape[2]        # noqa: E731
    t1_width = [width(T1[0])]
    for k in range(1, N):
        # (the count `:206` prints needs the width of T1[k] even where the tensor itself is not needed)
        t1_width.append(width(T1[k]) if T1[k] is not None else t1_width[-1] * (1 + model.GK.shape[1]))
    for k in range(N):
        Zk = Ze[-1]
        tubes.append((Zk[:, :, 0] + xbar[:, k], Zk[:, :, 1:], model.X_lo, model.X_hi))
        KZ = _contract(model.K, Zk)
        tubes.append((KZ[:, :, 0] + vaff[:, k], KZ[:, :, 1:], model.U_lo, model.U_hi))
        gens_per_step.append(t1_width[k] + width(T2[k]) - 2)          # the count `:206` prints
        if k + 1 < N or k == 0:          # Ze[N] is unconstrained (Q4); Ze[1] is always returned (:377)
            Ze.append(_plus(T1[k], T2[k]))
    ze1 = Ze[1]
    g1 = ze1.shape[2] - 1

    vsl = slice(1, 1 + nv)
    psl = np.r_[0, np.arange(1 + nv, nw)]    # [1; p] slots of w
    dforms = _Forms(nw)      # decision-dependent atoms -> epigraph variables
    pforms = _Forms(nw)      # parameter-only atoms     -> alpha
    # raw rows:  sense=+1:  cv.v + sum_i tw_i|dform_i| + centre(p) + sum_j aw_j alpha_j + cst <= bound
    #            sense=-1:  cv.v - sum_i tw_i|dform_i| + centre(p) - sum_j aw_j alpha_j - cst >= bound
    raw_rows = []
    zD = np.zeros(D)

    def absorb(gens_r: np.ndarray):
        """gens_r: (D, g, nw)"""
        tw, aw, cst = {}, {}, zD.copy()
        amax = np.max(np.abs(gens_r), axis=(0, 2)) if gens_r.shape[1] else np.zeros(0)
        for j in np.flatnonzero(amax > zero_tol):
            f = gens_r[:, j]
            if np.any(anyb(f[:, vsl])):
                i, wgt = dforms.add(f)
                tw[i] = tw.get(i, 0.0) + wgt
            elif np.any(anyb(f[:, 1 + nv:])):
                i, wgt = pforms.add(f)
                aw[i] = aw.get(i, 0.0) + wgt
            else:
                cst = cst + np.abs(f[:, 0])
        return tw, aw, cst

    for centre, gens, lo, hi in tubes:
        for r in range(centre.shape[1]):
            tw, aw, cst = absorb(gens[:, r])
            if np.isfinite(hi[r]):
                raw_rows.append((centre[:, r, vsl].copy(), tw, +1, float(hi[r]), centre[:, r][:, psl].copy(), aw, cst))
            if np.isfinite(lo[r]):
                raw_rows.append((centre[:, r, vsl].copy(), tw, -1, float(lo[r]), centre[:, r][:, psl].copy(), aw, cst))
    xrows = range(1, N + 1) if simplified else range(N + 1)
    for k in xrows:
        for j in range(n):
            for lim, sense in ((box.x_hi, +1), (box.x_lo, -1)):
                if lim is not None and np.isfinite(lim[j]):
                    raw_rows.append((xbar[:, k, j, vsl].copy(), {}, sense, float(lim[j]), xbar[:, k, j][:, psl].copy(), {}, zD.copy()))
    for k in range(N):
        for j in range(m):
            for lim, sense in ((box.v_hi, +1), (box.v_lo, -1)):
                if lim is not None and np.isfinite(lim[j]):
                    raw_rows.append((vaff[:, k, j, vsl].copy(), {}, sense, float(lim[j]), vaff[:, k, j][:, psl].copy(), {}, zD.copy()))

    # rows whose few decision-dependent atoms are expanded into sign combinations:
    #   a.v + w|f| <= b   <=>   a.v + w f <= b  and  a.v - w f <= b      (exact; no epigraph variable)
    expanded = []
    used_forms = set()
    for cv, tw, sense, bound, cpar, aw, cst in raw_rows:
        if 0 < len(tw) <= expand_max_atoms:
            items = list(tw.items())
            for signs in itertools.product((1.0, -1.0), repeat=len(items)):
                cv2, cpar2 = cv.copy(), cpar.copy()
                for (i, wgt), sg in zip(items, signs):
                    f = dforms.forms[i]
                    cv2 += (sense * sg * wgt)[:, None] * f[:, vsl]
                    cpar2 += (sense * sg * wgt)[:, None] * f[:, psl]
                expanded.append((cv2, {}, sense, bound, cpar2, aw, cst))
        else:
            used_forms.update(tw.keys())
            expanded.append((cv, tw, sense, bound, cpar, aw, cst))
    remap = {old: new for new, old in enumerate(sorted(used_forms))}
    tforms = [dforms.forms[i] for i in sorted(used_forms)]
    nt, nz = len(tforms), nv + len(tforms)

    # ---- objective (Q7) ----
    Pv = np.zeros((D, nv, nv))
    qv = np.zeros((D, nv, 1 + npar))         # q = qv [1; p]
    cost_rows = []                           # (cv, cpar incl. -ref, weight): weight*|cv.v + cpar.[1;p]|
    cost_a = {}                              # alpha index -> weight
    cc_lin = np.zeros((D, 1 + npar))
    CC2 = np.zeros((D, npar, npar))
    wmax = 0.0
    if simplified:
        xcost = [xbar[:, k] for k in range(1, N + 1)]
        ucost = [vaff[:, k] for k in range(N)]
    else:
        xcost = [xbar[:, k] for k in range(N)]
        ucost = []

    def quad(L_: np.ndarray, Qm_: np.ndarray, R_: np.ndarray) -> np.ndarray:
        """L' Q R per data set: L (D, dim, a), R (D, dim, b) -> (D, a, b), accumulated in a fixed order."""
        out = np.zeros((D, L_.shape[2], R_.shape[2]))
        for i in range(Qm_.shape[0]):
            for j in range(Qm_.shape[1]):
                if Qm_[i, j] != 0.0:
                    out = out + Qm_[i, j] * (L_[:, i, :, None] * R_[:, j, None, :])
        return out

    for rows, Qm, ref, wabs, dim in ((xcost, cost.Q, cost.x_ref, cost.w_abs, n), (ucost, cost.R, cost.u_ref, cost.r_abs, m)):
        ref = np.zeros(dim) if ref is None else np.asarray(ref, dtype=np.float64).reshape(dim)
        for aff in rows:
            cv, cp = aff[:, :, vsl], aff[:, :, psl].copy()
            cp[:, :, 0] -= ref
            if Qm is not None:
                Qm_ = np.asarray(Qm, dtype=np.float64)
                wmax = max(wmax, float(np.max(np.abs(Qm_))))
                Pv += 2.0 * quad(cv, Qm_, cv)
                qv += 2.0 * quad(cv, Qm_, cp)
                M2 = quad(cp, Qm_, cp)                # [1;p]' M2 [1;p]
                cc_lin[:, 0] += M2[:, 0, 0]
                cc_lin[:, 1:] += M2[:, 0, 1:] + M2[:, 1:, 0]
                CC2 += M2[:, 1:, 1:]
            if wabs is not None:
                for j, wj in enumerate(np.asarray(wabs, dtype=np.float64).reshape(dim)):
                    if wj == 0.0:
                        continue
                    assert wj > 0.0, "negative |.| weights are not convex"
                    wmax = max(wmax, float(wj))
                    if np.any(anyb(cv[:, j])):
                        cost_rows.append((cv[:, j].copy(), cp[:, j].copy(), float(wj), float(ref[j])))
                    elif np.any(anyb(cp[:, j, 1:])):
                        f = np.zeros((D, nw))
                        f[:, psl] = cp[:, j]
                        i, wgt = pforms.add(f)
                        cost_a[i] = cost_a.get(i, 0.0) + wj * wgt
                    else:
                        cc_lin[:, 0] += wj * np.abs(cp[:, j, 0])

    na = len(pforms.forms)
    ncol = 1 + npar + na

    def shift_of(sense: int, cpar: np.ndarray, aw: dict, cst: np.ndarray) -> np.ndarray:
        """hi-row: a z <= hi - centre(p) - sum h alpha - cst ;  lo-row: a z >= lo - centre(p) + sum h alpha + cst."""
        r = np.zeros((D, ncol))
        r[:, :1 + npar] = -cpar
        r[:, 0] -= sense * cst
        for i, wgt in aw.items():
            r[:, 1 + npar + i] -= sense * wgt
        return r

    # rows: [a (D,nz), l0, u0, kink0, wabs, shift (D,ncol)] ; merged when (a, shift) coincide in every data set
    rows, index, chk = [], {}, []

    def add_row(a, lo_, hi_, r, kink=0.0, w=0.0):
        key = (tuple(np.round(a[0], 12) + 0.0), tuple(np.round(r[0], 12) + 0.0))
        i = index.get(key)
        if i is not None and not (np.allclose(rows[i][0], a, rtol=0.0, atol=1e-12) and np.allclose(rows[i][5], r, rtol=0.0, atol=1e-12)):
            i = None                                     # coincide in data set 0 only
        if i is None:
            index.setdefault(key, len(rows))
            rows.append([a, lo_, hi_, kink, w, r])
        else:
            rows[i][1] = max(rows[i][1], lo_)
            rows[i][2] = min(rows[i][2], hi_)
            if w > 0.0:
                assert rows[i][4] == 0.0 or rows[i][3] == kink
                rows[i][3], rows[i][4] = kink, rows[i][4] + w

    for cv, tw, sense, bound, cpar, aw, cst in expanded:
        a = np.zeros((D, nz))
        a[:, :nv] = cv
        for i, wgt in tw.items():
            a[:, nv + remap[i]] = sense * wgt       # hi: +sum w t ; lo: a z - sum w t >= ...
        r = shift_of(sense, cpar, aw, cst)
        if not np.any(anyb(a)):
            # parameter-only row (e.g. xbar0 + e0 in X at k = 0): a feasibility check, 0 within [l, u]
            rr = r.copy()
            rr[:, 0] += bound
            chk.append(-rr if sense > 0 else rr)      # hi: 0 <= bound + r  ;  lo: 0 >= bound + r
            continue
        add_row(a, -INF if sense > 0 else bound, bound if sense > 0 else INF, r)
    # |.| cost rows: handled by the prox of  w|z - kink|  in the ADMM z-update (no slack variable)
    for cv, cp, wj, refj in cost_rows:
        a = np.zeros((D, nz))
        a[:, :nv] = cv
        r = np.zeros((D, ncol))
        r[:, :1 + npar] = -cp
        r[:, 0] -= refj                               # cp already has -ref folded in: undo, keep ref as the kink
        add_row(a, -INF, INF, r, kink=refj, w=wj)
    # epigraph rows of the remaining atoms:  f(v,p) - t <= 0  and  f(v,p) + t >= 0
    for i, f in enumerate(tforms):
        for sgn in (-1.0, +1.0):
            a = np.zeros((D, nz))
            a[:, :nv] = f[:, vsl]
            a[:, nv + i] = sgn
            r = np.zeros((D, ncol))
            r[:, :1 + npar] = -f[:, psl]
            add_row(a, -INF if sgn < 0 else 0.0, 0.0 if sgn < 0 else INF, r)
    nc = len(rows)
    A = np.stack([r_[0] for r_ in rows], axis=1).reshape(D, nc, nz) if nc else np.zeros((D, 0, nz))
    l0 = np.tile(np.asarray([r_[1] for r_ in rows], dtype=np.float64), (D, 1))
    u0 = np.tile(np.asarray([r_[2] for r_ in rows], dtype=np.float64), (D, 1))
    kink0 = np.tile(np.asarray([r_[3] for r_ in rows], dtype=np.float64), (D, 1))
    wabs_rows = np.tile(np.asarray([r_[4] for r_ in rows], dtype=np.float64), (D, 1))
    R = np.stack([r_[5] for r_ in rows], axis=1).reshape(D, nc, ncol) if nc else np.zeros((D, 0, ncol))

    P = np.zeros((D, nz, nz))
    P[:, :nv, :nv] = Pv
    q0 = np.zeros((D, nz))
    Qp = np.zeros((D, nz, npar))
    q0[:, :nv], Qp[:, :nv] = qv[:, :, 0], qv[:, :, 1:]
    cc = np.zeros((D, ncol))
    cc[:, :1 + npar] = cc_lin
    for i, wgt in cost_a.items():
        cc[:, 1 + npar + i] += wgt
    Bt = np.stack([f[:, 1 + nv:] for f in pforms.forms], axis=1).reshape(D, na, npar) if na else np.zeros((D, 0, npar))
    gam = np.stack([f[:, 0] for f in pforms.forms], axis=1).reshape(D, na) if na else np.zeros((D, 0))
    Rchk = np.stack(chk, axis=1).reshape(D, len(chk), ncol) if chk else np.zeros((D, 0, ncol))

    XB = xbar.reshape(D, (N + 1) * n, nw)
    # sparse term table of Ze[1].Z (row-major over (r, j)); pattern = union over the batch
    flat = ze1.reshape(D, n * (1 + g1), nw)
    pat = anyb(flat)
    ptr = np.zeros(flat.shape[1] + 1, dtype=np.int32)
    idx = []
    for e in range(flat.shape[1]):
        nzc = np.flatnonzero(pat[e])
        idx.extend(nzc.tolist())
        ptr[e + 1] = len(idx)
    ent_of = np.repeat(np.arange(flat.shape[1]), np.diff(ptr))
    idx = np.asarray(idx, dtype=np.int32)
    val = flat[:, ent_of, idx] if len(idx) else np.zeros((D, 0))
    Dsc, Esc, csc = ruiz_equilibrate(P, A, q0, Qp, wabs_rows)
    return CompiledProgramBatch(n=n, m=m, N=N, nv=nv, nt=nt, nz=nz, nc=nc, npar=npar, na=na, g1=g1,
                                P=P, q0=q0, Qp=Qp, A=A, l0=l0, u0=u0, kink0=kink0, wabs=wabs_rows, R=R, Bt=Bt, gam=gam, Rchk=Rchk,
                                cc=cc, CC2=CC2, XB=XB, ze1_ptr=ptr, ze1_idx=idx, ze1_val=np.ascontiguousarray(val),
                                gens_per_step=gens_per_step, D_=Dsc, E=Esc, c=csc, wmax=max(wmax, 1.0))


def ruiz_equilibrate(P: np.ndarray, A: np.ndarray, q0: np.ndarray, Qp: np.ndarray, wabs: np.ndarray, iters: int = 15):
    """Modified Ruiz equilibration of [[P, A'], [A, 0]] (OSQP, Stellato et al. 2020, Alg. 2), batched over the leading
    data-set axis.  Model-only, so it is done once here; the kernels scale q, l, u per scenario."""
    Dn, nz, nc = P.shape[0], P.shape[1], A.shape[1]
    Dv, Ev = np.ones((Dn, nz)), np.ones((Dn, nc))
    Pb, Ab = P.copy(), A.copy()
    for _ in range(iters):
        cn = np.maximum(np.max(np.abs(Pb), axis=1, initial=0.0), np.max(np.abs(Ab), axis=1, initial=0.0))
        rn = np.max(np.abs(Ab), axis=2, initial=0.0)
        d = 1.0 / np.sqrt(np.where(cn > 1e-12, cn, 1.0))
        e = 1.0 / np.sqrt(np.where(rn > 1e-12, rn, 1.0))
        Pb = d[:, :, None] * Pb * d[:, None, :]
        Ab = e[:, :, None] * Ab * d[:, None, :]
        Dv *= d
        Ev *= e
    qn = np.maximum(np.maximum(np.max(np.abs(Dv * q0), axis=1, initial=0.0), np.max(np.abs(Dv[:, :, None] * Qp), axis=(1, 2), initial=0.0)),
                    np.max(wabs / Ev, axis=1, initial=0.0))
    # (mean over the columns accumulated in order: np.mean's pairwise summation is batch-size independent here, nz is tiny)
    colmax = np.max(np.abs(Pb), axis=1) if nz else np.zeros((Dn, 0))
    pn = np.zeros(Dn)
    for j in range(nz):
        pn = pn + colmax[:, j]
    pn = pn / nz if nz else pn
    g = np.maximum(pn, qn)
    c = np.where(g > 1e-12, 1.0 / np.where(g > 1e-12, g, 1.0), 1.0)
    return Dv, Ev, c
