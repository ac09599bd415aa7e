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


def _mz_times(C: np.ndarray, G: np.ndarray, Zc: np.ndarray) -> np.ndarray:
    """MatrixZonotope x Zonotope on coefficient tensors Zc[row, column, coef]:
    [C Z, G_1 Z, ..., G_N Z] (pyzonotope column order; call sites
    tzddpc/tzddpc.py:175-176,181,185).  Zero columns are retained."""
    out = np.empty((C.shape[0], Zc.shape[1] * (1 + G.shape[0]), Zc.shape[2]))
    w = Zc.shape[1]
    out[:, :w] = np.einsum("rk,kjc->rjc", C, Zc)
    if G.shape[0]:
        out[:, w:] = np.einsum("irk,kjc->rijc", G, Zc).reshape(C.shape[0], -1, Zc.shape[2])
    return out


def _plus(Zc: np.ndarray, Zo: np.ndarray) -> np.ndarray:
    """Minkowski sum on coefficient tensors: [c1 + c2, G1, G2] (tzddpc/tzddpc.py:176,185,205)."""
    out = np.concatenate([Zc, Zo[:, 1:]], axis=1)
    out[:, 0] = Zc[:, 0] + Zo[:, 0]
    return out


class _Forms:
    """Registry of unit-normalised affine forms; proportional forms share one entry."""

    def __init__(self, nw: int):
        self.nw = nw
        self.forms: List[np.ndarray] = []
        self._index = {}

    def add(self, f: np.ndarray) -> Tuple[int, float]:
        """-> (index, weight) with f = +-weight * forms[index]."""
        nrm = float(np.sqrt(f @ f))
        big = np.flatnonzero(np.abs(f) > 1e-9 * np.max(np.abs(f)))
        s = 1.0 if f[big[0]] > 0 else -1.0
        u = s * f / nrm
        key = tuple(np.round(u, 9) + 0.0)
        i = self._index.get(key)
        if i is None:
            i = len(self.forms)
            self._index[key] = i
            self.forms.append(u)
        return i, nrm


def compile_program(model: TubeModel, horizon: int, cost: StageCost, box: Optional[BoxConstraint] = None,
                    k0: Optional[int] = None, zero_tol: float = 0.0, expand_max_atoms: int = 1) -> CompiledProgram:
    n, m, N = model.n, model.m, int(horizon)
    assert N >= 1
    box = box or BoxConstraint()
    simplified = k0 is not None
    nv, npar = N * m, 2 * n
    nw = 1 + nv + npar
    ix0, ie0 = 1 + nv, 1 + nv + n
    A_hat, B_hat = model.AB[:, :n], model.AB[:, n:]

    # ---- nominal trajectory, eliminated through the equality constraints (:166-170) ----
    xbar = np.zeros((N + 1, n, nw))
    xbar[0, np.arange(n), ix0 + np.arange(n)] = 1.0
    vaff = np.zeros((N, m, nw))
    for k in range(N):
        vaff[k, np.arange(m), 1 + k * m + np.arange(m)] = 1.0
        xbar[k + 1] = A_hat @ xbar[k] + B_hat @ vaff[k]

    # ---- error tubes (:172-186, simplified :283-302) ----
    Wc = np.zeros(model.WZ.shape + (nw,))
    Wc[..., 0] = model.WZ
    ze0 = np.zeros((n, 2, nw))
    ze0[np.arange(n), 0, ie0 + np.arange(n)] = 1.0
    Ze = [ze0]
    XU = []
    for k in range(N):
        z = np.zeros((n + m, 2, nw))
        z[:n, 0], z[n:, 0] = xbar[k], vaff[k]
        XU.append(z)
    zeroC = np.zeros((n, n + m))
    T1 = [_mz_times(model.Acl, model.GK, Ze[0])]
    Zn = [_plus(_mz_times(zeroC, model.GD, XU[k]), Wc) for k in range(N)]
    T2 = []
    for k in range(N):
        if not simplified:
            T1.append(_mz_times(model.Acl, model.GK, T1[-1]) if k + 1 < N else None)   # T1[N] is never used
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
    tubes = []      # (centre (d,nw), gens (d,g,nw), lo (d), hi (d))
    gens_per_step = []
    for k in range(N):
        Zk = Ze[-1]
        tubes.append((Zk[:, 0] + xbar[k], Zk[:, 1:], model.X_lo, model.X_hi))
        KZ = np.einsum("ir,rjc->ijc", model.K, Zk)
        tubes.append((KZ[:, 0] + vaff[k], KZ[:, 1:], model.U_lo, model.U_hi))
        gens_per_step.append(T1[k].shape[1] + T2[k].shape[1] - 2)     # the count `:206` prints
        if k + 1 < N or k == 0:          # Ze[N] is unconstrained (Q4); Ze[1] is always returned (:377)
            Ze.append(_plus(T1[k], T2[k]))
    ze1 = Ze[1]
    g1 = ze1.shape[1] - 1

    vsl = slice(1, 1 + nv)
    psl = np.r_[0, np.arange(1 + nv, nw)]    # [1; p] slots of w
    dforms = _Forms(nw)      # decision-dependent atoms -> epigraph variables
    pforms = _Forms(nw)      # parameter-only atoms     -> alpha
    # raw rows:  sense=+1:  cv.v + sum_i tw_i|dform_i| + centre(p) + sum_j aw_j alpha_j + cst <= bound
    #            sense=-1:  cv.v - sum_i tw_i|dform_i| + centre(p) - sum_j aw_j alpha_j - cst >= bound
    raw_rows = []

    def absorb(gens_r: np.ndarray):
        tw, aw, cst = {}, {}, 0.0
        amax = np.max(np.abs(gens_r), axis=1) if gens_r.shape[0] else np.zeros(0)
        for j in np.flatnonzero(amax > zero_tol):
            f = gens_r[j]
            if np.any(f[vsl] != 0.0):
                i, wgt = dforms.add(f)
                tw[i] = tw.get(i, 0.0) + wgt
            elif np.any(f[1 + nv:] != 0.0):
                i, wgt = pforms.add(f)
                aw[i] = aw.get(i, 0.0) + wgt
            else:
                cst += abs(f[0])
        return tw, aw, cst

    for centre, gens, lo, hi in tubes:
        for r in range(centre.shape[0]):
            tw, aw, cst = absorb(gens[r])
            if np.isfinite(hi[r]):
                raw_rows.append((centre[r, vsl].copy(), tw, +1, float(hi[r]), centre[r, psl].copy(), aw, cst))
            if np.isfinite(lo[r]):
                raw_rows.append((centre[r, vsl].copy(), tw, -1, float(lo[r]), centre[r, psl].copy(), aw, cst))
    xrows = range(1, N + 1) if simplified else range(N + 1)
    for k in xrows:
        for j in range(n):
            for lim, sense in ((box.x_hi, +1), (box.x_lo, -1)):
                if lim is not None and np.isfinite(lim[j]):
                    raw_rows.append((xbar[k, j, vsl].copy(), {}, sense, float(lim[j]), xbar[k, j, psl].copy(), {}, 0.0))
    for k in range(N):
        for j in range(m):
            for lim, sense in ((box.v_hi, +1), (box.v_lo, -1)):
                if lim is not None and np.isfinite(lim[j]):
                    raw_rows.append((vaff[k, j, vsl].copy(), {}, sense, float(lim[j]), vaff[k, j, psl].copy(), {}, 0.0))

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
                    cv2 += sense * sg * wgt * f[vsl]
                    cpar2 += sense * sg * wgt * f[psl]
                expanded.append((cv2, {}, sense, bound, cpar2, aw, cst))
        else:
            used_forms.update(tw.keys())
            expanded.append((cv, tw, sense, bound, cpar, aw, cst))
    remap = {old: new for new, old in enumerate(sorted(used_forms))}
    tforms = [dforms.forms[i] for i in sorted(used_forms)]
    nt, nz = len(tforms), nv + len(tforms)

    # ---- objective (Q7) ----
    Pv = np.zeros((nv, nv))
    qv = np.zeros((nv, 1 + npar))            # q = qv [1; p]
    cost_rows = []                           # (cv, cpar incl. -ref, weight): weight*|cv.v + cpar.[1;p]|
    cost_a = {}                              # alpha index -> weight
    cc_lin = np.zeros(1 + npar)
    CC2 = np.zeros((npar, npar))
    wmax = 0.0
    if simplified:
        xcost = [xbar[k] for k in range(1, N + 1)]
        ucost = [vaff[k] for k in range(N)]
    else:
        xcost = [xbar[k] for k in range(N)]
        ucost = []
    for rows, Qm, ref, wabs, dim in ((xcost, cost.Q, cost.x_ref, cost.w_abs, n), (ucost, cost.R, cost.u_ref, cost.r_abs, m)):
        ref = np.zeros(dim) if ref is None else np.asarray(ref, dtype=np.float64).reshape(dim)
        for aff in rows:
            cv, cp = aff[:, vsl], aff[:, psl].copy()
            cp[:, 0] -= ref
            if Qm is not None:
                Qm_ = np.asarray(Qm, dtype=np.float64)
                wmax = max(wmax, float(np.max(np.abs(Qm_))))
                Pv += 2.0 * cv.T @ Qm_ @ cv
                qv += 2.0 * cv.T @ Qm_ @ cp
                M2 = cp.T @ Qm_ @ cp                 # [1;p]' M2 [1;p]
                cc_lin[0] += M2[0, 0]
                cc_lin[1:] += M2[0, 1:] + M2[1:, 0]
                CC2 += M2[1:, 1:]
            if wabs is not None:
                for j, wj in enumerate(np.asarray(wabs, dtype=np.float64).reshape(dim)):
                    if wj == 0.0:
                        continue
                    assert wj > 0.0, "negative |.| weights are not convex"
                    wmax = max(wmax, float(wj))
                    if np.any(cv[j] != 0.0):
                        cost_rows.append((cv[j].copy(), cp[j].copy(), float(wj), float(ref[j])))
                    elif np.any(cp[j, 1:] != 0.0):
                        f = np.zeros(nw)
                        f[psl] = cp[j]
                        i, wgt = pforms.add(f)
                        cost_a[i] = cost_a.get(i, 0.0) + wj * wgt
                    else:
                        cc_lin[0] += wj * abs(cp[j, 0])

    na = len(pforms.forms)
    ncol = 1 + npar + na

    def shift_of(sense: int, cpar: np.ndarray, aw: dict, cst: float) -> np.ndarray:
        """hi-row: a z <= hi - centre(p) - sum h alpha - cst ;  lo-row: a z >= lo - centre(p) + sum h alpha + cst."""
        r = np.zeros(ncol)
        r[:1 + npar] = -cpar
        r[0] -= sense * cst
        for i, wgt in aw.items():
            r[1 + npar + i] -= sense * wgt
        return r

    # rows: [a (nz), l0, u0, kink0, wabs, shift (ncol)] ; merged when (a, shift) coincide
    rows, index, chk = [], {}, []

    def add_row(a, lo_, hi_, r, kink=0.0, w=0.0):
        key = (tuple(np.round(a, 12) + 0.0), tuple(np.round(r, 12) + 0.0))
        i = index.get(key)
        if i is None:
            index[key] = len(rows)
            rows.append([a, lo_, hi_, kink, w, r])
        else:
            rows[i][1] = max(rows[i][1], lo_)
            rows[i][2] = min(rows[i][2], hi_)
            if w > 0.0:
                assert rows[i][4] == 0.0 or rows[i][3] == kink
                rows[i][3], rows[i][4] = kink, rows[i][4] + w

    for cv, tw, sense, bound, cpar, aw, cst in expanded:
        a = np.zeros(nz)
        a[:nv] = cv
        for i, wgt in tw.items():
            a[nv + remap[i]] = sense * wgt          # hi: +sum w t ; lo: a z - sum w t >= ...
        r = shift_of(sense, cpar, aw, cst)
        if not np.any(a != 0.0):
            # parameter-only row (e.g. xbar0 + e0 in X at k = 0): a feasibility check, 0 within [l, u]
            rr = r.copy()
            rr[0] += bound
            chk.append(-rr if sense > 0 else rr)      # hi: 0 <= bound + r  ;  lo: 0 >= bound + r
            continue
        add_row(a, -INF if sense > 0 else bound, bound if sense > 0 else INF, r)
    # |.| cost rows: handled by the prox of  w|z - kink|  in the ADMM z-update (no slack variable)
    for cv, cp, wj, refj in cost_rows:
        a = np.zeros(nz)
        a[:nv] = cv
        r = np.zeros(ncol)
        r[:1 + npar] = -cp
        r[0] -= refj                                  # cp already has -ref folded in: undo, keep ref as the kink
        add_row(a, -INF, INF, r, kink=refj, w=wj)
    # epigraph rows of the remaining atoms:  f(v,p) - t <= 0  and  f(v,p) + t >= 0
    for i, f in enumerate(tforms):
        for sgn in (-1.0, +1.0):
            a = np.zeros(nz)
            a[:nv] = f[vsl]
            a[nv + i] = sgn
            r = np.zeros(ncol)
            r[:1 + npar] = -f[psl]
            add_row(a, -INF if sgn < 0 else 0.0, 0.0 if sgn < 0 else INF, r)
    nc = len(rows)
    A = np.asarray([r_[0] for r_ in rows]).reshape(nc, nz)
    l0 = np.asarray([r_[1] for r_ in rows], dtype=np.float64)
    u0 = np.asarray([r_[2] for r_ in rows], dtype=np.float64)
    kink0 = np.asarray([r_[3] for r_ in rows], dtype=np.float64)
    wabs_rows = np.asarray([r_[4] for r_ in rows], dtype=np.float64)
    R = np.asarray([r_[5] for r_ in rows]).reshape(nc, ncol)

    P = np.zeros((nz, nz))
    P[:nv, :nv] = Pv
    q0 = np.zeros(nz)
    Qp = np.zeros((nz, npar))
    q0[:nv], Qp[:nv] = qv[:, 0], qv[:, 1:]
    cc = np.zeros(ncol)
    cc[:1 + npar] = cc_lin
    for i, wgt in cost_a.items():
        cc[1 + npar + i] += wgt
    Bt = np.asarray([f[1 + nv:] for f in pforms.forms]).reshape(na, npar)
    gam = np.asarray([f[0] for f in pforms.forms]).reshape(na)
    Rchk = np.asarray(chk).reshape(len(chk), ncol)

    XB = xbar.reshape((N + 1) * n, nw)
    # sparse term table of Ze[1].Z (row-major over (r, j))
    flat = ze1.reshape(n * (1 + g1), nw)
    ptr = np.zeros(flat.shape[0] + 1, dtype=np.int32)
    idx, val = [], []
    for e in range(flat.shape[0]):
        nzc = np.flatnonzero(flat[e])
        idx.extend(nzc.tolist())
        val.extend(flat[e, nzc].tolist())
        ptr[e + 1] = len(idx)
    prog = CompiledProgram(n=n, m=m, N=N, nv=nv, nt=nt, nz=nz, nc=nc, npar=npar, na=na, g1=g1,
                           P=P, q0=q0, Qp=Qp, A=A, l0=l0, u0=u0, kink0=kink0, wabs=wabs_rows, R=R, Bt=Bt, gam=gam, Rchk=Rchk, cc=cc, CC2=CC2,
                           XB=XB, ze1_ptr=ptr, ze1_idx=np.asarray(idx, dtype=np.int32),
                           ze1_val=np.asarray(val, dtype=np.float64),
                           gens_per_step=gens_per_step, wmax=max(wmax, 1.0))
    prog.D, prog.E, prog.c = ruiz_equilibrate(P, A, q0, Qp, wabs_rows)
    return prog


def ruiz_equilibrate(P: np.ndarray, A: np.ndarray, q0: np.ndarray, Qp: np.ndarray, wabs: np.ndarray, iters: int = 15):
    """Modified Ruiz equilibration of [[P, A'], [A, 0]] (OSQP, Stellato et al. 2020, Alg. 2).
    Model-only, so it is done once here; the kernels scale q, l, u per scenario."""
    nz, nc = P.shape[0], A.shape[0]
    D, E = np.ones(nz), np.ones(nc)
    c = 1.0
    Pb, Ab = P.copy(), A.copy()
    for _ in range(iters):
        cn = np.maximum(np.max(np.abs(Pb), axis=0, initial=0.0), np.max(np.abs(Ab), axis=0, initial=0.0))
        rn = np.max(np.abs(Ab), axis=1, initial=0.0)
        d = 1.0 / np.sqrt(np.where(cn > 1e-12, cn, 1.0))
        e = 1.0 / np.sqrt(np.where(rn > 1e-12, rn, 1.0))
        Pb = d[:, None] * Pb * d[None, :]
        Ab = e[:, None] * Ab * d[None, :]
        D *= d
        E *= e
    qn = max(float(np.max(np.abs(D * q0), initial=0.0)), float(np.max(np.abs(D[:, None] * Qp), initial=0.0)),
             float(np.max(wabs / E, initial=0.0)))
    pn = float(np.mean(np.max(np.abs(c * Pb), axis=0))) if nz else 0.0
    g = max(pn, qn)
    c = 1.0 / g if g > 1e-12 else 1.0
    return D, E, c
