"""Glue used by the tests: builds the ORACLE controller and (on a GPU) the product controller from the
same example config, data set and gain.  Only tests may import `oracle`."""
from __future__ import annotations

import numpy as np

import oracle
from oracle.program import generate_trajectories  # noqa: F401
from tzddpc_b200 import configs

COST_RTOL = 1e-6      # north_star: nominal inputs and costs to 1e-6 (the QP solve tolerance)
GEN_RTOL = 1e-9       # north_star: generator matrices and interval bounds to 1e-9 relative


def dataset(cfg, seed=None):
    rng = np.random.default_rng(cfg.seed if seed is None else seed)
    u, x = configs.generate_dataset(cfg, rng)
    return u, x


def oracle_zonotopes(cfg):
    Z = oracle.Zonotope
    return oracle.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))


def make_oracle(cfg, u, x, K=None, horizon=None, k0=None):
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    z = oracle_zonotopes(cfg)
    if K is None:
        o.build_zonotopes(z)
        C = o.Mdata.center
        K = configs.lqr_gain(C[:, :cfg.n], C[:, cfg.n:])
    o.build_zonotopes_theta(z, K)
    box = oracle.BoxConstraint(**cfg.box) if cfg.box else None
    o.build_problem(horizon or cfg.horizon, oracle.StageCost(**cfg.cost), box, k0=k0)
    return o, K


def make_product(cfg, u, x, K, horizon=None, k0=None, verbose=False):
    import tzddpc_b200 as tz
    t = tz.TZDDPC(tz.Data(u, x))
    t.verbose = verbose
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    t.build_zonotopes_theta(zon, K=K)
    box = tz.BoxConstraint(**cfg.box) if cfg.box else tz.BoxConstraint()
    if k0 is None:
        t.build_problem(horizon or cfg.horizon, tz.StageCost(**cfg.cost), box)
    else:
        t.build_problem_simplified(k0, horizon or cfg.horizon, tz.StageCost(**cfg.cost), box)
    return t


def noise_for(cfg, steps, S, rng):
    """Closed-loop noise realisations (steps, S, n): W.sample() (examples/2.pulley_sim.py:92) or a random vertex
    of W (examples/1.double_integrator_sim.py:85)."""
    cW, GW = cfg.W
    if cfg.noise == "sample":
        beta = rng.uniform(-1, 1, size=(steps, S, GW.shape[1]))
    else:
        beta = rng.choice([-1.0, 1.0], size=(steps, S, GW.shape[1]))
    return cW[None, None] + beta @ GW.T


def cost_close(a, b, wmax, rtol=COST_RTOL):
    scale = np.maximum(1.0, np.maximum(np.abs(b), wmax))
    return np.abs(a - b) <= rtol * scale


def sort_columns(G):
    """Generator matrices are compared as column multisets."""
    idx = np.lexsort(G[::-1])
    return G[:, idx]


def make_compiled(cfg, o, horizon=None, k0=None):
    """The PRODUCT's host canonicaliser (tzddpc_b200/program.py, no GPU needed) fed with the oracle's model."""
    from tzddpc_b200 import program as P
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    model = P.TubeModel(AB=o.Mdata.center, Acl=o.MdataK.center, GK=o.MdataK.generators, GD=o.Mdelta.generators,
                        K=o.theta.K, WZ=o.zonotopes.W.Z, X_lo=Xi.left_limit, X_hi=Xi.right_limit,
                        U_lo=Ui.left_limit, U_hi=Ui.right_limit)
    box = P.BoxConstraint(**cfg.box) if cfg.box else P.BoxConstraint()
    return P.compile_program(model, horizon or cfg.horizon, P.StageCost(**cfg.cost), box, k0=k0)


def solve_compiled(prog, xbar0, e0):
    """Evaluates a CompiledProgram at p = [xbar0; e0] on the CPU with the oracle's interior-point solver:
    the reading of include/tzddpc.h's TzProgramDesc that the CUDA kernels implement.
    Returns dict(status, cost, v, xbar (N+1, n), ze1 (n, 1+g1))."""
    p = np.r_[xbar0, e0]
    alpha = np.abs(prog.Bt @ p + prog.gam) if prog.na else np.zeros(0)
    w = np.r_[1.0, p, alpha]
    if prog.Rchk.shape[0]:
        t = prog.Rchk * w[None]
        if np.any(t.sum(axis=1) > 1e-9 * np.maximum(1.0, np.abs(t).max(axis=1))):
            return dict(status=2, cost=np.inf)
    r = prog.R @ w
    l, u = prog.l0 + r, prog.u0 + r
    kink = prog.kink0 + r
    kr = np.flatnonzero(prog.wabs > 0)
    nz, ns = prog.nz, len(kr)
    G, h = [], []
    for i in range(prog.nc):
        a = np.r_[prog.A[i], np.zeros(ns)]
        if np.isfinite(u[i]):
            G.append(a); h.append(u[i])
        if np.isfinite(l[i]):
            G.append(-a); h.append(-l[i])
    for j, i in enumerate(kr):                 # s_j >= |a_i z - kink_i|
        a = np.r_[prog.A[i], np.zeros(ns)]
        sj = np.zeros(nz + ns); sj[nz + j] = 1.0
        G.append(a - sj); h.append(kink[i])
        G.append(-a - sj); h.append(-kink[i])
    G, h = np.asarray(G).reshape(-1, nz + ns), np.asarray(h)
    P = np.zeros((nz + ns, nz + ns)); P[:nz, :nz] = prog.P
    q = np.r_[prog.q0 + prog.Qp @ p, prog.wabs[kr]]
    from scipy.optimize import linprog
    fr = linprog(np.zeros(nz + ns), A_ub=G, b_ub=h + 1e-9 * np.maximum(1.0, np.abs(h)), bounds=[(None, None)] * (nz + ns),
                 method="highs")
    if fr.status == 2:
        return dict(status=2, cost=np.inf)
    y, _, info = oracle.solve_qp_ipm(P, q, G, h)
    c0 = prog.cc @ w + p @ prog.CC2 @ p
    cost = 0.5 * y @ P @ y + q @ y + c0
    om = np.r_[1.0, y[:prog.nv], p]
    xbar = (prog.XB @ om).reshape(prog.N + 1, prog.n)
    ze1 = np.zeros(prog.n * (1 + prog.g1))
    for e in range(ze1.shape[0]):
        t0, t1 = prog.ze1_ptr[e], prog.ze1_ptr[e + 1]
        ze1[e] = prog.ze1_val[t0:t1] @ om[prog.ze1_idx[t0:t1]]
    return dict(status=0 if info["status"] == "optimal" else 1, cost=float(cost), v=y[:prog.nv].reshape(prog.N, prog.m),
                xbar=xbar, ze1=ze1.reshape(prog.n, 1 + prog.g1))
