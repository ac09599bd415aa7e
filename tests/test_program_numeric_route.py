"""A third, NUMERIC route to the per-step program, independent of the coefficient-tensor machinery that both
oracle/program.py and tzddpc_b200/program.py use: for given numbers (xbar0, e0, v) the statements of
tzddpc/tzddpc.py:163-207 are executed literally on numeric zonotopes written out here (matrix-zonotope product, Minkowski
sum, interval hull -- SURVEY.md App. A.1-A.4), and the result is compared with what the PRODUCT's compiled program says at
the same point:

  * Ze[1].Z (generators as a column multiset),
  * the largest constraint violation (tightened state / input sets of :191-197 and the user's box constraints) -- i.e.
    the feasible set of the canonical program is the feasible set of the literal statements,
  * the objective value.

No solver, no GPU.  (The oracle only supplies the identified model matrices here.)"""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs


# ---- numeric zonotope arithmetic, written out ---------------------------------------------------
def mz_times_z(C, Gm, Z):
    """<C, {G_i}> * <c, G>  =  < C c,  [C G | G_1 c ... G_N c | G_1 G ... G_N G] >."""
    c, G = Z[:, :1], Z[:, 1:]
    cols = [C @ c, C @ G]
    cols += [Gi @ c for Gi in Gm]
    cols += [Gi @ G for Gi in Gm]
    return np.hstack(cols)


def plus(Z1, Z2):
    return np.hstack([Z1[:, :1] + Z2[:, :1], Z1[:, 1:], Z2[:, 1:]])


def radius(Z):
    return np.abs(Z[:, 1:]).sum(axis=1)


def literal(o, cfg, xbar0, e0, v):
    """tzddpc/tzddpc.py:163-207 with numbers.  Returns (Ze list, xbar, worst violation of the tube constraints)."""
    n, m, N = cfg.n, cfg.m, v.shape[0]
    A, B = o.Mdata.center[:, :n], o.Mdata.center[:, n:]                                   # :163
    K = o.theta.K
    MK = (o.MdataK.center, list(o.MdataK.generators))
    MD = (o.Mdelta.center, list(o.Mdelta.generators))
    W = o.zonotopes.W.Z
    xbar = [np.asarray(xbar0, dtype=np.float64)]
    for k in range(N):                                                                    # :166-170
        xbar.append(A @ xbar[k] + B @ v[k])
    Ze = [np.hstack([np.asarray(e0, dtype=np.float64)[:, None], np.zeros((n, 1))])]       # :172
    XU = [np.hstack([np.r_[xbar[k], v[k]][:, None], np.zeros((n + m, 1))]) for k in range(N)]      # :174
    T1 = [mz_times_z(*MK, Ze[0])]                                                         # :175
    Zn = [plus(mz_times_z(*MD, XU[k]), W) for k in range(N)]                              # :176
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    viol = -np.inf
    for k in range(N):                                                                    # :178-209
        T1.append(mz_times_z(*MK, T1[-1]))                                                # :181
        noise = Zn[0]
        for j in range(1, k):                                                             # :183-185
            noise = plus(mz_times_z(*MK, noise), Zn[j])
        Zk = Ze[-1]
        c, r = Zk[:, 0] + xbar[k], radius(Zk)                                             # :191,194-195
        viol = max(viol, (Xi.left_limit - (c - r)).max(), ((c + r) - Xi.right_limit).max())
        KZ = K @ Zk                                                                       # :192,196-197
        cu, ru = KZ[:, 0] + v[k], radius(KZ)
        viol = max(viol, (Ui.left_limit - (cu - ru)).max(), ((cu + ru) - Ui.right_limit).max())
        Ze.append(plus(T1[k], noise))                                                     # :205
    return Ze, np.array(xbar), viol


def literal_simplified(o, cfg, xbar0, e0, v, k0):
    """tzddpc/tzddpc.py:274-327 (build_problem_simplified) with numbers."""
    n, m, N = cfg.n, cfg.m, v.shape[0]
    A, B = o.Mdata.center[:, :n], o.Mdata.center[:, n:]                                   # :274
    K = o.theta.K
    MK = (o.MdataK.center, list(o.MdataK.generators))
    MD = (o.Mdelta.center, list(o.Mdelta.generators))
    W = o.zonotopes.W.Z
    xbar = [np.asarray(xbar0, dtype=np.float64)]
    for k in range(N):                                                                    # :277-281
        xbar.append(A @ xbar[k] + B @ v[k])
    Ze = [np.hstack([np.asarray(e0, dtype=np.float64)[:, None], np.zeros((n, 1))])]       # :283
    XU = [np.hstack([np.r_[xbar[k], v[k]][:, None], np.zeros((n + m, 1))]) for k in range(N)]      # :285
    T1 = [mz_times_z(*MK, Ze[0])]                                                         # :286
    Zn = [plus(mz_times_z(*MD, XU[k]), W) for k in range(N)]                              # :287
    T2 = []
    for k in range(N):                                                                    # :291-302
        T1.append(T1[-1] if k > k0 else mz_times_z(*MK, T1[-1]))
        start = max(0, k - k0)
        noise = Zn[start]
        for j in range(1, min(k, k0)):
            noise = plus(mz_times_z(*MK, noise), Zn[start + j])
        T2.append(noise)
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    viol = -np.inf
    for k in range(N):                                                                    # :305-323
        Zk = Ze[-1]
        c, r = Zk[:, 0] + xbar[k], radius(Zk)
        viol = max(viol, (Xi.left_limit - (c - r)).max(), ((c + r) - Xi.right_limit).max())
        KZ = K @ Zk
        cu, ru = KZ[:, 0] + v[k], radius(KZ)
        viol = max(viol, (Ui.left_limit - (cu - ru)).max(), ((cu + ru) - Ui.right_limit).max())
        Ze.append(plus(T1[k], T2[k]))
    return Ze, np.array(xbar), viol


def stage_cost_simplified(cfg, xbar, v):
    """build_loss(v, xbar[1:]) (tzddpc/tzddpc.py:336): the x terms over xbar_1..xbar_N, the u terms over v."""
    c = cfg.cost
    total = 0.0
    for k in range(v.shape[0]):
        d = xbar[k + 1] - (c["x_ref"] if c.get("x_ref") is not None else 0.0)
        if c.get("Q") is not None:
            total += float(d @ np.asarray(c["Q"]) @ d)
        if c.get("w_abs") is not None:
            total += float(np.abs(d) @ np.asarray(c["w_abs"]))
        du = v[k] - (c["u_ref"] if c.get("u_ref") is not None else 0.0)
        if c.get("R") is not None:
            total += float(du @ np.asarray(c["R"]) @ du)
        if c.get("r_abs") is not None:
            total += float(np.abs(du) @ np.asarray(c["r_abs"]))
    return total


def stage_cost(cfg, xbar, N):
    """The loss callbacks of the examples receive the FREE variable u (quirk Q7: the u terms sit at their floor 0) and the
    rows xbar_0..xbar_{N-1} (tzddpc/tzddpc.py:160,222)."""
    c = cfg.cost
    total = 0.0
    for k in range(N):
        d = xbar[k] - (c["x_ref"] if c.get("x_ref") is not None else 0.0)
        if c.get("Q") is not None:
            total += float(d @ np.asarray(c["Q"]) @ d)
        if c.get("w_abs") is not None:
            total += float(np.abs(d) @ np.asarray(c["w_abs"]))
    return total


def box_violation(cfg, xbar, v):
    b = cfg.box or {}
    viol = -np.inf
    if b.get("x_lo") is not None:
        viol = max(viol, (np.asarray(b["x_lo"]) - xbar).max())
    if b.get("x_hi") is not None:
        viol = max(viol, (xbar - np.asarray(b["x_hi"])).max())
    if b.get("v_lo") is not None:
        viol = max(viol, (np.asarray(b["v_lo"]) - v).max())
    if b.get("v_hi") is not None:
        viol = max(viol, (v - np.asarray(b["v_hi"])).max())
    return viol


def program_at(prog, p, z):
    """Worst row violation and objective of the CompiledProgram (include/tzddpc.h, TzProgramDesc) at parameters p, variables z."""
    alpha = np.abs(prog.Bt @ p + prog.gam) if prog.na else np.zeros(0)
    w = np.r_[1.0, p, alpha]
    r = prog.R @ w
    Az = prog.A @ z
    viol = max((prog.l0 + r - Az).max(), (Az - prog.u0 - r).max())
    if prog.Rchk.shape[0]:
        viol = max(viol, (prog.Rchk @ w).max())
    obj = 0.5 * z @ prog.P @ z + (prog.q0 + prog.Qp @ p) @ z + float(prog.wabs @ np.abs(Az - prog.kink0 - r)) + prog.cc @ w + p @ prog.CC2 @ p
    return viol, float(obj)


@pytest.mark.parametrize("name,horizon", [("double_integrator", None), ("pulley", None), ("fivedim", None),
                                          ("double_integrator", 1), ("double_integrator", 3)])
def test_compiled_program_equals_the_literal_numeric_statements(name, horizon):
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, _ = common.make_oracle(cfg, u, x, horizon=horizon)
    prog = common.make_compiled(cfg, o, horizon=horizon)
    assert prog.nz == prog.nv, "no epigraph variable (z = v) in these programs"
    n, m, N = cfg.n, cfg.m, horizon or cfg.horizon
    rng = np.random.default_rng(17)
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    feas = infeas = 0
    for trial in range(60):
        spread = 0.45 if trial % 3 else 1.2                      # inside the state set / well outside it
        mid, half = 0.5 * (Xi.left_limit + Xi.right_limit), 0.5 * (Xi.right_limit - Xi.left_limit)
        xbar0 = mid + spread * half * rng.uniform(-1, 1, n)
        e0 = 0.02 * rng.uniform(-1, 1, n)
        v = (0.5 * (Ui.left_limit + Ui.right_limit) + spread * 0.5 * (Ui.right_limit - Ui.left_limit) * rng.uniform(-1, 1, (N, m)))
        if trial % 4 == 0:                                       # a point that is feasible for sure: the optimum near X0
            xbar0 = np.asarray(cfg.X0[0], dtype=np.float64) + 0.02 * rng.uniform(-1, 1, n)
            e0 = 0.005 * rng.uniform(-1, 1, n)
            r0 = o.solve_status(xbar0, e0)
            if r0.status == 0:
                v = np.asarray(r0.v, dtype=np.float64).reshape(N, m)
        Ze, xbar, viol_tube = literal(o, cfg, xbar0, e0, v)
        viol_lit = max(viol_tube, box_violation(cfg, xbar, v))
        p = np.r_[xbar0, e0]
        viol_prog, obj_prog = program_at(prog, p, v.ravel())
        scale = max(1.0, np.abs(xbar).max(), np.abs(v).max())
        assert abs(viol_prog - viol_lit) <= 1e-9 * scale, (name, trial, viol_prog, viol_lit)
        obj_lit = stage_cost(cfg, xbar, N)
        assert abs(obj_prog - obj_lit) <= 1e-9 * max(1.0, abs(obj_lit)), (name, trial, obj_prog, obj_lit)
        # Ze[1].Z as the program's term table evaluates it, against the literal set: same centre, same generator multiset
        om = np.r_[1.0, v.ravel(), p]
        Zp = np.zeros((n, 1 + prog.g1))
        for e in range(n * (1 + prog.g1)):
            t0, t1 = prog.ze1_ptr[e], prog.ze1_ptr[e + 1]
            Zp.flat[e] = float(prog.ze1_val[t0:t1] @ om[prog.ze1_idx[t0:t1]])
        Zl = Ze[1]
        np.testing.assert_allclose(Zp[:, 0], Zl[:, 0], rtol=1e-12, atol=1e-13)
        keep_p = Zp[:, 1:][:, np.any(Zp[:, 1:] != 0, axis=0)]
        keep_l = Zl[:, 1:][:, np.any(Zl[:, 1:] != 0, axis=0)]
        assert keep_p.shape == keep_l.shape, (name, trial, keep_p.shape, keep_l.shape)
        np.testing.assert_allclose(common.sort_columns(keep_p), common.sort_columns(keep_l), rtol=1e-12, atol=1e-13)
        feas += viol_lit <= 1e-7 * scale                         # (an optimum sits ON its active constraints)
        infeas += viol_lit > 1e-7 * scale
    assert feas >= 5 and infeas >= 5, (feas, infeas)


@pytest.mark.parametrize("horizon,k0", [(2, 1), (3, 1), (3, 2)])
def test_simplified_program_equals_the_literal_numeric_statements(horizon, k0):
    """build_problem_simplified (tzddpc/tzddpc.py:243-355; the `-m stzddpc` rows of the complexity sweep) on the sweep system,
    for the (horizon, k0) whose canonical program needs no epigraph variable."""
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, _ = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
    prog = common.make_compiled(cfg, o, horizon=horizon, k0=k0)
    assert prog.nz == prog.nv
    n, m, N = cfg.n, cfg.m, horizon
    rng = np.random.default_rng(23)
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    feas = infeas = 0
    for trial in range(60):
        spread = 0.4 if trial % 3 else 2.5                       # inside the state set / well outside it
        mid, half = 0.5 * (Xi.left_limit + Xi.right_limit), 0.5 * (Xi.right_limit - Xi.left_limit)
        xbar0 = mid + spread * half * rng.uniform(-1, 1, n)
        e0 = 0.01 * rng.uniform(-1, 1, n)
        v = 0.5 * (Ui.left_limit + Ui.right_limit) + spread * 0.3 * (Ui.right_limit - Ui.left_limit) * rng.uniform(-1, 1, (N, m))
        if trial % 4 == 0:
            r0 = o.solve_status(xbar0, e0)
            if r0.status == 0:
                v = np.asarray(r0.v, dtype=np.float64).reshape(N, m)
        Ze, xbar, viol_lit = literal_simplified(o, cfg, xbar0, e0, v, k0)
        p = np.r_[xbar0, e0]
        viol_prog, obj_prog = program_at(prog, p, v.ravel())
        scale = max(1.0, np.abs(xbar).max(), np.abs(v).max())
        assert abs(viol_prog - viol_lit) <= 1e-9 * scale, (trial, viol_prog, viol_lit)
        obj_lit = stage_cost_simplified(cfg, xbar, v)
        assert abs(obj_prog - obj_lit) <= 1e-9 * max(1.0, abs(obj_lit)), (trial, obj_prog, obj_lit)
        feas += viol_lit <= 1e-7 * scale
        infeas += viol_lit > 1e-7 * scale
    assert feas >= 5 and infeas >= 5, (feas, infeas)


@pytest.mark.parametrize("horizon,k0", [(4, None), (5, None), (4, 1), (6, 1)])
def test_programs_with_epigraph_variables_have_the_literal_feasible_set(horizon, k0):
    """Longer horizons put several |.| atoms into one tightened row and the canonical program carries epigraph variables
    s (z = [v; s]).  For fixed (p, v) the program is feasible iff some s makes every row hold -- an LP in s, solved here
    with HiGHS -- and that must be exactly when the literal numeric statements hold."""
    from scipy.optimize import linprog
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, _ = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
    prog = common.make_compiled(cfg, o, horizon=horizon, k0=k0)
    ns = prog.nz - prog.nv
    assert ns > 0
    n, m, N = cfg.n, cfg.m, horizon
    rng = np.random.default_rng(31)
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    feas = infeas = 0
    for trial in range(40):
        spread = 0.35 if trial % 2 else 1.6
        mid, half = 0.5 * (Xi.left_limit + Xi.right_limit), 0.5 * (Xi.right_limit - Xi.left_limit)
        xbar0 = mid + spread * half * rng.uniform(-1, 1, n)
        e0 = 0.01 * rng.uniform(-1, 1, n)
        v = 0.5 * (Ui.left_limit + Ui.right_limit) + spread * 0.25 * (Ui.right_limit - Ui.left_limit) * rng.uniform(-1, 1, (N, m))
        if trial % 4 == 1:
            r0 = o.solve_status(xbar0, e0)
            if r0.status == 0:
                v = np.asarray(r0.v, dtype=np.float64).reshape(N, m)
        if k0 is None:
            _, xbar, viol_lit = literal(o, cfg, xbar0, e0, v)
        else:
            _, xbar, viol_lit = literal_simplified(o, cfg, xbar0, e0, v, k0)
        scale = max(1.0, np.abs(xbar).max(), np.abs(v).max())
        if abs(viol_lit) < 1e-6 * scale:
            continue                                            # on the boundary: the verdict is a matter of rounding
        # min_t  s.t.  l - A z <= t,  A z - u <= t  over (s, t) with v fixed
        p = np.r_[xbar0, e0]
        alpha = np.abs(prog.Bt @ p + prog.gam) if prog.na else np.zeros(0)
        w = np.r_[1.0, p, alpha]
        r = prog.R @ w
        Av, As = prog.A[:, :prog.nv], prog.A[:, prog.nv:]
        base = Av @ v.ravel()
        lo, hi = prog.l0 + r, prog.u0 + r
        rows, rhs = [], []
        for i in range(prog.nc):
            if np.isfinite(hi[i]):
                rows.append(np.r_[As[i], -1.0]); rhs.append(hi[i] - base[i])
            if np.isfinite(lo[i]):
                rows.append(np.r_[-As[i], -1.0]); rhs.append(base[i] - lo[i])
        res = linprog(np.r_[np.zeros(ns), 1.0], A_ub=np.array(rows), b_ub=np.array(rhs), bounds=[(None, None)] * (ns + 1), method="highs")
        assert res.status == 0, res.message
        t_min = res.x[-1]
        if prog.Rchk.shape[0]:
            t_min = max(t_min, float((prog.Rchk @ w).max()))
        assert (t_min <= 1e-9 * scale) == (viol_lit <= 0), (horizon, k0, trial, t_min, viol_lit)
        feas += viol_lit <= 0
        infeas += viol_lit > 0
    assert feas >= 3 and infeas >= 3, (feas, infeas)
