"""The oracle's solver and `solve` against independent solvers (scipy HiGHS for the LP examples, SLSQP for the QP
example) and against its own committed outputs (tests/golden/oracle_*.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
from scipy.optimize import linprog, minimize

import oracle
from tests import common
from tzddpc_b200 import configs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ipm_random_qp_against_slsqp():
    rng = np.random.default_rng(0)
    for trial in range(5):
        n, m = 4, 9
        M = rng.normal(size=(n, n))
        P = M @ M.T + 0.1 * np.eye(n)
        q = rng.normal(size=n)
        G = rng.normal(size=(m, n))
        h = rng.uniform(0.1, 1.0, size=m)
        x, z, info = oracle.solve_qp_ipm(P, q, G, h)
        assert info["status"] == "optimal"
        res = minimize(lambda y: 0.5 * y @ P @ y + q @ y, np.zeros(n), jac=lambda y: P @ y + q, method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda y: h - G @ y, "jac": lambda y: -G}],
                       options={"ftol": 1e-14, "maxiter": 500})
        np.testing.assert_allclose(x, res.x, atol=2e-6)
        # KKT
        assert np.all(G @ x <= h + 1e-9) and np.all(z >= -1e-12)
        assert np.abs(P @ x + q + G.T @ z).max() < 1e-8
        assert np.abs(z * (h - G @ x)).max() < 1e-8


@pytest.mark.parametrize("name", ["pulley", "fivedim"])
def test_lp_examples_against_highs(name):
    cfg = configs.CONFIGS[name]()
    fx = np.load(os.path.join(GOLD, f"oracle_{name}.npz"))
    o, _ = common.make_oracle(cfg, fx["u_data"], fx["x_data"], K=fx["K"])
    done = 0
    for i in range(0, 24):
        P, q, c0, G, h, ok = o._assemble(fx["xbar0"][i], fx["e0"][i])
        assert not np.any(P)                       # ex.2 / ex.3 are LPs
        res = linprog(q, A_ub=G, b_ub=h, bounds=[(None, None)] * len(q), method="highs")
        if not ok or res.status == 2:
            assert fx["status"][i] == 2
            continue
        assert fx["status"][i] == 0 and res.status == 0
        done += 1
        wmax = float(np.max(cfg.cost["w_abs"]))
        assert common.cost_close(res.fun + c0, fx["cost"][i], wmax, rtol=1e-7)
    assert done >= 5


def test_qp_example_against_slsqp():
    cfg = configs.double_integrator()
    fx = np.load(os.path.join(GOLD, "oracle_double_integrator.npz"))
    o, _ = common.make_oracle(cfg, fx["u_data"], fx["x_data"], K=fx["K"])
    done = 0
    for i in range(0, 16):
        if fx["status"][i] != 0:
            continue
        P, q, c0, G, h, ok = o._assemble(fx["xbar0"][i], fx["e0"][i])
        y0 = np.r_[fx["v"][i].ravel(), np.zeros(len(q) - fx["v"][i].size)]
        y0[fx["v"][i].size:] = np.maximum(0.0, np.max(G[:, :fx["v"][i].size] @ fx["v"][i].ravel() - h)) + 1.0
        res = minimize(lambda y: 0.5 * y @ P @ y + q @ y, y0, jac=lambda y: P @ y + q, method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda y: h - G @ y, "jac": lambda y: -G}],
                       options={"ftol": 1e-13, "maxiter": 1000})
        assert np.all(G @ res.x <= h + 1e-7)           # (SLSQP may stop on its line search at the optimum)
        assert abs(res.fun + c0 - fx["cost"][i]) <= 1e-6 * max(1.0, abs(fx["cost"][i]))
        done += 1
    assert done >= 5


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_oracle_reproduces_committed_golden(name):
    cfg = configs.CONFIGS[name]()
    fx = np.load(os.path.join(GOLD, f"oracle_{name}.npz"))
    u, x = common.dataset(cfg)
    np.testing.assert_array_equal(u, fx["u_data"])       # the seeded data set itself (configs.generate_dataset)
    np.testing.assert_array_equal(x, fx["x_data"])
    o, K = common.make_oracle(cfg, u, x)
    np.testing.assert_allclose(K, fx["K"], rtol=1e-10)
    np.testing.assert_allclose(o.Mdata.center, fx["AB"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(o.Mdelta.generators, fx["GD"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(o.MdataK.generators, fx["GK"], rtol=1e-9, atol=1e-14)
    for i in range(0, fx["xbar0"].shape[0], 3):
        r = o.solve_status(fx["xbar0"][i], fx["e0"][i])
        assert r.status == fx["status"][i]
        if r.status == 2:
            assert np.isinf(fx["cost"][i])
            continue
        wmax = max(1.0, float(np.max(cfg.cost.get("w_abs", np.ones(1)))))
        assert common.cost_close(r.cost, fx["cost"][i], wmax, rtol=1e-8)
        np.testing.assert_allclose(r.v[0], fx["v"][i, 0], rtol=1e-7, atol=1e-8)
        np.testing.assert_allclose(r.Ze1, fx["ze1"][i], rtol=1e-7, atol=1e-8)


def test_known_answer_pulley_unconstrained_step():
    """with the tightened constraints inactive the N = 2 pulley step gives xbar_1[0] = 1 and
    v0 = (1 - Ahat[0,:] xbar0) / Bhat[0]; the returned cost includes the constant |xbar0[0] - 1| (quirk Q7)."""
    cfg = configs.pulley()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    AB = o.Mdata.center
    xb = np.array([0.9, 1.1, 1.0, 0.95])
    r = o.solve_status(xb, np.zeros(4))
    assert r.status == 0
    v0 = (1.0 - AB[0, :4] @ xb) / AB[0, 4]
    np.testing.assert_allclose(r.v[0, 0], v0, rtol=1e-8)
    np.testing.assert_allclose(r.xbar[1, 0], 1.0, atol=1e-9)
    np.testing.assert_allclose(r.cost, abs(xb[0] - 1.0), atol=1e-8)


def test_solve_raises_like_the_reference_when_infeasible():
    cfg = configs.pulley()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    with pytest.raises(Exception, match="Problem is unbounded"):          # tzddpc/tzddpc.py:374-375
        o.solve(np.array([10.0, 10, 10, 10]), np.zeros(4))
