"""Host logic without a GPU: the product's one-off canonicalisation (tzddpc_b200/program.py, the counterpart of
cvxpy's canonicalisation of tzddpc/tzddpc.py:132-241) must describe the same convex program as the oracle's literal
restatement -- same optimum, same nominal trajectory, same Ze[1].Z, same feasibility verdicts."""
import os

import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", params=["double_integrator", "pulley", "fivedim"])
def setup(request):
    cfg = configs.CONFIGS[request.param]()
    fx = np.load(os.path.join(GOLD, f"oracle_{request.param}.npz"))
    o, _ = common.make_oracle(cfg, fx["u_data"], fx["x_data"], K=fx["K"])
    return cfg, o, common.make_compiled(cfg, o), fx


def test_compiled_program_shapes(setup):
    cfg, o, prog, fx = setup
    assert prog.nv == cfg.horizon * cfg.m and prog.npar == 2 * cfg.n
    assert prog.g1 == o.num_generators_log[0] == fx["ze1"].shape[2] - 1
    assert list(prog.gens_per_step) == list(o.num_generators_log)          # what tzddpc/tzddpc.py:206 prints
    assert prog.A.shape == (prog.nc, prog.nz) and prog.R.shape == (prog.nc, 1 + prog.npar + prog.na)
    assert np.all(prog.D > 0) and np.all(prog.E > 0) and prog.c > 0
    # every generator entry of Ze[1].Z has at most one term (boxed M_K / M_Delta): what the CUDA output phase relies on
    ld = 1 + prog.g1
    cnt = np.diff(prog.ze1_ptr).reshape(cfg.n, ld)
    assert cnt[:, 1:].max() <= 1


def test_compiled_program_matches_oracle_golden(setup):
    cfg, o, prog, fx = setup
    n_ok = 0
    for i in range(0, fx["xbar0"].shape[0], 2):
        r = common.solve_compiled(prog, fx["xbar0"][i], fx["e0"][i])
        assert (r["status"] == 2) == (fx["status"][i] == 2), f"point {i}: feasibility verdict differs"
        if r["status"] == 2:
            continue
        n_ok += 1
        assert common.cost_close(r["cost"], fx["cost"][i], prog.wmax, rtol=1e-7), (i, r["cost"], fx["cost"][i])
        np.testing.assert_allclose(r["v"][0], fx["v"][i, 0], rtol=1e-6, atol=1e-7)        # Q12: v[0], xbar[1] are unique
        np.testing.assert_allclose(r["xbar"][:2], fx["xbar"][i, :2], rtol=1e-6, atol=1e-7)
        # the tube is a function of v[0] alone: exact comparison at the same v
        Zo = o.evaluate_tube(fx["xbar0"][i], fx["e0"][i], r["v"].ravel(), 1)
        np.testing.assert_allclose(r["ze1"], Zo, rtol=1e-12, atol=1e-14)
    assert n_ok >= 5


@pytest.mark.parametrize("horizon,k0", [(1, None), (3, None), (3, 1), (4, 2)])
def test_longer_horizons_and_simplified_variant(horizon, k0):
    """build_problem for N != 2 and build_problem_simplified(k0, N) (tzddpc/tzddpc.py:243-355) on the double integrator."""
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, _ = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
    prog = common.make_compiled(cfg, o, horizon=horizon, k0=k0)
    assert list(prog.gens_per_step) == list(o.num_generators_log)
    rng = np.random.default_rng(horizon * 10 + (k0 or 0))
    Xi = o.zonotopes.X.interval
    done = 0
    for _ in range(6):
        xb = Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.3, 0.7, cfg.n)
        e = rng.uniform(-0.01, 0.01, cfg.n)
        ro = o.solve_status(xb, e)
        rp = common.solve_compiled(prog, xb, e)
        assert (ro.status == 2) == (rp["status"] == 2)
        if ro.status == 2:
            continue
        done += 1
        assert common.cost_close(rp["cost"], ro.cost, prog.wmax, rtol=1e-7)
        np.testing.assert_allclose(rp["xbar"][1], ro.xbar[1], rtol=1e-6, atol=1e-7)
        if horizon >= 2:
            np.testing.assert_allclose(rp["ze1"], o.evaluate_tube(xb, e, rp["v"].ravel(), 1), rtol=1e-12, atol=1e-14)
    assert done >= 2


def test_callbacks_reduce_to_the_structured_cost():
    """The reference's loss/constraint callbacks (examples/*.py) written against tzddpc_b200.cvx give the same
    StageCost / BoxConstraint as the structured presets in configs.py."""
    from tzddpc_b200 import cvx as cp

    def loss_ex1(u, y):                      # examples/1.double_integrator_sim.py:22-28
        horizon, dim_u, dim_x = u.shape[0], u.shape[1], y.shape[1]
        cost = 0
        for i in range(horizon):
            cost += cp.norm(y[i, :], p=2) ** 2 + 1e-2 * cp.norm(u[i], p=1)
        return cost

    def loss_ex2(u, y):                      # examples/2.pulley_sim.py:17-22
        cost = 0
        for i in range(u.shape[0]):
            cost += cp.norm(y[i, 0] - 1, p=2)
        return cost

    def loss_ex3(u, y):                      # examples/3.5dimsystem_sim.py:14-20
        cost = 0
        for i in range(u.shape[0]):
            cost += 1e9 * cp.norm(y[i, 1] - 2, p=2) + 1e-1 * cp.norm(u[i], p=2)
        return cost

    def cons_ex3(u, y):                      # examples/3.5dimsystem_sim.py:23-26
        return [y[:, 1] <= 10, y[:, 1] >= 2]

    c1 = cp.extract_stage_cost(loss_ex1, 2, 2, 1, False)
    np.testing.assert_allclose(c1.Q, np.eye(2))
    np.testing.assert_allclose(c1.r_abs, [0.01])
    c2 = cp.extract_stage_cost(loss_ex2, 2, 4, 1, False)
    np.testing.assert_allclose(c2.w_abs, [1, 0, 0, 0]); np.testing.assert_allclose(c2.x_ref[0], 1.0)
    c3 = cp.extract_stage_cost(loss_ex3, 2, 5, 1, False)
    np.testing.assert_allclose(c3.w_abs, [0, 1e9, 0, 0, 0]); np.testing.assert_allclose(c3.x_ref[1], 2.0)
    np.testing.assert_allclose(c3.r_abs, [0.1])          # the 2-norm of a scalar input is its absolute value
    b3 = cp.extract_box_constraints(cons_ex3, 2, 5, 1, False)
    assert b3.x_hi[1] == 10 and b3.x_lo[1] == 2 and np.isinf(b3.x_hi[0]) and np.isinf(b3.x_lo[0])
    b0 = cp.extract_box_constraints(lambda u, y: [], 2, 5, 1, False)
    assert b0.x_lo is None or not np.any(np.isfinite(b0.x_lo))


def test_batched_canonicalisation_equals_the_single_one_bit_for_bit():
    """compile_program_batch (the data-set axis: D models of one structure canonicalised at once) against compile_program on
    each model: every field of every slice identical, whatever the batch the model sits in."""
    from tzddpc_b200 import program as P
    for name, D in (("fivedim", 5), ("double_integrator", 3)):
        cfg = configs.CONFIGS[name]()
        models = []
        for d in range(D):
            u, x = common.dataset(cfg, seed=cfg.seed + 101 * d)
            o, _ = common.make_oracle(cfg, u, x)
            Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
            models.append(P.TubeModel(AB=o.Mdata.center, Acl=o.MdataK.center, GK=o.MdataK.generators, GD=o.Mdelta.generators,
                                      K=o.theta.K, WZ=o.zonotopes.W.Z, X_lo=Xi.left_limit, X_hi=Xi.right_limit,
                                      U_lo=Ui.left_limit, U_hi=Ui.right_limit))
        cost, box = P.StageCost(**cfg.cost), (P.BoxConstraint(**cfg.box) if cfg.box else P.BoxConstraint())
        B = P.compile_program_batch(P.TubeModelBatch.of(models), cfg.horizon, cost, box)
        B2 = P.compile_program_batch(P.TubeModelBatch.of(models[::-1] + models), cfg.horizon, cost, box)     # another batch composition
        assert B.num == D
        for d in range(D):
            s, b, b2 = P.compile_program(models[d], cfg.horizon, cost, box), B.program(d), B2.program(D + d)
            for f in ("P", "q0", "Qp", "A", "l0", "u0", "kink0", "wabs", "R", "Bt", "gam", "Rchk", "cc", "CC2", "XB", "ze1_ptr", "ze1_idx",
                      "ze1_val", "D", "E"):
                assert np.array_equal(getattr(s, f), getattr(b, f)), (name, d, f)
                assert np.array_equal(getattr(s, f), getattr(b2, f)), (name, d, f, "batch composition")
            assert s.c == b.c and (s.nc, s.nz, s.na, s.g1) == (b.nc, b.nz, b.na, b.g1)


def test_boxed_model_batch_equals_the_generator_form():
    """TubeModelBatch.boxed (straight from tz_identify's centre and boxes) builds the model the MatrixZonotope route builds."""
    from tzddpc_b200 import program as P
    cfg = configs.fivedim()
    u, x = common.dataset(cfg)
    o, _ = common.make_oracle(cfg, u, x)
    n, m = cfg.n, cfg.m
    dAB = np.abs(o.Mdelta.generators).sum(axis=0)
    dK = np.abs(o.MdataK.generators).sum(axis=0)
    Xi, Ui = o.zonotopes.X.interval, o.zonotopes.U.interval
    mb = P.TubeModelBatch.boxed(o.Mdata.center[None], dAB[None], dK[None], o.theta.K[None], o.zonotopes.W.Z, Xi.left_limit,
                                Xi.right_limit, Ui.left_limit, Ui.right_limit)
    np.testing.assert_allclose(mb.Acl[0], o.MdataK.center, rtol=1e-13, atol=1e-15)
    # the generators in Girard's diag order: single-entry matrices d[r, c] E_rc
    np.testing.assert_array_equal(mb.GD[0], o.Mdelta.generators)
    np.testing.assert_array_equal(mb.GK[0], o.MdataK.generators)
