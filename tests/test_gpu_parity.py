"""GPU parity: the CUDA path (through the C ABI / torch ops) against the oracle on identical inputs."""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["double_integrator", "pulley", "fivedim"])
def pair(request, cuda_lib):
    cfg = configs.CONFIGS[request.param]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    return cfg, o, t


def test_identify_matches_oracle(pair):
    cfg, o, t = pair
    np.testing.assert_allclose(t.Mdata.center, o.Mdata.center, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(t.MdataK.center, o.MdataK.center, rtol=1e-9, atol=1e-12)
    for a, b in ((t.Mdata, o.Mdata), (t.MdataK, o.MdataK), (t.Mdelta, o.Mdelta)):
        assert a.num_generators == b.num_generators
        np.testing.assert_allclose(a.generators, b.generators, rtol=common.GEN_RTOL, atol=1e-13)


def _points(cfg, o, rng, count):
    Xi = o.zonotopes.X.interval
    pts = []
    while len(pts) < count:
        xb = Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.05, 0.95, cfg.n)
        e = rng.uniform(-0.4, 0.4, cfg.n)
        pts.append((xb, e))
    return pts


def test_solve_matches_oracle(pair):
    cfg, o, t = pair
    rng = np.random.default_rng(7)
    pts = _points(cfg, o, rng, 64)
    xb = np.array([p[0] for p in pts])
    ee = np.array([p[1] for p in pts])
    cost, v, xbar, tube, status = t.solve(xb, ee)
    Z = tube.Z.value
    wmax = t._program.compiled.wmax
    n_ok = 0
    for i, (a, b) in enumerate(pts):
        r = o.solve_status(a, b)
        if r.status == 2:
            assert status[i] == 2, f"point {i}: oracle infeasible, GPU status {status[i]}"
            assert np.isinf(cost[i])
            continue
        assert status[i] == 0, f"point {i}: GPU status {status[i]}"
        n_ok += 1
        assert common.cost_close(cost[i], r.cost, wmax), (i, cost[i], r.cost)
        np.testing.assert_allclose(v[i, 0], r.v[0], rtol=1e-6, atol=1e-6)          # Q12: only v[0], xbar[1] are unique
        np.testing.assert_allclose(xbar[i, :2], r.xbar[:2], rtol=1e-6, atol=1e-6)
        # Ze[1].Z depends on v[0] only: compare at 1e-9 after evaluating the oracle tube at the GPU's v
        Zo = o.evaluate_tube(a, b, np.r_[v[i].ravel()], 1)
        np.testing.assert_allclose(Z[i], Zo, rtol=common.GEN_RTOL, atol=1e-12)
        np.testing.assert_allclose(Z[i], r.Ze1, rtol=1e-6, atol=1e-6)
    assert n_ok >= 8


def test_solve_batch1_reference_tuple(pair):
    cfg, o, t = pair
    x0 = np.asarray(cfg.X0[0], dtype=np.float64)
    r = o.solve_status(x0, np.zeros(cfg.n))
    if r.status == 2:
        with pytest.raises(Exception, match="unbounded"):
            t.solve(x0, np.zeros(cfg.n))
        return
    result, v, xbar, Ze1 = t.solve(x0, np.zeros(cfg.n))
    assert isinstance(result, float) and v.shape == (cfg.horizon, cfg.m) and xbar.shape == (cfg.horizon + 1, cfg.n)
    assert Ze1.Z.value.shape == (cfg.n, 1 + o.num_generators_log[0])
    assert common.cost_close(result, r.cost, t._program.compiled.wmax)
    np.testing.assert_allclose(v[0], r.v[0], rtol=1e-6, atol=1e-6)


def test_closed_loop_matches_oracle(pair):
    cfg, o, t = pair
    rng = np.random.default_rng(11)
    steps, S = min(cfg.steps, 25), 4
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    out = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True)
    for s in range(S):
        r = o.closed_loop(cfg.A, cfg.B, x0[s], noise[:, s], keep_tubes=True)
        assert np.array_equal(out["status"][:, s] == 2, r["status"] == 2)
        ok = r["status"] == 0
        last = int(np.argmin(ok)) if not ok.all() else steps
        np.testing.assert_allclose(out["x"][:last + 1, s], r["x"][:last + 1], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(out["xbar"][:last + 1, s], r["xbar"][:last + 1], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(out["u"][:last, s], r["u"][:last], rtol=1e-6, atol=1e-6)
        for k in range(last):
            np.testing.assert_allclose(out["tubes"][k, s], r["tubes"][k], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("mode", [1, 2])
def test_warm_start_modes_give_the_cold_result(pair, mode):
    """warm_start 1 (previous x, y) and 2 (active-set hint, KKT-certified) only change the work, not the answer."""
    import tzddpc_b200 as tz
    cfg, o, t = pair
    rng = np.random.default_rng(5)
    steps, S = min(cfg.steps, 30), 64
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    cold = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True)
    hot = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, options=tz.SolverOptions(warm_start=mode))
    assert np.array_equal(cold["status"], hot["status"])
    np.testing.assert_allclose(hot["x"], cold["x"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(hot["u"], cold["u"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(hot["tubes"], cold["tubes"], rtol=1e-7, atol=1e-7)
    if mode == 2:       # after the first step almost every solve is certified from the hint without iterating
        assert hot["iters"][1:].mean() < 0.5 * max(cold["iters"][1:].mean(), 1.0)


def test_restart_after_an_infeasible_step():
    """The reference raises when a step is infeasible (tzddpc/tzddpc.py:374-375) and the run ends; with restart=True the
    scenario starts a new run from its x0.  The 5-dim example runs into its tightened constraints after ~60 steps."""
    cfg = configs.fivedim()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    rng = np.random.default_rng(0)
    steps, S = 90, 6
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    out = t.simulate(cfg.A, cfg.B, x0, noise, restart=True)
    for s in range(S):
        r = o.closed_loop(cfg.A, cfg.B, x0[s], noise[:, s])
        bad = np.flatnonzero(r["status"] == 2)
        assert len(bad), "expected the oracle run to end infeasible"
        k = bad[0]
        assert (out["status"][:k, s] == 0).all() and out["status"][k, s] == 2
        np.testing.assert_allclose(out["x"][:k + 1, s], r["x"][:k + 1], rtol=1e-6, atol=1e-6)
        # new run from x0: state after the infeasible step, then the oracle again from there
        np.testing.assert_array_equal(out["x"][k + 1, s], x0[s])
        np.testing.assert_array_equal(out["xbar"][k + 1, s], x0[s])
        np.testing.assert_array_equal(out["e"][k + 1, s], 0.0)
        r2 = o.closed_loop(cfg.A, cfg.B, x0[s], noise[k + 1:k + 11, s])
        np.testing.assert_allclose(out["x"][k + 1:k + 12, s], r2["x"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("S", [1, 3, 4, 7, 21])
def test_partial_warps_in_every_solver_mode(pair, S):
    """Batches that do not fill a warp tile (dead lanes) in all three warm-start modes, with restart: regression test for a
    short-circuited warp collective in the hint path that only showed with fewer than 8 scenarios."""
    import tzddpc_b200 as tz
    cfg, o, t = pair
    steps = min(cfg.steps, 14)
    noise = np.repeat(common.noise_for(cfg, steps, 1, np.random.default_rng(0)), S, axis=1)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    r = o.closed_loop(cfg.A, cfg.B, x0[0], noise[:, 0])
    ok = r["status"] == 0
    last = int(np.argmin(ok)) if not ok.all() else steps
    for mode in (0, 1, 2):
        out = t.simulate(cfg.A, cfg.B, x0, noise, options=tz.SolverOptions(warm_start=mode), restart=True)
        for s in (0, S - 1):
            assert (out["status"][:last, s] == 0).all(), (mode, out["status"][:, s])
            np.testing.assert_allclose(out["x"][:last + 1, s], r["x"][:last + 1], rtol=1e-6, atol=1e-6)
