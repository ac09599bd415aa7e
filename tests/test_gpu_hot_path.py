"""fast_step_kernel (one thread per scenario, closed-form KKT certificate of the hinted active set, csrc/tz_fast.cuh +
csrc/tz_cert2.cuh) against the lane-group ADMM kernel and the oracle.  `hot_path=0` sends every tile through step_kernel;
the two paths must agree BIT FOR BIT (both evaluate hint-certified scenarios with the same closed-form function, and a
scenario the hint cannot decide is solved by step_kernel either way)."""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu

KEYS = ("x", "xbar", "e", "u", "v", "cost", "status", "iters", "tubes")


@pytest.fixture(scope="module", params=["double_integrator", "pulley", "fivedim"])
def pair(request, cuda_lib):
    cfg = configs.CONFIGS[request.param]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    return cfg, o, t


def _same(a, b, tag):
    for k in KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"{tag}: {k}")


@pytest.mark.parametrize("S", [1, 15, 16, 33, 64, 257, 4096 + 48])
@pytest.mark.parametrize("packed", [0, 1])
def test_hot_path_equals_the_admm_kernel_bitwise(pair, S, packed):
    import tzddpc_b200 as tz
    cfg, o, t = pair
    rng = np.random.default_rng(100 + S)
    steps = min(cfg.steps, 40 if S <= 257 else 12)
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    hot = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, restart=True,
                     options=tz.SolverOptions(warm_start=2, hot_path=1, tube_packed=packed))
    ref = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, restart=True,
                     options=tz.SolverOptions(warm_start=2, hot_path=0, tube_packed=packed))
    _same(hot, ref, f"S={S} packed={packed}")
    # statistics are sums over scenarios accumulated with atomics: equal up to the order of the additions
    np.testing.assert_allclose(hot["stats"], ref["stats"], rtol=1e-9, atol=1e-9)
    assert (hot["stats"][:, 7] == S).all()


@pytest.mark.parametrize("knobs", [dict(TZDDPC_FAST_TPB="32"), dict(TZDDPC_FAST_TPB="64", TZDDPC_FAST_H="1"),
                                   dict(TZDDPC_FAST_TPB="64", TZDDPC_FAST_H="2"), dict(TZDDPC_FAST_TPB="64", TZDDPC_FAST_H="4"),
                                   dict(TZDDPC_FAST_TPB="128"), dict(TZDDPC_FAST_TPB="256"), dict(TZDDPC_PDL="2"), dict(TZDDPC_PDL="0")])
@pytest.mark.parametrize("packed", [0, 1])
def test_results_do_not_depend_on_the_launch_shape(pair, knobs, packed, monkeypatch):
    """CTA size, helper groups per scenario and programmatic dependent launch (the diagnostic knobs of csrc/tz_fast.cuh and
    csrc/tz_common.cuh, read per launch) change how the scenarios are laid over threads, never a scenario's arithmetic."""
    import tzddpc_b200 as tz
    cfg, o, t = pair
    S = 1000 + 2 * packed                         # ragged last block; even: the 16-byte zero stores are in play
    rng = np.random.default_rng(55)
    steps = min(cfg.steps, 14)
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    opts = tz.SolverOptions(warm_start=2, hot_path=1, tube_packed=packed)
    for k in ("TZDDPC_FAST_TPB", "TZDDPC_FAST_H", "TZDDPC_PDL"):
        monkeypatch.delenv(k, raising=False)
    ref = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, restart=True, options=opts)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    got = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, restart=True, options=opts)
    _same(got, ref, f"{knobs} packed={packed}")


def test_hot_path_matches_the_oracle_closed_loop(pair):
    import tzddpc_b200 as tz
    cfg, o, t = pair
    rng = np.random.default_rng(21)
    steps, S = min(cfg.steps, 30), 6
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    out = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, options=tz.SolverOptions(warm_start=2, hot_path=1))
    wmax = t._program.compiled.wmax
    for s in range(S):
        r = o.closed_loop(cfg.A, cfg.B, x0[s], noise[:, s], keep_tubes=True)
        assert np.array_equal(out["status"][:, s] == 2, r["status"] == 2)
        ok = r["status"] == 0
        last = int(np.argmin(ok)) if not ok.all() else steps
        np.testing.assert_allclose(out["x"][:last + 1, s], r["x"][:last + 1], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(out["u"][:last, s], r["u"][:last], rtol=1e-6, atol=1e-6)
        assert common.cost_close(out["cost"][:last, s], r["cost"][:last], wmax).all()
        for k in range(last):
            np.testing.assert_allclose(out["tubes"][k, s], r["tubes"][k], rtol=1e-6, atol=1e-6)
    # from the second step on the hint decides (no ADMM iteration)
    assert out["iters"][2:].mean() < 1.0


def test_solve_with_hints_uses_the_hot_path_and_matches_cold(pair):
    """tz_solve (no closed-loop update) with the hint buffer: second call on the same points is decided by the hints."""
    import torch
    import tzddpc_b200 as tz
    cfg, o, t = pair
    rng = np.random.default_rng(3)
    Xi = o.zonotopes.X.interval
    S = 200
    xb = Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.05, 0.95, (S, cfg.n))
    ee = rng.uniform(-0.4, 0.4, (S, cfg.n))
    xbt, eet = t._t(xb.T), t._t(ee.T)
    cold = t.solve_batch(xbt, eet)
    warm = torch.zeros((t._program.warm_rows, S), dtype=torch.float64, device=xbt.device)
    first = t.solve_batch(xbt, eet, warm=warm, options=tz.SolverOptions(warm_start=2))
    second = t.solve_batch(xbt, eet, warm=warm, options=tz.SolverOptions(warm_start=2))
    assert torch.equal(cold.status, first.status) and torch.equal(cold.status, second.status)
    ok = (cold.status == 0).cpu().numpy()
    assert ok.sum() >= 20
    # (a hint whose closed-form point -- the minimum-norm one when the cost leaves a direction free -- is infeasible is
    # not re-certified and the scenario iterates again: a few per cent of random points)
    it2 = second.iters[torch.as_tensor(ok, device=xbt.device)]
    assert float((it2 == 0).double().mean().item()) >= 0.8
    wmax = t._program.compiled.wmax
    assert common.cost_close(second.cost.cpu().numpy()[ok], cold.cost.cpu().numpy()[ok], wmax).all()
    np.testing.assert_allclose(second.v.cpu().numpy()[0][ok], cold.v.cpu().numpy()[0][ok], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(second.tube.Z.value[ok], cold.tube.Z.value[ok], rtol=1e-6, atol=1e-6)


def test_kink_row_resting_on_a_bound_away_from_its_kink(cuda_lib):
    """a row that carries both an |.| cost and a finite bound (program.py merges them when (a, r) coincide).
    Cost |xbar_1[0] - 1| with the box xbar[:, 0] <= hi < 1 puts the row on a bound BELOW its kink: the subgradient of that
    side has to enter the stationarity condition.  Cold start, hint mode (both kernels) and the oracle must agree."""
    import tzddpc_b200 as tz
    cfg = configs.pulley()
    n = cfg.n
    hi = np.full(n, np.inf); hi[0] = 0.6
    cfg.box = dict(x_hi=hi)
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    prog = t._program.compiled
    merged = np.flatnonzero((prog.wabs > 0) & np.isfinite(prog.u0))
    assert len(merged) >= 1, "expected the |.| cost row and the box row to be merged"
    rng = np.random.default_rng(1)
    steps, S = 30, 48
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    outs = {name: t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=False, options=opt)
            for name, opt in (("cold", tz.SolverOptions()), ("hint", tz.SolverOptions(warm_start=2, hot_path=0)),
                              ("hot", tz.SolverOptions(warm_start=2, hot_path=1)))}
    _ = outs["hot"]
    for k in ("x", "xbar", "e", "u", "v", "cost", "status", "iters"):
        np.testing.assert_array_equal(outs["hot"][k], outs["hint"][k], err_msg=k)
    on_bound = 0
    for s in range(0, S, 8):
        r = o.closed_loop(cfg.A, cfg.B, x0[s], noise[:, s])
        assert (r["status"] == 0).all()
        on_bound += int((np.abs(r["xbar"][1:, 0] - 0.6) < 1e-7).sum())
        for name, out in outs.items():
            assert (out["status"][:, s] == 0).all(), name
            np.testing.assert_allclose(out["x"][:, s], r["x"], rtol=1e-6, atol=1e-6, err_msg=name)
            np.testing.assert_allclose(out["u"][:, s], r["u"], rtol=1e-6, atol=1e-6, err_msg=name)
    assert on_bound > 0, "the test must drive the merged row onto its bound"
