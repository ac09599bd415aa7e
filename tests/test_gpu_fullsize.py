"""BASELINE.json's full sizes (pulley x 4,096 and 5-dim x 65,536 scenarios on one GPU) through properties that do not
need the oracle at that size:
  * batch invariance: scenario i gives bit-identical results in a batch of 65,536 and in a batch of 64
    (= shard invariance: results do not depend on how scenarios are split over GPUs);
  * the fused kernel's Ze[1].Z equals the recomposition MdataK*Ze0 (+) Mdelta*[xbar0;v0] (+) W done with the
    stand-alone reach_step kernel (two independent CUDA paths, 1e-9), tzddpc/tzddpc.py:172-176,205;
  * tightened constraints hold at the solution: hull(Ze[1]) + xbar_1 inside X (tzddpc/tzddpc.py:191-195);
  * closed loop: e+ = x+ - xbar+, x+ = A x + B (K e + v0) + w, and the realised error lies in the tube Ze[1]
    (the robust guarantee the method is built on); the on-device statistics equal torch reductions.
A sample of the batch is also compared with the oracle directly."""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu

SIZES = {"pulley": 4096, "fivedim": 65536}


@pytest.fixture(scope="module", params=["pulley", "fivedim"])
def big(request, cuda_lib):
    import torch
    cfg = configs.CONFIGS[request.param]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    S = SIZES[request.param]
    rng = np.random.default_rng(17)
    x0 = np.asarray(cfg.X0[0], dtype=np.float64)
    # scenarios: the example's start plus a spread of nominal states / errors around it
    xb = x0[None] + rng.uniform(-0.3, 0.3, size=(S, cfg.n))
    e = rng.uniform(-0.2, 0.2, size=(S, cfg.n))
    xb[0], e[0] = x0, 0.0
    return cfg, o, t, torch, xb, e


def test_batch_invariance_and_tube_recomposition(big):
    cfg, o, t, torch, xb, e = big
    n, m, S = cfg.n, cfg.m, xb.shape[0]
    dev = t.device
    xbt = torch.as_tensor(xb.T.copy()).to(dev)
    et = torch.as_tensor(e.T.copy()).to(dev)
    r = t.solve_batch(xbt, et)
    status = r.status.cpu().numpy()
    ok = status == 0
    assert ok.mean() > 0.5 and set(np.unique(status)) <= {0, 2}
    # ---- batch invariance (bitwise)
    for sl in (slice(0, 64), slice(S - 64, S), slice(S // 2 - 7, S // 2 + 57)):
        rs = t.solve_batch(xbt[:, sl].contiguous(), et[:, sl].contiguous())
        assert torch.equal(rs.status, r.status[sl])
        for a, b in ((rs.cost, r.cost[sl]), (rs.v, r.v[:, sl]), (rs.xbar, r.xbar[:, sl]), (rs.tube._ze1, r.tube._ze1[:, sl])):
            a, b = a.cpu().numpy(), b.cpu().numpy()
            good = np.broadcast_to(ok[sl], a.shape)
            assert np.array_equal(a[good], b[good])
    # ---- Ze[1] = MdataK * <e0, 0> (+) (Mdelta * <[xbar0; v0], 0> (+) W)
    g1 = t._program.compiled.g1
    Z = r.tube.device_tensor.permute(2, 0, 1).contiguous()                     # S x n x (1+g1)
    f = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)        # noqa: E731
    Ze0 = torch.zeros((S, n, 2), dtype=torch.float64, device=dev)
    Ze0[:, :, 0] = et.t()
    XU0 = torch.zeros((S, n + m, 2), dtype=torch.float64, device=dev)
    XU0[:, :n, 0] = xbt.t()
    XU0[:, n:, 0] = r.v[:m].t()
    T1 = torch.ops.tzddpc.reach_step(f(t.MdataK.center), f(t.MdataK.generators), Ze0, None)
    Zn = torch.ops.tzddpc.reach_step(f(np.zeros((n, n + m))), f(t.Mdelta.generators), XU0, f(t.zonotopes.W.Z))
    ref = torch.cat([T1, Zn[:, :, 1:]], dim=2)
    ref[:, :, 0] += Zn[:, :, 0]
    assert ref.shape == (S, n, 1 + g1)
    okt = torch.as_tensor(ok).to(dev)
    err = (Z[okt] - ref[okt]).abs()
    scale = ref[okt].abs().clamp_min(1e-3)
    assert float((err / scale).max()) < common.GEN_RTOL
    # ---- tightened state constraint at k = 1 (tzddpc/tzddpc.py:191-195)
    lo, hi = torch.ops.tzddpc.interval_hull(Z)
    xbar1 = r.xbar[n:2 * n].t()
    Xi = t.zonotopes.X.interval
    tol = 1e-6
    assert bool(((lo + xbar1)[okt] >= f(Xi.left_limit) - tol * (1 + f(np.abs(Xi.left_limit)))).all())
    assert bool(((hi + xbar1)[okt] <= f(Xi.right_limit) + tol * (1 + f(np.abs(Xi.right_limit)))).all())
    # ---- a sample against the oracle
    wmax = t._program.compiled.wmax
    cost = r.cost.cpu().numpy()
    v = r.v.cpu().numpy()
    for i in list(range(0, S, S // 24))[:24]:
        ro = o.solve_status(xb[i], e[i])
        assert (ro.status == 2) == (status[i] == 2)
        if ro.status == 0:
            assert common.cost_close(cost[i], ro.cost, wmax)
            np.testing.assert_allclose(v[0, i], ro.v[0, 0], rtol=1e-6, atol=1e-6)


def test_closed_loop_properties_full_size(big):
    cfg, o, t, torch, xb, e = big
    n, m, S = cfg.n, cfg.m, xb.shape[0]
    steps = 6
    rng = np.random.default_rng(3)
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    out = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True)
    st = out["status"]
    assert set(np.unique(st)) <= {0, 2}
    K = t.theta.K
    alive = np.ones(S, dtype=bool)
    for k in range(steps):
        okk = st[k] == 0
        alive &= okk
        x, xbar, ee = out["x"][k], out["xbar"][k], out["e"][k]
        u = ee @ K.T + out["v"][k][:, :m]
        np.testing.assert_allclose(out["u"][k][alive], u[alive], rtol=1e-12, atol=1e-12)
        xn = x @ cfg.A.T + u @ cfg.B.T + noise[k]
        np.testing.assert_allclose(out["x"][k + 1][alive], xn[alive], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(out["e"][k + 1][alive], (out["x"][k + 1] - out["xbar"][k + 1])[alive])
        # realised error inside the predicted tube Ze[1] = <c, G>: |e+ - c| <= sum |G|
        Zk = out["tubes"][k]
        c, rad = Zk[:, :, 0], np.abs(Zk[:, :, 1:]).sum(axis=2)
        assert np.all(np.abs(out["e"][k + 1][alive] - c[alive]) <= rad[alive] + 1e-9)
        # statistics row: [sum |x+|, sum |x+|^2, sum cost, #infeasible, #maxiter, sum iters, #nonfinite, S]
        nrm = np.linalg.norm(out["x"][k + 1][okk], axis=1)
        np.testing.assert_allclose(out["stats"][k, 0], nrm.sum(), rtol=1e-10)
        np.testing.assert_allclose(out["stats"][k, 1], (nrm ** 2).sum(), rtol=1e-10)
        np.testing.assert_allclose(out["stats"][k, 2], out["cost"][k][okk].sum(), rtol=1e-10)
        assert out["stats"][k, 3] == (st[k] == 2).sum() and out["stats"][k, 7] == S
        assert out["stats"][k, 5] == out["iters"][k].sum()
    assert alive.mean() > 0.99
    # every scenario saw the same x0 and the loop is deterministic: scenario 0 against the oracle
    ro = o.closed_loop(cfg.A, cfg.B, x0[0], noise[:, 0])
    np.testing.assert_allclose(out["x"][:, 0], ro["x"], rtol=1e-6, atol=1e-6)


def test_bench_configuration_hint_mode_equals_cold(cuda_lib):
    """What bench.py times: 65,536 scenarios of the 5-dim example in closed loop with the active-set hint and restart of
    infeasible scenarios, through the wave in which the example runs into its constraints (steps ~45-66).  The hint only
    changes the work: states, inputs and statuses equal those of the cold solver, step by step."""
    import tzddpc_b200 as tz
    cfg = configs.fivedim()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    S, steps = 65536, 72
    rng = np.random.default_rng(8)
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    cold = t.simulate(cfg.A, cfg.B, x0, noise, restart=True)
    hot = t.simulate(cfg.A, cfg.B, x0, noise, restart=True, options=tz.SolverOptions(warm_start=2))
    assert np.array_equal(cold["status"], hot["status"])
    assert set(np.unique(cold["status"])) <= {0, 2}
    assert (cold["status"] == 2).sum() > S // 2                      # the wave is inside the window
    np.testing.assert_allclose(hot["x"], cold["x"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(hot["u"], cold["u"], rtol=1e-7, atol=1e-7, equal_nan=True)
    assert hot["iters"].mean() < 0.2 * cold["iters"].mean()
    # restarted scenarios are back at x0 right after their infeasible step
    k, s = np.argwhere(cold["status"] == 2)[0]
    np.testing.assert_array_equal(hot["x"][k + 1, s], x0[s])
    # scenario 0 against the oracle up to its own infeasible step
    r = o.closed_loop(cfg.A, cfg.B, x0[0], noise[:, 0])
    last = int(np.argmax(r["status"] == 2)) if (r["status"] == 2).any() else steps
    np.testing.assert_allclose(hot["x"][:last + 1, 0], r["x"][:last + 1], rtol=1e-6, atol=1e-6)
    assert hot["status"][last, 0] == 2 if last < steps else True
