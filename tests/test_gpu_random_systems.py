"""Random systems of other sizes than the three examples: dim_x from 1 to 8 (the kernel's maximum), dim_u 1 or 2,
quadratic + absolute-value costs, user boxes -- solve and closed loop against the oracle; empty batches."""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


def _random_config(n, m, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    A *= 0.9 / max(np.abs(np.linalg.eigvals(A)).max(), 1e-9)
    B = rng.normal(size=(n, m))
    w = np.zeros(n); w[0] = 2.0
    r = rng.uniform(-0.5, 0.5, n)
    lo = np.full(n, -np.inf); hi = np.full(n, np.inf)
    if n >= 2:
        lo[1], hi[1] = -3.0, 3.0
    return configs.ExampleConfig(f"random{n}x{m}", A, B, X0=(np.zeros(n), np.zeros((n, 1))), U=(np.zeros(m), 2.0 * np.eye(m)),
                                 W=(np.zeros(n), 0.02 * np.ones((n, 1))), X=(np.zeros(n), 4.0 * np.eye(n)), T=60 + 20 * n,
                                 horizon=2, steps=10, cost=dict(Q=0.5 * np.eye(n), w_abs=w, x_ref=r),
                                 box=dict(x_lo=lo, x_hi=hi), noise="sample", seed=seed)


@pytest.mark.parametrize("n,m", [(1, 1), (3, 1), (3, 2), (6, 1), (6, 2), (8, 1), (8, 2)])
def test_random_system_matches_oracle(cuda_lib, n, m):
    cfg = _random_config(n, m, 100 * n + m)
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    rng = np.random.default_rng(n + m)
    S = 24
    xb = rng.uniform(-1.0, 1.0, size=(S, n))
    e = rng.uniform(-0.05, 0.05, size=(S, n))
    cost, v, xbar, tube, status = t.solve(xb, e)
    Z = tube.Z.value
    wmax = t._program.compiled.wmax
    n_ok = 0
    for i in range(S):
        r = o.solve_status(xb[i], e[i])
        assert (r.status == 2) == (status[i] == 2), (i, r.status, status[i])
        if r.status == 2:
            continue
        n_ok += 1
        assert status[i] == 0
        assert common.cost_close(cost[i], r.cost, wmax), (i, cost[i], r.cost)
        np.testing.assert_allclose(xbar[i, 1], r.xbar[1], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(Z[i], o.evaluate_tube(xb[i], e[i], v[i].ravel(), 1), rtol=common.GEN_RTOL, atol=1e-12)
    assert n_ok >= 4, f"only {n_ok} feasible points: the random configuration is too tight to be a useful test"
    # closed loop, three solver modes agree with the oracle
    import tzddpc_b200 as tz
    steps = 8
    noise = common.noise_for(cfg, steps, 3, rng)
    x0 = rng.uniform(-0.5, 0.5, size=(3, n))
    for mode in (0, 2):
        out = t.simulate(cfg.A, cfg.B, x0, noise, options=tz.SolverOptions(warm_start=mode))
        for s in range(3):
            rr = o.closed_loop(cfg.A, cfg.B, x0[s], noise[:, s])
            ok = rr["status"] == 0
            last = int(np.argmin(ok)) if not ok.all() else steps
            assert (out["status"][:last, s] == 0).all()
            np.testing.assert_allclose(out["x"][:last + 1, s], rr["x"][:last + 1], rtol=1e-6, atol=1e-6)


def test_empty_batches(cuda_lib):
    import torch
    from tzddpc_b200 import ops  # noqa: F401
    cfg = configs.pulley()
    u, x = common.dataset(cfg)
    t = common.make_product(cfg, u, x, configs.lqr_gain(cfg.A, cfg.B))
    n = cfg.n
    z = torch.zeros((n, 0), dtype=torch.float64, device="cuda")
    r = t.solve_batch(z, z)
    assert r.cost.shape == (0,) and r.status.shape == (0,) and r.v.shape[1] == 0
    lo, hi = torch.ops.tzddpc.interval_hull(torch.zeros((0, n, 5), dtype=torch.float64, device="cuda"))
    assert lo.shape == (0, n)
    lo, hi = torch.ops.tzddpc.interval_hull(torch.ones((2, n, 1), dtype=torch.float64, device="cuda"))      # no generators
    assert torch.equal(lo, hi) and float(lo.min()) == 1.0
    out, g = torch.ops.tzddpc.girard_reduce(torch.ones((2, n, 1), dtype=torch.float64, device="cuda"), 2.0, 0, 4)
    assert g.tolist() == [0, 0] and float(out[:, :, 1:].abs().max()) == 0.0
