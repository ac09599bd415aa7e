"""BASELINE.json configs[1]: the double-integrator complexity sweep at batch 1 -- `build_problem(N, ...)` for horizons other
than 2 and `build_problem_simplified(k0, N, ...)` (examples/1.double_integrator_computation_complexity.py:48-122), solved on
the GPU (kernel buckets B1 / B2 / B3 and, beyond 12 variables -- `-m stzddpc -ho 10 -k0 1` has 31 -- the generic
warp-per-scenario path of csrc/tz_big.cu) and compared with the oracle.  Only uniquely determined outputs
are compared (quirk Q12): cost, xbar[1], v[0], status, and Ze[1].Z at equal v."""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("horizon,k0", [(1, None), (3, None), (4, None), (5, None), (3, 1), (5, 1), (4, 2),
                                        (6, 1), (8, 1), (10, 1), (8, 2), (10, 3)])
def test_sweep_horizons_batch1(cuda_lib, horizon, k0):
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
    t = common.make_product(cfg, u, x, K, horizon=horizon, k0=k0)
    assert list(t._program.compiled.gens_per_step) == list(o.num_generators_log)        # what tzddpc/tzddpc.py:206 prints
    rng = np.random.default_rng(horizon * 7 + (k0 or 0))
    Xi = o.zonotopes.X.interval
    wmax = t._program.compiled.wmax
    done = 0
    x0 = np.asarray(cfg.X0[0], dtype=np.float64)
    pts = [(x0, np.zeros(cfg.n))]
    for _ in range(7):
        pts.append((Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.3, 0.7, cfg.n), rng.uniform(-0.01, 0.01, cfg.n)))
    for xb, e in pts:
        r = o.solve_status(xb, e)
        if r.status == 2:
            with pytest.raises(Exception, match="unbounded"):
                t.solve(xb, e)                                        # batch 1: the reference's exception (:374-375)
            continue
        cost, v, xbar, tube = t.solve(xb, e)
        assert v.shape == (horizon, cfg.m) and xbar.shape == (horizon + 1, cfg.n)
        assert common.cost_close(cost, r.cost, wmax), (cost, r.cost)
        if horizon >= 2 or k0 is not None:        # N = 1 of build_problem: the cost is xbar_0's alone (quirk Q7), v is not unique
            # generic path (more than 12 variables): degenerate LP vertices whose active set the certificate cannot close
            # leave by ADMM's residual test at eps = 1e-6, i.e. a few 1e-6 in the minimiser
            tol = 1e-6 if t._program.compiled.nz <= 12 else 2e-5
            np.testing.assert_allclose(xbar[1], r.xbar[1], rtol=tol, atol=tol)
            np.testing.assert_allclose(v[0], r.v[0], rtol=tol, atol=tol)
        if horizon >= 2:
            Zo = o.evaluate_tube(xb, e, v.ravel(), 1)
            np.testing.assert_allclose(tube.Z.value, Zo, rtol=common.GEN_RTOL, atol=1e-12)
        done += 1
    assert done >= 3


def test_program_too_large_is_reported(cuda_lib):
    """Horizons beyond the compiled buckets fail loudly with TZ_ERANGE (no silent fallback)."""
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x, horizon=2)
    with pytest.raises(RuntimeError, match="exceeds every compiled bucket"):
        common.make_product(cfg, u, x, K, horizon=14, k0=3)          # 46 variables: beyond the 32 of the generic path
