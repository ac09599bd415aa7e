"""GPU parity of the batched gain synthesis tz_gain_synthesis (compute_theta / is_gain_robust, tzddpc/utils.py:58-129)
against its numpy restatement oracle/gain.py on identical inputs and identical Philox draws."""
import numpy as np
import pytest
import torch

import oracle
from oracle import gain
from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


def _model(cfg, seed):
    u, x = common.dataset(cfg, seed=seed)
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    o.build_zonotopes(common.oracle_zonotopes(cfg))
    Pinv = np.linalg.pinv(np.vstack([o.dataset.Xm.T, o.dataset.Um.T]))
    return o.Mdata.center, Pinv


def _wz(cfg, scale=1.0):
    return np.hstack([np.asarray(cfg.W[0], dtype=float)[:, None], scale * np.asarray(cfg.W[1], dtype=float)])


@pytest.mark.parametrize("name,scale", [("double_integrator", 1.0), ("pulley", 1.0), ("fivedim", 1.0), ("fivedim", 14.0), ("fivedim", 20.0),
                                        ("double_integrator", 3.0)])
def test_gain_synthesis_matches_oracle(cuda_lib, name, scale):
    """scale > 1 inflates the noise zonotope until the adversary destabilises the first gain: the outer loop iterates."""
    from tzddpc_b200 import ops
    cfg = configs.CONFIGS[name]()
    D = 3
    models = [_model(cfg, cfg.seed + 17 * d) for d in range(D)]
    WZ = _wz(cfg, scale)
    dev = torch.device("cuda")
    AB = torch.tensor(np.stack([mdl[0] for mdl in models]), device=dev)
    Pinv = torch.tensor(np.stack([mdl[1] for mdl in models]), device=dev)
    kw = dict(tol=1e-5, max_iter=6, num_init=4, accuracy=0.03, confidence=1e-2)
    K, dA, dB, rho, robust, iters, status = ops.gain_synthesis(AB, Pinv, torch.tensor(WZ, device=dev), kw["tol"], kw["max_iter"],
                                                               kw["num_init"], kw["accuracy"], kw["confidence"], 25, 40)
    iterated = 0
    for d in range(D):
        r = gain.gain_synthesis(models[d][0], models[d][1], WZ, seed=25, dataset=40 + d, **kw)
        assert bool(status[d].item() == 0) == r["ok"]
        if not r["ok"]:
            continue
        assert int(iters[d]) == r["iters"]
        iterated += r["iters"] > 0
        np.testing.assert_allclose(K[d].cpu().numpy(), r["K"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(dA[d].cpu().numpy(), r["dA"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(dB[d].cpu().numpy(), r["dB"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(rho[d].cpu().numpy(), [r["rho0"], r["rho_adv"], r["rho_mc"]], rtol=1e-8)
        assert bool(robust[d].item()) == r["robust"]
        # independent check of the reported radii
        F0 = models[d][0][:, :cfg.n] + models[d][0][:, cfg.n:] @ r["K"]
        assert float(rho[d, 0]) == pytest.approx(np.abs(np.linalg.eigvals(F0)).max(), rel=1e-6)
    if scale > 1.0:
        assert iterated > 0, "the inflated noise zonotope was meant to exercise the outer loop"


def test_compute_theta_default_runs_the_gpu_synthesis(cuda_lib):
    """TZDDPC.build_zonotopes_theta without K: Theta comes from tz_gain_synthesis; with no outer iteration K is the LQR gain of
    the identified centre and the closed loop built on it matches the oracle built on the same K."""
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS["pulley"]()
    u, x = common.dataset(cfg)
    t = tz.TZDDPC(tz.Data(u, x))
    t.verbose = False
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    theta, _ = t.build_zonotopes_theta(zon, num_initial_points=3)
    assert t.theta_info["robust"] and t.theta_info["iterations"] == 0
    AB = t._AB
    np.testing.assert_allclose(theta.K, configs.lqr_gain(AB[:, :cfg.n], AB[:, cfg.n:]), rtol=1e-8, atol=1e-10)
    assert theta.deltaA.shape == (cfg.n, cfg.n) and theta.deltaB.shape == (cfg.n, cfg.m) and np.abs(theta.deltaA).max() > 0
    o, _ = common.make_oracle(cfg, u, x, K=theta.K)
    t.build_problem(cfg.horizon, tz.StageCost(**cfg.cost), tz.BoxConstraint())
    x0 = np.asarray(cfg.X0[0], dtype=float)
    cost, v, xbar, tube = t.solve(x0, np.zeros(cfg.n))
    r = o.solve_status(x0, np.zeros(cfg.n))
    assert common.cost_close(cost, r.cost, t._program.compiled.wmax)
    np.testing.assert_allclose(v[0], r.v[0], rtol=1e-6, atol=1e-6)
