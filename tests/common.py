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
