#!/usr/bin/env python
"""Regenerates the fixtures of tests/golden/ (run in the build container, where /root/reference is mounted):

    python tests/golden/make_golden.py [--reference /root/reference]

1. ref_pulley_xtzddpc.npy  (5 x 201 x 4 float64)  -- REFERENCE-PRODUCED closed-loop states written by
   examples/2.pulley_sim.py:100-103,147 and shipped as examples/results/pulley.xtzddpc.npy; the only numerical
   output of the reference's hot path that exists (the reference has no tests; its dependencies cvxpy /
   pyzonotope / pydatadrivenreachability cannot be installed here, so it cannot be run).  Copied verbatim.
   ref_pulley_tzddpc_times.npy: the wall times of the same 5 runs (examples/2.pulley_sim.py:97-99).
2. ref_pulley_derived.npz -- quantities recovered from (1) with nothing but the plant of
   examples/2.pulley_sim.py:39-42: the applied inputs u_t, the noise draws beta_t (w_t = 0.1 beta_t 1,
   examples/2.pulley_sim.py:54,92), and the affine closed-loop law u_t = K(x_t - 1) + c fitted for t >= 6.
3. oracle_<config>.npz -- outputs of the CPU oracle (oracle/) at fixed seeded inputs for the three shipped
   configurations: identification (centre and order-1 boxes), and `solve` at 48 parameter points
   (cost, v, xbar trajectory, Ze[1].Z, status).  They pin the oracle against accidental change (CPU test)
   and are what the CUDA path is compared with on the GPU box without re-deriving anything there.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

T_FIT = 6          # the nominal state has reached its fixed point after n + 2 steps


def derive_from_reference(X: np.ndarray, A: np.ndarray, B: np.ndarray):
    """X: runs x 201 x 4.  Rows 1..3 of B are zero, so the residual x+ - A x gives the noise in rows 1..3 and
    B u + w in row 0."""
    runs, T1, n = X.shape
    u = np.zeros((runs, T1 - 1))
    beta = np.zeros((runs, T1 - 1))
    Kfit = np.zeros((runs, n))
    cfit = np.zeros(runs)
    fit_err = np.zeros(runs)
    for r in range(runs):
        res = X[r, 1:] - X[r, :-1] @ A.T
        beta[r] = res[:, 1] / 0.1
        u[r] = (res[:, 0] - 0.1 * beta[r]) / B[0, 0]
        M = np.hstack([X[r, T_FIT:-1] - 1.0, np.ones((T1 - 1 - T_FIT, 1))])
        sol, *_ = np.linalg.lstsq(M, u[r, T_FIT:], rcond=None)
        Kfit[r], cfit[r] = sol[:n], sol[n]
        fit_err[r] = np.abs(M @ sol - u[r, T_FIT:]).max()
    return u, beta, Kfit, cfit, fit_err


def oracle_points(cfg, o, rng, count):
    Xi = o.zonotopes.X.interval
    lo, hi = Xi.left_limit.copy(), Xi.right_limit.copy()
    xb = lo + (hi - lo) * rng.uniform(0.05, 0.95, size=(count, cfg.n))
    if cfg.box:      # three quarters of the points also respect the user box (examples/3.5dimsystem_sim.py:23-26)
        lo2 = np.maximum(lo, cfg.box.get("x_lo", lo))
        hi2 = np.minimum(hi, cfg.box.get("x_hi", hi))
        k = (3 * count) // 4
        xb[:k] = lo2 + (hi2 - lo2) * rng.uniform(0.1, 0.9, size=(k, cfg.n))
    e = rng.uniform(-0.4, 0.4, size=(count, cfg.n))
    return xb, e


def oracle_fixture(name: str, count: int = 48):
    from tests import common
    from tzddpc_b200 import configs
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    rng = np.random.default_rng(2024)
    xb, e = oracle_points(cfg, o, rng, count)
    # the three examples' own starting point (x0, e = 0) goes first
    xb[0], e[0] = np.asarray(cfg.X0[0], dtype=np.float64), 0.0
    N, n, m = cfg.horizon, cfg.n, cfg.m
    g1 = o.num_generators_log[0]
    cost = np.zeros(count)
    v = np.zeros((count, N, m))
    xbar = np.zeros((count, N + 1, n))
    ze1 = np.zeros((count, n, 1 + g1))
    status = np.zeros(count, dtype=np.int32)
    for i in range(count):
        r = o.solve_status(xb[i], e[i])
        status[i] = r.status
        cost[i] = r.cost
        if r.status != 2:
            v[i], xbar[i], ze1[i] = r.v, r.xbar, r.Ze1
    return dict(u_data=u, x_data=x, K=K, AB=o.Mdata.center, AclK=o.MdataK.center,
                GD=o.Mdelta.generators, GK=o.MdataK.generators, xbar0=xb, e0=e, cost=cost, v=v, xbar=xbar, ze1=ze1,
                status=status, gens_per_step=np.asarray(o.num_generators_log))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--skip-reference", action="store_true")
    args = ap.parse_args()
    from tzddpc_b200 import configs
    if not args.skip_reference:
        res = os.path.join(args.reference, "examples", "results")
        X = np.load(os.path.join(res, "pulley.xtzddpc.npy"))
        for r in range(X.shape[0]):          # the stacked file equals the per-run files
            assert np.array_equal(X[r], np.load(os.path.join(res, f"pulley.xtzddpc.{r}.npy")))
        np.save(os.path.join(HERE, "ref_pulley_xtzddpc.npy"), X)
        np.save(os.path.join(HERE, "ref_pulley_tzddpc_times.npy"), np.load(os.path.join(res, "pulley.tzddpc_times.npy")))
        cfg = configs.pulley()
        u, beta, Kfit, cfit, fit_err = derive_from_reference(X, cfg.A, cfg.B)
        np.savez(os.path.join(HERE, "ref_pulley_derived.npz"), u=u, beta=beta, Kfit=Kfit, cfit=cfit, fit_err=fit_err)
        print("reference pulley runs:", X.shape, "affine-law fit error per run:", fit_err)
    for name in ("double_integrator", "pulley", "fivedim"):
        fx = oracle_fixture(name)
        np.savez_compressed(os.path.join(HERE, f"oracle_{name}.npz"), **fx)
        print(name, "status histogram:", np.bincount(fx["status"], minlength=3), "g1 =", fx["ze1"].shape[2] - 1)


if __name__ == "__main__":
    main()
