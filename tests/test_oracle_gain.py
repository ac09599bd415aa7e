"""Pins for oracle/gain.py (the numpy restatement of tz_gain_synthesis; reference: tzddpc/utils.py:8-129).
The reference's gain needs cvxpy + DCCP + MOSEK and is solver-dependent, so these are independent known answers:
numpy eigenvalues, scipy's DARE solver, brute force over the vertices of a small M_Sigma."""
import itertools

import numpy as np
import pytest
from scipy.linalg import solve_discrete_are

import oracle
from oracle import gain
from tests import common
from tzddpc_b200 import configs


def test_spectral_radius_by_squaring_matches_eigenvalues():
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 5, 8):
        for _ in range(5):
            M = rng.standard_normal((n, n)) * rng.uniform(0.1, 2.0)
            assert gain.spectral_radius(M) == pytest.approx(np.abs(np.linalg.eigvals(M)).max(), rel=1e-7)
    rot = 0.97 * np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])       # complex pair
    assert gain.spectral_radius(rot) == pytest.approx(0.97, rel=1e-9)
    jordan = np.array([[0.9, 1.0, 0.0], [0.0, 0.9, 1.0], [0.0, 0.0, 0.9]])                    # defective: slowest convergence
    assert gain.spectral_radius(jordan) == pytest.approx(0.9, rel=1e-6)
    assert gain.spectral_radius(np.array([[0.0, 1.0], [0.0, 0.0]])) == 0.0                    # nilpotent
    assert gain.spectral_radius(np.zeros((3, 3))) == 0.0


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_lqr_gain_matches_scipy_dare(name):
    cfg = configs.CONFIGS[name]()
    rng = np.random.default_rng(2)
    for A, B in [(cfg.A, cfg.B), (cfg.A + 0.05 * rng.standard_normal(cfg.A.shape), cfg.B * 1.3)]:
        K, ok = gain.lqr_gain(A, B)
        assert ok
        P = solve_discrete_are(A, B, np.eye(cfg.n), np.eye(cfg.m))
        Kref = -np.linalg.solve(np.eye(cfg.m) + B.T @ P @ B, B.T @ P @ A)
        np.testing.assert_allclose(K, Kref, rtol=1e-9, atol=1e-11)
        assert gain.spectral_radius(A + B @ K) < 1.0


def test_lqr_gain_reports_unstabilisable_pairs():
    A = np.diag([1.5, 0.5])
    B = np.array([[0.0], [1.0]])           # the unstable mode is not reachable
    K, ok = gain.lqr_gain(A, B)
    assert (not ok) or gain.spectral_radius(A + B @ K) >= 1.0


def test_adversary_reaches_the_brute_force_maximum_on_a_small_model():
    rng = np.random.default_rng(3)
    n, m, Tm, gW = 2, 1, 5, 1
    A0, B0 = np.array([[1.0, 1.0], [0.0, 1.0]]), np.array([[0.5], [1.0]])
    Pinv = 0.05 * rng.standard_normal((Tm, n + m))
    GW = np.array([[0.1], [0.05]])
    K = np.array([[-0.4, -0.9]])
    An, Bn, f = gain.adversary(A0, B0, Pinv, GW, K, num_init=10, seed=25, dataset=0)
    assert f == pytest.approx(((An + Bn @ K) ** 2).sum(), rel=1e-12)
    best = -np.inf
    for bits in itertools.product([-1.0, 1.0], repeat=2 * Tm):
        bA, bB = np.array(bits[:Tm])[None], np.array(bits[Tm:])[None]
        F = A0 - GW @ (bA @ Pinv[:, :n]) + (B0 - GW @ (bB @ Pinv[:, n:])) @ K
        best = max(best, (F ** 2).sum())
    assert f <= best * (1 + 1e-12)
    assert f == pytest.approx(best, rel=1e-12)          # convex maximisation over a box: the optimum is a vertex
    assert f >= ((A0 + B0 @ K) ** 2).sum()
    # beta_A and beta_B are independent (the reference's formulation, utils.py:19-35): each part lies in its own zonotope
    box = np.abs(GW).sum(1)[:, None] * np.abs(Pinv).sum(0)[None]
    assert np.all(np.abs(np.hstack([An - A0, Bn - B0])) <= box * (1 + 1e-12))


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_gain_synthesis_on_the_examples(name):
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    o.build_zonotopes(common.oracle_zonotopes(cfg))
    AB = o.Mdata.center
    Pinv = np.linalg.pinv(np.vstack([o.dataset.Xm.T, o.dataset.Um.T]))
    WZ = np.hstack([np.asarray(cfg.W[0], dtype=float)[:, None], np.asarray(cfg.W[1], dtype=float)])
    r = gain.gain_synthesis(AB, Pinv, WZ, num_init=3, accuracy=0.05, confidence=1e-2)
    assert r["ok"] and r["robust"] and r["iters"] == 0
    assert max(r["rho0"], r["rho_adv"], r["rho_mc"]) < 1.0
    assert r["rho_adv"] >= r["rho0"] - 1e-9 or r["rho_mc"] < 1          # the adversary does not make the loop more stable
    # no outer iteration: the gain is the LQR gain of the identified centre
    np.testing.assert_allclose(r["K"], configs.lqr_gain(AB[:, :cfg.n], AB[:, cfg.n:]), rtol=1e-9, atol=1e-11)
    # the adversarial pair is a member of M_Sigma's interval hull
    Mbox = np.abs(WZ[:, 1:]).sum(1)[:, None] * np.abs(Pinv).sum(0)[None]
    assert np.all(np.abs(np.hstack([r["dA"], r["dB"]])) <= Mbox * (1 + 1e-12))
