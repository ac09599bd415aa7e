"""The reference's shipped example set-ups as plain data (no zonotope classes, no GPU), so that
tests, the benchmark and the oracle all start from the same numbers.

  double_integrator  examples/1.double_integrator_sim.py:37-53,73   (n=2, m=1, T=100, N=2, 12 steps)
  sweep              examples/1.double_integrator_computation_complexity.py:48-55
  pulley             examples/2.pulley_sim.py:39-59,70,75            (n=4, m=1, T=400, N=2, 200 steps)
  fivedim            examples/3.5dimsystem_sim.py:29-58,71,73        (n=5, m=1, T=400, N=2, 200 steps)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np


@dataclass
class ExampleConfig:
    name: str
    A: np.ndarray
    B: np.ndarray
    X0: tuple          # (center, generators)
    U: tuple
    W: tuple
    X: tuple
    T: int             # data length
    horizon: int
    steps: int
    cost: Dict[str, np.ndarray]          # StageCost fields
    box: Optional[Dict[str, np.ndarray]] # BoxConstraint fields
    noise: str                           # 'vertex' (ex.1, :85) or 'sample' (ex.2/3, :92)
    seed: int = 25

    @property
    def n(self) -> int:
        return self.B.shape[0]

    @property
    def m(self) -> int:
        return self.B.shape[1]


def double_integrator() -> ExampleConfig:
    A = np.array([[1.0, 1.0], [0.0, 1.0]])
    B = np.array([[0.5], [1.0]])
    return ExampleConfig("double_integrator", A, B,
                         X0=(np.array([-5.0, -2.0]), np.zeros((2, 2))), U=(np.array([0.0]), np.ones((1, 1))),
                         W=(np.zeros(2), 0.1 * np.array([[1.0, 0.5], [0.5, 1.0]])),
                         X=(np.array([-4.0, 0.0]), 0.95 * np.diag([5.0, 2.5])), T=100, horizon=2, steps=12,
                         cost=dict(Q=np.eye(2), r_abs=np.array([0.01])), box=None, noise="vertex")


def sweep() -> ExampleConfig:
    c = double_integrator()
    c.name = "sweep"
    c.W = (np.zeros(2), 0.001 * np.array([[1.0, 0.5], [0.5, 1.0]]))
    c.X = (np.array([-4.0, 0.0]), 1.2 * np.diag([5.0, 2.5]))
    return c


def pulley() -> ExampleConfig:
    import scipy.signal as sig
    num, den = [0.28261, 0.50666], [1, -1.41833, 1.58939, -1.31608, 0.88642]
    ss = sig.TransferFunction(num, den, dt=0.05).to_ss()
    A, B = np.asarray(ss.A, dtype=np.float64), np.asarray(ss.B, dtype=np.float64)
    n, m = B.shape
    w = np.zeros(n); w[0] = 1.0
    r = np.zeros(n); r[0] = 1.0
    return ExampleConfig("pulley", A, B, X0=(np.zeros(n), np.zeros((n, 1))), U=(np.ones(m), 3.0 * np.ones((m, 1))),
                         W=(np.zeros(n), 0.1 * np.ones((n, 1))), X=(np.ones(n), 2.0 * np.ones((n, 1))), T=400, horizon=2,
                         steps=200, cost=dict(w_abs=w, x_ref=r), box=None, noise="sample")


def fivedim() -> ExampleConfig:
    import scipy.signal as sig
    Ac = np.array([[-1, -4, 0, 0, 0], [4, -1, 0, 0, 0], [0, 0, -3, 1, 0], [0, 0, -1, -3, 0], [0, 0, 0, 0, -2.0]])
    Bc = np.ones((5, 1))
    A, B, _, _, _ = sig.cont2discrete((Ac, Bc, np.eye(5), 0 * Bc), dt=0.05)
    n, m = B.shape
    Id = 20.0 * np.ones((n, 1)); Id[1] = 19.0
    w = np.zeros(n); w[1] = 1e9
    r = np.zeros(n); r[1] = 2.0
    lo = np.full(n, -np.inf); lo[1] = 2.0
    hi = np.full(n, np.inf); hi[1] = 10.0
    return ExampleConfig("fivedim", A, B, X0=(np.array([-2, 4, 3, -2.5, 5.5]), np.zeros((n, n))),
                         U=(np.array([7.0]), 100.0 * np.eye(m)), W=(np.zeros(n), 0.1 * np.ones((n, 1))),
                         X=(np.array([1.0, 20, 1, 1, 1]), Id), T=400, horizon=2, steps=200,
                         cost=dict(w_abs=w, x_ref=r, r_abs=np.array([0.1])), box=dict(x_lo=lo, x_hi=hi), noise="sample")


CONFIGS = {"double_integrator": double_integrator, "sweep": sweep, "pulley": pulley, "fivedim": fivedim}


def lqr_gain(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """A stabilising K for the identified centre (gain synthesis, tzddpc/utils.py:60-103, is out of scope:
    K is an INPUT of the hot path; tests, bench and oracle all use this one)."""
    from scipy.linalg import solve_discrete_are
    n, m = B.shape
    P = solve_discrete_are(A, B, np.eye(n), np.eye(m))
    return -np.linalg.solve(np.eye(m) + B.T @ P @ B, B.T @ P @ A)


def generate_dataset(cfg: ExampleConfig, rng: np.random.Generator, num_trajectories: int = 1):
    """examples/utils.py:6-45 on plain arrays, including quirk Q9 (first returned row is the origin) and
    vertex noise (`:37`).  Returns (u, x) each (T, dim).  Data synthesis is adjacent to, not on, the hot path."""
    n, m, T = cfg.n, cfg.m, cfg.T
    cU, GU = cfg.U
    cW, GW = cfg.W
    cX0, GX0 = cfg.X0
    total = T * num_trajectories
    u = (cU[None] + rng.uniform(-1, 1, size=(total, GU.shape[1])) @ GU.T).reshape(num_trajectories, T, m)
    import itertools
    nzc = [j for j in range(GW.shape[1]) if np.any(GW[:, j] != 0)]
    Wv = np.unique(np.round(np.array([cW + GW[:, nzc] @ np.array(s) for s in itertools.product((-1.0, 1.0), repeat=len(nzc))]), 14), axis=0)
    X = np.zeros((num_trajectories, T, n))
    Y = np.zeros((num_trajectories, T, n))
    for j in range(num_trajectories):
        X[j, 0] = cX0 + GX0 @ rng.uniform(-1, 1, size=GX0.shape[1])
        for i in range(1, T):
            X[j, i] = cfg.A @ X[j, i - 1] + cfg.B @ u[j, i - 1] + Wv[rng.integers(len(Wv))]
            Y[j, i] = X[j, i]
    return u.reshape(total, m), Y.reshape(total, n)
