"""The oracle against the only numerical output of the reference that exists: the closed-loop states shipped in
examples/results/pulley.xtzddpc.npy (copied to tests/golden/ by tests/golden/make_golden.py).

examples/2.pulley_sim.py is unseeded and does not save its data set, so the runs cannot be reproduced bit for
bit.  What they do pin is everything that does not depend on the data set:
  * the plant, the scalar input and the noise model w = 0.1 beta 1 (to 1e-15);
  * the closed-loop law: with the reference's gain K_r (recovered from run r) and the reference's noise draws,
    the oracle's closed loop and the reference's differ ONLY through the data-set dependent constant
    c = v0* + K(1 - xbar*), i.e. d_t = x_oracle - x_reference obeys d+ = (A + B K_r) d + B (c_oracle - c_r)
    for every t after the nominal state has settled -- checked to 5e-9 on all 5 x 190 steps;
  * constraint satisfaction and the tail statistics.
"""
import os

import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    X = np.load(os.path.join(GOLD, "ref_pulley_xtzddpc.npy"))
    d = np.load(os.path.join(GOLD, "ref_pulley_derived.npz"))
    return X, d


def test_reference_runs_follow_the_pulley_plant(ref):
    """x+ - A x in span{B, 0.1*1}: confirms A, B (examples/2.pulley_sim.py:39-42) and W.sample() = 0.1 beta 1 (:54,92)."""
    X, d = ref
    cfg = configs.pulley()
    assert X.shape == (5, cfg.steps + 1, cfg.n) and np.all(X[:, 0] == 0.0)          # X0 = <0, 0> (:51)
    for r in range(5):
        res = X[r, 1:] - X[r, :-1] @ cfg.A.T
        beta = res[:, 1:] / 0.1
        assert np.abs(beta - beta[:, :1]).max() < 1e-13          # the same beta in rows 1..3: one generator 0.1*1
        assert np.abs(beta).max() <= 1.0 + 1e-12                  # beta ~ U[-1, 1]
        recon = X[r, :-1] @ cfg.A.T + d["u"][r][:, None] * cfg.B[:, 0] + 0.1 * d["beta"][r][:, None]
        assert np.abs(recon - X[r, 1:]).max() < 1e-14
        # U = <1, 3> = [-2, 4], X = <1, 2> = [-1, 3] hold along the run (:52-53)
        assert d["u"][r].min() >= -2.0 and d["u"][r].max() <= 4.0
        assert X[r].min() >= -1.0 and X[r].max() <= 3.0
    assert np.all(d["fit_err"] < 2e-7)                            # the affine law holds to the solver's accuracy


@pytest.mark.parametrize("run", range(5))
def test_oracle_closed_loop_reproduces_reference_law(ref, run):
    X, d = ref
    cfg = configs.pulley()
    u_data, x_data = common.dataset(cfg)
    K = d["Kfit"][run][None]                                     # the reference's gain for this run
    o, _ = common.make_oracle(cfg, u_data, x_data, K=K)
    noise = 0.1 * d["beta"][run][:, None] * np.ones((1, cfg.n))   # the reference's noise realisation
    out = o.closed_loop(cfg.A, cfg.B, np.zeros(cfg.n), noise)
    assert (out["status"] == 0).all()
    x, u = out["x"], out["u"][:, 0]
    # (a) the oracle's own law is affine with the same gain once xbar has settled
    xs = out["xbar"][-1]
    assert np.abs(out["xbar"][12:] - xs).max() < 1e-8
    assert abs(xs[0] - 1.0) < 1e-9                                # N = 2 pulley step: xbar_1[0] = 1 (known answer)
    c_oracle = out["v0"][-1, 0] + float((K @ (1.0 - xs))[0])
    law = (x[12:-1] - 1.0) @ K[0] + c_oracle
    assert np.abs(law - u[12:]).max() < 1e-8
    # (b) difference to the reference run = response to the constant offset only
    dc = c_oracle - d["cfit"][run]
    assert abs(dc) < 5e-3                                         # both are 1 - sum_j Ahat[0, j] up to estimation error
    Acl = cfg.A + cfg.B @ K
    dd = x - X[run]
    rec = dd[10:-1] @ Acl.T + cfg.B[:, 0] * dc
    assert np.abs(rec - dd[11:]).max() < 5e-9
    # (c) transient: same shape, data-set level differences only
    assert np.abs(u[:10] - d["u"][run][:10]).max() < 3e-2
    assert abs(u[0] - 1.0) < 2e-2 and abs(d["u"][run][0] - 1.0) < 2e-2     # v0 = 1 / Bhat[0] from x = xbar = 0


def test_reference_tail_statistics(ref):
    """E||x_t|| of examples/2.pulley_analyse_results.ipynb (cell with 1.96 sigma / sqrt(N)): tail in [1.9, 2.1]."""
    X, _ = ref
    nrm = np.linalg.norm(X, axis=2)
    tail = nrm[:, 50:].mean()
    assert 1.9 < tail < 2.1
    assert abs(X[:, 50:, :].mean() - 1.0) < 0.05
