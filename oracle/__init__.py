"""CPU oracle for the TZDDPC hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 numpy restatement of what the reference
(rssalessio/TZDDPC) computes on the path `BASELINE.json:north_star` names.
It is imported only by `tests/`, by `__graft_entry__.smoke()` and by the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- never by
`tzddpc_b200/` (the product), which must fail loudly without its CUDA library.

PARITY UNPINNED.  The reference's arithmetic lives in third-party packages that
are neither vendored nor pinned (`setup.py:12`: pyzonotope,
pydatadrivenreachability, cvxpy, dccp) and cannot be installed here (no
network), and the reference ships no tests.  The oracle therefore restates

  * the reference's own call sites, `tzddpc/tzddpc.py:45-241,357-377`, and the
    closed loop of `examples/2.pulley_sim.py:81-96`, literally, and
  * the published semantics of the libraries behind them,
    every recalled-not-evidenced convention being a switch in `Conventions`.

It is pinned only by the shipped closed-loop results in
`examples/results/pulley.xtzddpc.*.npy` (see `tests/golden/`) and by analytic
known answers (tests/test_oracle_*.py).
"""
from .zono import (Conventions, Interval, Zonotope, MatrixZonotope,            # noqa: F401
                   concatenate_zonotope, compute_LTI_matrix_zonotope)
from .program import StageCost, BoxConstraint, OracleTZDDPC, Data, SystemZonotopes, Theta  # noqa: F401
from .qp import solve_qp_ipm                                                  # noqa: F401
