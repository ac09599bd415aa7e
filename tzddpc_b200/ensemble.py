"""`TZDDPCEnsemble` -- the data-set axis of the scenario batch (BASELINE.json north_star: scenarios = noise realisations
x initial states x data sets).

The reference builds one `TZDDPC` object per data set -- one model M_Sigma (`tzddpc/tzddpc.py:67-85`), one gain, one
problem (`:132-241`) -- and runs the closed loops one after another (`examples/2.pulley_sim.py:62-103` repeats the whole
script per run).  Here D controllers of identical structure are fused into a *program set* (`tz_program_set_create`):
scenarios `[begin[j], begin[j+1])` of the batch are stepped with the program of data set j, all in ONE kernel launch per
closed-loop step (`tz_closed_loop_step_set`).
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _abi, ops
from .objects import Data, SystemZonotopes
from .ops import SolverOptions
from .tzddpc import TZDDPC


class _SetProgram:
    """Duck-typed stand-in for `_abi.Program` (handle / warm_rows / compiled / tube_pattern / bucket) backed by a program set."""

    def __init__(self, pset: _abi.ProgramSet):
        p0 = pset.programs[0]
        self.pset = pset
        self.handle = pset.handle
        self.warm_rows, self.compiled, self.tube_pattern, self.bucket = p0.warm_rows, p0.compiled, p0.tube_pattern, p0.bucket


class TZDDPCEnsemble(object):
    _solve_op = staticmethod(ops.solve_set)
    _step_op = staticmethod(ops.closed_loop_step_set)

    def __init__(self, controllers: Sequence[TZDDPC], scenarios_per_dataset: Union[int, Sequence[int]]):
        """controllers: D `TZDDPC` objects on which `build_problem` has been called with the same horizon, cost and
        constraints (one per data set).  scenarios_per_dataset: scenarios of each data set (an int, or one count per data
        set; every count but the last must be a multiple of 16)."""
        assert len(controllers) >= 1
        c0 = controllers[0]
        for c in controllers:
            assert c._program is not None, "call build_problem on every controller first"
            assert (c.dim_x, c.dim_u, c.horizon, c.device) == (c0.dim_x, c0.dim_u, c0.horizon, c0.device)
        D = len(controllers)
        counts = [int(scenarios_per_dataset)] * D if np.isscalar(scenarios_per_dataset) else [int(v) for v in scenarios_per_dataset]
        assert len(counts) == D and all(v >= 0 for v in counts)
        assert all(v % 16 == 0 for v in counts[:-1]), "scenarios per data set must be a multiple of 16 (all but the last)"
        self.controllers = list(controllers)
        self.begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        with torch.cuda.device(c0.device):                # the set's entry table lives on the controllers' device
            self._program = _SetProgram(_abi.ProgramSet([c._program for c in controllers], self.begin))
        self.device, self.solver_options, self.verbose = c0.device, SolverOptions(), False
        self.dim_x, self.dim_u, self.horizon, self._dims = c0.dim_x, c0.dim_u, c0.horizon, list(c0._dims)
        self.zonotopes = c0.zonotopes
        self.num_scenarios = int(self.begin[-1])

    @classmethod
    def from_datasets(cls, datasets: Sequence[Data], zonotopes: SystemZonotopes, horizon: int, build_loss, build_constraints=None,
                      scenarios_per_dataset: Union[int, Sequence[int]] = 16, K: Optional[np.ndarray] = None, k0: Optional[int] = None,
                      device=None, tol: float = 1e-5, num_max_iterations: int = 20, num_initial_points: int = 10,
                      accuracy: float = 1e-2, confidence: float = 1e-5, seed: int = 25) -> "TZDDPCEnsemble":
        """What the reference does once per data set -- TZDDPC(data), build_zonotopes_theta, build_problem
        (examples/2.pulley_sim.py:59-75) -- for D data sets of equal length, with the device work batched over the data sets:
        ONE tz_identify launch (model + pseudo-inverse), ONE tz_gain_synthesis launch (a gain per data set; or `K`, (m, n)
        shared / (D, m, n) per data set), ONE tz_identify launch for the boxes of M_K; then the host canonicalisation and
        tz_program_create per data set."""
        import torch
        ctls = [TZDDPC(d, device=device) for d in datasets]
        c0 = ctls[0]
        D, n, m = len(ctls), c0.dim_x, c0.dim_u
        for c in ctls:
            c.verbose = False
            assert (c.dim_x, c.dim_u, c.num_samples) == (n, m, c0.num_samples), "data sets must have the same shape"
        X = torch.stack([c0._t(d.x) for d in datasets]).contiguous()
        U = torch.stack([c0._t(d.u) for d in datasets]).contiguous()
        WZ = c0._t(zonotopes.W.Z)
        AB, dAB, _, Pinv, status = ops.identify(X, U, WZ, None, True)
        if int((status != 0).sum().item()) != 0:
            raise Exception('Identification failed: [X0; U0] does not have full row rank')
        info = None
        if K is None:
            Kd, dA, dB, rho, robust, iters, st = ops.gain_synthesis(AB, Pinv, WZ, tol, num_max_iterations, num_initial_points,
                                                                    accuracy, confidence, seed, 0)
            if int((st != 0).sum().item()) != 0:
                raise Exception('Gain synthesis failed: a Riccati iteration did not converge')
            assert bool(robust.bool().all().item()), f'K is not robust with accuracy-confidence of {accuracy, 1 - confidence}'
            info = {"rho": rho.cpu().numpy(), "iterations": iters.cpu().numpy(), "robust": robust.cpu().numpy().astype(bool)}
            dA_h, dB_h = dA.cpu().numpy(), dB.cpu().numpy()
        else:
            Kh = np.asarray(K, dtype=np.float64)
            Kd = c0._t(np.broadcast_to(Kh.reshape((-1, m, n)) if Kh.ndim == 3 else Kh.reshape(1, m, n), (D, m, n)))
            dA_h, dB_h = np.zeros((D, n, n)), np.zeros((D, n, m))
        _, _, dK, _, _ = ops.identify(X, U, WZ, Kd.contiguous(), False)
        AB_h, dAB_h, dK_h, K_h = AB.cpu().numpy(), dAB.cpu().numpy(), dK.cpu().numpy(), Kd.cpu().numpy()
        from .objects import Theta
        for d, c in enumerate(ctls):
            theta = Theta(K_h[d].copy(), dA_h[d], dB_h[d])
            if not c._adopt_model(zonotopes, AB_h[d], dAB_h[d], dK_h[d], theta):
                c.build_zonotopes_theta(zonotopes, K=K_h[d])
            c._build(int(horizon), build_loss, build_constraints, k0)
        ens = cls(ctls, scenarios_per_dataset)
        ens.theta_info = info
        return ens

    @property
    def num_datasets(self) -> int:
        return len(self.controllers)

    def dataset_of(self) -> np.ndarray:
        """data set index of every scenario of the batch"""
        return np.repeat(np.arange(self.num_datasets), np.diff(self.begin))

    @property
    def K(self) -> np.ndarray:
        """(D, m, n) feedback gains theta.K of the data sets"""
        return np.stack([c.theta.K for c in self.controllers])

    _t = TZDDPC._t
    solve_batch = TZDDPC.solve_batch

    def simulate(self, A_true, B_true, x0, *args, **kwargs):
        x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
        assert x0.shape[0] == self.num_scenarios, f"x0 must have one row per scenario ({self.num_scenarios})"
        return TZDDPC.simulate(self, A_true, B_true, x0, *args, **kwargs)

    simulate.__doc__ = TZDDPC.simulate.__doc__
