"""Gain-synthesis helpers under the reference's names (`tzddpc/utils.py:8-129`, re-exported by `tzddpc/__init__.py:2-7`).

The reference needs cvxpy + DCCP + MOSEK for these (utils.py:5-6,37) and its gain is "any feasible point" of an LMI
(utils.py:43-56), i.e. solver dependent; here they are thin wrappers over the batched CUDA routines `tz_gain_synthesis` /
`tz_gain_adversary` (csrc/tz_gain.cu), so the NAMES, ARGUMENTS and RETURN TYPES are the reference's but
K is the LQR gain of the adversarial pair, not the reference's LMI point.  `compute_A_B`, `is_gain_robust` and
`compute_theta` need the rank-one structure of the data-driven model (generators -g_k P[j,:]), which
`TZDDPC.build_zonotopes` attaches to the MatrixZonotope it returns; a MatrixZonotope built by hand has no such structure
and is refused (no CPU fallback)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ops
from .objects import Theta
from .zonotope import MatrixZonotope

SEED = 25


def spectral_radius(X: np.ndarray) -> float:
    """Returns the spectral radius of a matrix (tzddpc/utils.py:8-11)."""
    X = np.asarray(X, dtype=np.float64)
    assert len(X.shape) == 2 and X.shape[0] == X.shape[1], 'X is not  a square matrix'
    return float(np.abs(np.linalg.eigvals(X)).max())


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("tzddpc_b200.utils needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _t(a, dev) -> torch.Tensor:
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def _structure(Mdata: MatrixZonotope):
    s = getattr(Mdata, "rank_one", None)
    if s is None:
        raise NotImplementedError("this MatrixZonotope does not carry the rank-one structure of a data-driven model "
                                  "(use the one TZDDPC.build_zonotopes returns)")
    return s


def compute_control_gain(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Compute a stabilising matrix K given a pair (A, B) (tzddpc/utils.py:43-58: an LMI feasibility point).  Here: the LQR
    gain (Q = R = I) from the DARE, solved on the GPU by the doubling algorithm of tz_gain_synthesis on a noise-free model."""
    dev = _dev()
    A, B = np.asarray(A, dtype=np.float64), np.asarray(B, dtype=np.float64)
    n, m = B.shape
    AB = _t(np.hstack([A, B])[None], dev)
    K, _, _, _, _, _, status = ops.gain_synthesis(AB, torch.zeros((1, 1, n + m), dtype=torch.float64, device=dev),
                                                  torch.zeros((n, 2), dtype=torch.float64, device=dev), 1e-5, 1, 1, 0.5, 0.5, SEED, 0)
    if int(status[0].item()) != 0:
        raise Exception('compute_control_gain: the Riccati iteration did not converge (pair not stabilisable?)')
    return K[0].cpu().numpy()


def _adversary(Mdata: MatrixZonotope, K: np.ndarray, num_init: int, accuracy: float, confidence: float):
    Pinv, WZ = _structure(Mdata)
    dev = Pinv.device
    n = Mdata.shape[0]
    K = np.asarray(K, dtype=np.float64).reshape(-1, n)
    dA, dB, rho, robust, status = ops.gain_adversary(_t(Mdata.center[None], dev), Pinv[None].contiguous(), WZ, _t(K[None], dev),
                                                     int(num_init), float(accuracy), float(confidence), SEED, 0)
    if int(status[0].item()) != 0:
        raise Exception('gain adversary failed (non-finite data)')
    return dA[0].cpu().numpy(), dB[0].cpu().numpy(), rho[0].cpu().numpy(), bool(robust[0].item())


def compute_A_B(Mdata: MatrixZonotope, K: np.ndarray, num_init: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    """Computes the adversarial matrices (A, B) for a given control gain K and set of matrices M (tzddpc/utils.py:13-41:
    argmax ||A + B K||_F over Mdata with independent beta_A, beta_B; DCCP + MOSEK there, the closed-form convex-concave
    iteration of tz_gain_adversary here, from the centre and num_init - 1 random starts)."""
    n = Mdata.shape[0]
    dA, dB, _, _ = _adversary(Mdata, K, num_init, 0.5, 0.5)
    return Mdata.center[:, :n] + dA, Mdata.center[:, n:] + dB


def is_gain_robust(Mdata: MatrixZonotope, K: np.ndarray, accuracy: float, confidence: float) -> bool:
    """tzddpc/utils.py:105-129: N = ceil(ln(1/confidence) / ln(1/(1-accuracy))) samples of Mdata, all with rho(A + BK) < 1."""
    assert np.asarray(K).shape[1] == Mdata.shape[0], 'Wrong dimensionality for K'
    assert accuracy > 0 and accuracy < 1, 'Accuracy should be in (0,1)'
    assert confidence > 0 and confidence < 1, 'confidence should be in (0,1)'
    _, _, _, robust = _adversary(Mdata, K, 1, accuracy, confidence)
    return robust


def compute_theta(Mdata: MatrixZonotope, A0: np.ndarray, B0: np.ndarray, tolerance: float = 1e-5, initial_points: int = 10,
                  max_iterations: int = 20, accuracy: float = 1e-2, confidence: float = 1e-5) -> Theta:
    """tzddpc/utils.py:60-103 with the reference's argument list.  (A0, B0) must be the centre of Mdata (what
    TZDDPC.compute_theta passes, tzddpc/tzddpc.py:90-93)."""
    Pinv, WZ = _structure(Mdata)
    dev = Pinv.device
    n = Mdata.shape[0]
    assert np.allclose(np.hstack([A0, B0]), Mdata.center), 'Mdata does not contain (A0,B0)'
    K, dA, dB, rho, robust, iters, status = ops.gain_synthesis(_t(Mdata.center[None], dev), Pinv[None].contiguous(), WZ,
                                                               float(tolerance), int(max_iterations), int(initial_points),
                                                               float(accuracy), float(confidence), SEED, 0)
    if int(status[0].item()) != 0:
        raise Exception('Gain synthesis failed: the Riccati iteration of the identified pair did not converge')
    print(f'Optimization completed. Closed loop spectral radius: {float(max(rho[0, 0], rho[0, 1]))} - K {K[0].flatten().tolist()}')
    assert bool(robust[0].item()), f'K is not robust with accuracy-confidence of {accuracy, 1 - confidence}'
    return Theta(K[0].cpu().numpy(), dA[0].cpu().numpy(), dB[0].cpu().numpy())
