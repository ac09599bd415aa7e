"""Minimal `Zonotope` / `MatrixZonotope` / `Interval` value types.

The reference takes these from the third-party `pyzonotope` package
(`tzddpc/tzddpc.py:6`, `tzddpc/objects.py:5`), which is not vendored; the repo ships its own
so that `SystemZonotopes(X0, U, X, W)` and the example scripts keep working.  The objects
are host-side containers (numpy float64); every operation that is arithmetic on the hot
path -- interval hull, Girard reduction, MatrixZonotope x Zonotope, linear maps -- runs on
the GPU through `torch.ops.tzddpc.*` (there is no CPU implementation here).
Semantics and the [R] conventions follow SURVEY.md Appendix A.
"""
from __future__ import annotations

import itertools
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import ops

_METRICS = {"l1-linf": 0, "l1": 1, "l2": 2}


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("tzddpc_b200 needs a CUDA device: zonotope arithmetic has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _gpu(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(_dev())


class Interval(NamedTuple):
    """tzddpc/tzddpc.py:193-197 reads `.left_limit` / `.right_limit`."""
    left_limit: np.ndarray
    right_limit: np.ndarray


class Zonotope:
    """<c, G>, stored as Z = [c, G] (examples/1.double_integrator_sim.py:49-52,89-90)."""

    def __init__(self, center, generators):
        c = np.asarray(center, dtype=np.float64).reshape(-1)
        G = np.asarray(generators, dtype=np.float64)
        if G.ndim == 1:
            G = G.reshape(c.shape[0], -1)
        assert G.shape[0] == c.shape[0], "center/generator dimension mismatch"
        self.Z = np.ascontiguousarray(np.hstack([c[:, None], G]))

    @property
    def center(self) -> np.ndarray:
        return self.Z[:, 0]

    @property
    def generators(self) -> np.ndarray:
        return self.Z[:, 1:]

    @property
    def dimension(self) -> int:
        return self.Z.shape[0]

    @property
    def num_generators(self) -> int:
        return self.Z.shape[1] - 1

    @property
    def order(self) -> float:
        return self.num_generators / self.dimension

    @property
    def interval(self) -> Interval:
        lo, hi = ops.interval_hull(_gpu(self.Z)[None])
        return Interval(lo[0].cpu().numpy(), hi[0].cpu().numpy())

    def __add__(self, other):
        """Minkowski sum with a Zonotope, shift by a vector (tzddpc/tzddpc.py:176,191,205)."""
        if isinstance(other, Zonotope):
            return Zonotope(self.center + other.center, np.hstack([self.generators, other.generators]))
        return Zonotope(self.center + np.asarray(other, dtype=np.float64).reshape(-1), self.generators)

    __radd__ = __add__

    def __mul__(self, M):
        """`Z * M` is the linear map M Z (left multiplication, tzddpc/tzddpc.py:192)."""
        M = np.atleast_2d(np.asarray(M, dtype=np.float64))
        out = ops.reach_step(_gpu(M), torch.zeros((0,) + M.shape, dtype=torch.float64, device=_dev()), _gpu(self.Z)[None], None)
        Z = out[0].cpu().numpy()
        return Zonotope(Z[:, 0], Z[:, 1:])

    def reduce(self, order: float, metric: str = "l1-linf") -> "Zonotope":
        """Girard order reduction (examples/1.double_integrator_sim.py:170)."""
        n, g = self.dimension, self.num_generators
        cap = max(g, int(np.ceil(order * n)) + n)
        out, gout = ops.girard_reduce(_gpu(self.Z)[None], float(order), _METRICS[metric], cap)
        k = int(gout[0].item())
        assert k >= 0, "internal: reduction capacity too small"
        Z = out[0, :, :1 + k].cpu().numpy()
        return Zonotope(Z[:, 0], Z[:, 1:])

    def sample(self, batch_size: int = 1, rng=None) -> np.ndarray:
        """c + G b, b ~ U[-1, 1]^g (examples/2.pulley_sim.py:68,92).  Host RNG: not on the hot path."""
        rng = np.random if rng is None else rng
        beta = rng.uniform(-1.0, 1.0, size=(batch_size, self.num_generators))
        return self.center[None, :] + beta @ self.generators.T

    def compute_vertices(self) -> np.ndarray:
        """c + G s over sign patterns (examples/utils.py:28-29,37; examples/2.pulley_sim.py:56)."""
        nz = [j for j in range(self.num_generators) if np.any(self.generators[:, j] != 0.0)]
        assert len(nz) <= 16, "vertex enumeration is exponential"
        pts = [self.center + self.generators[:, nz] @ np.asarray(s) for s in itertools.product((-1.0, 1.0), repeat=len(nz))]
        return np.unique(np.round(np.asarray(pts).reshape(-1, self.dimension), 14), axis=0)

    def polygon_vertices(self) -> np.ndarray:
        """Boundary of a 2-D zonotope, counter-clockwise, as a (k, 2) array: the generators, flipped into the upper half
        plane and sorted by angle, are walked once forth and once back (O(g log g); no vertex enumeration).  Plotting
        helper of examples/1.double_integrator_sim.py:170 (`Z.reduce(min(3, Z.order)).polygon`); host side."""
        assert self.dimension == 2, "polygon export is for 2-D zonotopes"
        G = self.generators[:, np.any(self.generators != 0.0, axis=0)].T.copy()          # (g, 2)
        if G.shape[0] == 0:
            return self.center[None].copy()
        flip = (G[:, 1] < 0) | ((G[:, 1] == 0) & (G[:, 0] < 0))
        G[flip] *= -1.0
        G = G[np.argsort(np.arctan2(G[:, 1], G[:, 0]), kind="stable")]
        start = self.center - G.sum(axis=0)                                            # lowest vertex
        up = start + np.cumsum(2.0 * G, axis=0)
        down = up[-1] - np.cumsum(2.0 * G, axis=0)
        return np.vstack([start[None], up, down[:-1]])

    @property
    def polygon(self):
        """matplotlib patch of a 2-D zonotope, as pyzonotope's `Zonotope.polygon` (examples/1.double_integrator_sim.py:170)."""
        try:
            from matplotlib.patches import Polygon
        except ImportError as exc:         # matplotlib is a plotting-only dependency of the examples
            raise ImportError("Zonotope.polygon needs matplotlib; Zonotope.polygon_vertices() returns the boundary as an array") from exc
        return Polygon(self.polygon_vertices(), closed=True)

    def __repr__(self) -> str:
        return f"Zonotope(dimension={self.dimension}, num_generators={self.num_generators})"


class MatrixZonotope:
    """<C, {G_i}> (tzddpc/utils.py:19-31; tzddpc/tzddpc.py:119,123,126-128,175-176)."""

    def __init__(self, center, generators):
        self.center = np.ascontiguousarray(center, dtype=np.float64)
        G = np.asarray(generators, dtype=np.float64)
        if G.size == 0:
            G = np.zeros((0,) + self.center.shape)
        assert G.ndim == 3 and G.shape[1:] == self.center.shape
        self.generators = np.ascontiguousarray(G)

    @property
    def num_generators(self) -> int:
        return self.generators.shape[0]

    @property
    def shape(self):
        return self.center.shape

    @property
    def dimension(self) -> int:
        return self.center.shape[0]

    def __add__(self, other):
        return MatrixZonotope(self.center + np.asarray(other, dtype=np.float64), self.generators)

    def __rmul__(self, scalar):
        return MatrixZonotope(float(scalar) * self.center, float(scalar) * self.generators)

    def __mul__(self, other):
        if isinstance(other, Zonotope):
            out = ops.reach_step(_gpu(self.center), _gpu(self.generators), _gpu(other.Z)[None], None)
            Z = out[0].cpu().numpy()
            return Zonotope(Z[:, 0], Z[:, 1:])
        if isinstance(other, np.ndarray):
            # right-multiply centre and every generator (tzddpc/tzddpc.py:119): (G M)' = M' G'
            M = np.asarray(other, dtype=np.float64)
            stack = np.concatenate([self.center[None], self.generators], axis=0)          # (1+N, n, p)
            Zt = _gpu(np.transpose(stack, (0, 2, 1)))                                     # (1+N, p, n) as zonotopes
            out = ops.reach_step(_gpu(M.T), torch.zeros((0,) + M.T.shape, dtype=torch.float64, device=_dev()), Zt, None)
            res = np.transpose(out.cpu().numpy(), (0, 2, 1))
            return MatrixZonotope(res[0], res[1:])
        return NotImplemented

    def reduce(self, order: float, metric: str = "l1-linf", vec_order: str = "C") -> "MatrixZonotope":
        """Vectorise, Girard-reduce in dimension n*p, reshape (tzddpc/tzddpc.py:126-128)."""
        n, p = self.shape
        N = self.num_generators
        Gv = np.stack([G.flatten(order=vec_order) for G in self.generators], axis=1) if N else np.zeros((n * p, 0))
        zv = Zonotope(np.zeros(n * p), Gv).reduce(order, metric)
        gens = np.stack([zv.generators[:, j].reshape((n, p), order=vec_order) for j in range(zv.num_generators)], axis=0) \
            if zv.num_generators else np.zeros((0, n, p))
        return MatrixZonotope(self.center, gens)

    def sample(self, batch_size: int = 1, rng=None) -> np.ndarray:
        rng = np.random if rng is None else rng
        beta = rng.uniform(-1.0, 1.0, size=(batch_size, self.num_generators))
        return self.center[None] + np.tensordot(beta, self.generators, axes=(1, 0))

    def __repr__(self) -> str:
        return f"MatrixZonotope(shape={self.shape}, num_generators={self.num_generators})"


def boxed_generators(d: np.ndarray, vec_order: str = "C") -> np.ndarray:
    """The n*p single-entry generators d[r,c] E_rc of an order-1-reduced matrix zonotope, in the
    order Girard's diag(d) produces them for the given vectorisation."""
    n, p = d.shape
    G = np.zeros((n * p, n, p))
    for i in range(n * p):
        r, c = (i // p, i % p) if vec_order == "C" else (i % n, i // n)
        G[i, r, c] = d[r, c]
    return G


def concatenate_zonotope(W: Zonotope, N: int) -> MatrixZonotope:
    """Matrix zonotope of N-step noise sequences (tzddpc/tzddpc.py:81).  Structural
    (no arithmetic): one generator g_i e_j' per (W-generator i, column j), generator index outer."""
    n = W.dimension
    gens = np.zeros((W.num_generators * N, n, N))
    for i in range(W.num_generators):
        for j in range(N):
            gens[i * N + j, :, j] = W.generators[:, i]
    return MatrixZonotope(np.tile(W.center[:, None], (1, N)), gens)
