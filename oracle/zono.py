"""Zonotope / matrix-zonotope algebra -- ORACLE (test infrastructure, parity unpinned).

Restates, in plain float64 numpy, the semantics of the third-party `pyzonotope`
and `pydatadrivenreachability` packages (absent, un-pinned: reference
`setup.py:12`) as the reference uses them; see SURVEY.md Appendix A.  Each
function cites the reference call site it serves.  [E] = evidenced by the
reference's own code/data, [R] = recalled library behaviour, kept switchable
through `Conventions`.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import NamedTuple, Sequence

import numpy as np


@dataclass(frozen=True)
class Conventions:
    """Every [R] convention of SURVEY.md App. A as an explicit switch."""
    girard_metric: str = "l1-linf"      # App. A.5: 'l1-linf' (CORA default) | 'l2' | 'l1'
    vec_order: str = "C"                # App. A.6: matrix vectorisation, 'C' (numpy flatten) | 'F' (MATLAB)
    concat_gen_major: bool = True       # App. A.7: generator index outer, column index inner
    drop_zero_generators_on_reduce: bool = True   # App. A.5: CORA nonzeroFilter before Girard
    keep_zero_box_rows: bool = True     # App. A.5: diag(d) keeps all-zero columns


DEFAULT = Conventions()


class Interval(NamedTuple):
    """`Z.interval` result [E]: tzddpc/tzddpc.py:193-197 reads .left_limit/.right_limit."""
    left_limit: np.ndarray
    right_limit: np.ndarray


def _girard_metric(G: np.ndarray, metric: str) -> np.ndarray:
    """h_j per generator column, rows accumulated r = 0..n-1 in order (the CUDA kernel
    uses the same order so that the selection below is bit-reproducible)."""
    n, g = G.shape
    a = np.abs(G)
    if metric == "l1-linf":
        s = np.zeros(g)
        mx = np.zeros(g)
        for r in range(n):
            s = s + a[r]
            mx = np.maximum(mx, a[r])
        return s - mx
    if metric == "l1":
        s = np.zeros(g)
        for r in range(n):
            s = s + a[r]
        return s
    if metric == "l2":
        s = np.zeros(g)
        for r in range(n):
            s = s + a[r] * a[r]
        return s                        # monotone in the 2-norm; no sqrt needed for selection
    raise ValueError(f"unknown Girard metric {metric!r}")


def girard_reduce_generators(G: np.ndarray, order: float, conv: Conventions = DEFAULT) -> np.ndarray:
    """Girard order reduction of a generator matrix (n x g) -> (n x g'), App. A.5 [R]
    (port of CORA reduceGirard/pickedGenerators as used by `Zonotope.reduce`,
    reference call sites: tzddpc/tzddpc.py:126-128 via MatrixZonotope.reduce and
    examples/1.double_integrator_sim.py:170).

    * all-zero generators are dropped first;
    * if g <= order*n the (filtered) matrix is returned unchanged;
    * else the nReduced = g - floor(n(order-1)) generators with the SMALLEST metric
      (ties: lowest column index first) are replaced by diag(sum |g_red|); kept
      generators stay in their original order and come first.
    """
    G = np.asarray(G, dtype=np.float64)
    n = G.shape[0]
    if conv.drop_zero_generators_on_reduce and G.shape[1]:
        G = G[:, np.any(G != 0.0, axis=0)]
    g = G.shape[1]
    if g <= order * n:
        return G.copy()
    n_unreduced = int(np.floor(n * (order - 1)))
    n_reduced = g - n_unreduced
    h = _girard_metric(G, conv.girard_metric)
    idx = np.lexsort((np.arange(g), h))            # primary h, secondary index: stable ascending
    red = np.sort(idx[:n_reduced])
    keep = np.sort(idx[n_reduced:])
    a = np.abs(G[:, red])
    d = np.zeros(n)
    for j in range(a.shape[1]):                    # column order = ascending original index
        d = d + a[:, j]
    box = np.diag(d)
    if not conv.keep_zero_box_rows:
        box = box[:, d != 0.0]
    return np.hstack([G[:, keep], box])


class Zonotope:
    """<c, G> = {c + G b : |b|_inf <= 1}, stored as Z = [c, G] (App. A.1 [E]:
    examples/1.double_integrator_sim.py:49-52,89-90)."""

    def __init__(self, center, generators, conv: Conventions = DEFAULT):
        c = np.asarray(center, dtype=np.float64).reshape(-1)
        G = np.asarray(generators, dtype=np.float64)
        if G.ndim == 1:
            G = G.reshape(c.shape[0], -1)
        assert G.shape[0] == c.shape[0], "center/generator dimension mismatch"
        self.Z = np.hstack([c[:, None], G])
        self.conv = conv

    # -- accessors [E] tzddpc/tzddpc.py:72-75,416 --------------------------------
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

    # -- App. A.3 [E] tzddpc/tzddpc.py:193-197 -----------------------------------
    @property
    def interval(self) -> Interval:
        G = self.generators
        delta = np.zeros(self.dimension)
        for j in range(G.shape[1]):                 # column order, as the CUDA hull kernel
            delta = delta + np.abs(G[:, j])
        return Interval(self.center - delta, self.center + delta)

    # -- App. A.2 [E] tzddpc/tzddpc.py:176,191-192,205 ----------------------------
    def __add__(self, other):
        if isinstance(other, Zonotope):
            return Zonotope(self.center + other.center, np.hstack([self.generators, other.generators]), self.conv)
        return Zonotope(self.center + np.asarray(other, dtype=np.float64).reshape(-1), self.generators, self.conv)

    __radd__ = __add__

    def __mul__(self, M):
        """`Z * M` means M @ Z (left multiplication) [E] tzddpc/tzddpc.py:192."""
        M = np.atleast_2d(np.asarray(M, dtype=np.float64))
        Z = M @ self.Z
        return Zonotope(Z[:, 0], Z[:, 1:], self.conv)

    def reduce(self, order: float) -> "Zonotope":
        return Zonotope(self.center, girard_reduce_generators(self.generators, order, self.conv), self.conv)

    # -- App. A.9 [E] examples/2.pulley_sim.py:68,92; examples/utils.py:27-28 ------
    def sample(self, batch_size: int = 1, rng=None) -> np.ndarray:
        rng = np.random if rng is None else rng
        beta = rng.uniform(-1.0, 1.0, size=(batch_size, self.num_generators))
        return self.center[None, :] + beta @ self.generators.T

    def compute_vertices(self) -> np.ndarray:
        """Candidate vertices c + G s, s in {-1,1}^g, duplicates removed
        (examples/utils.py:28-29,37 only index into the list uniformly)."""
        g = self.num_generators
        assert g <= 16, "vertex enumeration is exponential"
        nz = [j for j in range(g) if np.any(self.generators[:, j] != 0.0)]
        pts = []
        for s in itertools.product((-1.0, 1.0), repeat=len(nz)):
            pts.append(self.center + self.generators[:, nz] @ np.asarray(s))
        pts = np.unique(np.round(np.asarray(pts).reshape(-1, self.dimension), 14), axis=0) if pts else self.center[None]
        return pts

    def support(self, direction: np.ndarray) -> float:
        d = np.asarray(direction, dtype=np.float64)
        return float(d @ self.center + np.abs(d @ self.generators).sum())


class MatrixZonotope:
    """<C, {G_i}> (App. A.4 [E] tzddpc/utils.py:19-31,69,112,122)."""

    def __init__(self, center, generators, conv: Conventions = DEFAULT):
        self.center = np.asarray(center, dtype=np.float64)
        G = np.asarray(generators, dtype=np.float64)
        if G.size == 0:
            G = np.zeros((0,) + self.center.shape)
        assert G.ndim == 3 and G.shape[1:] == self.center.shape
        self.generators = G
        self.conv = conv

    @property
    def num_generators(self) -> int:
        return self.generators.shape[0]

    @property
    def shape(self):
        return self.center.shape

    def __add__(self, other):
        """`M + ndarray` shifts the centre [E] tzddpc/tzddpc.py:123."""
        return MatrixZonotope(self.center + np.asarray(other, dtype=np.float64), self.generators, self.conv)

    def __rmul__(self, scalar):
        s = float(scalar)
        return MatrixZonotope(s * self.center, s * self.generators, self.conv)

    def __mul__(self, other):
        if isinstance(other, Zonotope):
            Z = matzono_times_Z(self.center, self.generators, other.Z)
            return Zonotope(Z[:, 0], Z[:, 1:], other.conv)
        if isinstance(other, np.ndarray):
            # right-multiplication of centre and every generator [E] tzddpc/tzddpc.py:119
            return MatrixZonotope(self.center @ other, self.generators @ other, self.conv)
        return NotImplemented

    # -- App. A.6 [R] tzddpc/tzddpc.py:126-128 -------------------------------------
    def reduce(self, order: float) -> "MatrixZonotope":
        n, p = self.shape
        o = self.conv.vec_order
        Gv = np.stack([G.flatten(order=o) for G in self.generators], axis=1) if self.num_generators else np.zeros((n * p, 0))
        Gr = girard_reduce_generators(Gv, order, self.conv)
        gens = np.stack([Gr[:, j].reshape((n, p), order=o) for j in range(Gr.shape[1])], axis=0) if Gr.shape[1] else np.zeros((0, n, p))
        return MatrixZonotope(self.center, gens, self.conv)

    def sample(self, batch_size: int = 1, rng=None) -> np.ndarray:
        rng = np.random if rng is None else rng
        beta = rng.uniform(-1.0, 1.0, size=(batch_size, self.num_generators))
        return self.center[None] + np.tensordot(beta, self.generators, axes=(1, 0))

    def contains(self, M: np.ndarray, tol: float = 1e-9) -> bool:
        """Is M = C + sum b_i G_i for some |b|_inf <= 1?  (tzddpc/utils.py:69,99) -- LP feasibility."""
        from scipy.optimize import linprog
        N = self.num_generators
        rhs = (np.asarray(M, dtype=np.float64) - self.center).reshape(-1)
        if N == 0:
            return bool(np.max(np.abs(rhs), initial=0.0) <= tol)
        Aeq = self.generators.reshape(N, -1).T
        # min t  s.t. Aeq b = rhs, -t <= b <= t
        c = np.r_[np.zeros(N), 1.0]
        A_ub = np.block([[np.eye(N), -np.ones((N, 1))], [-np.eye(N), -np.ones((N, 1))]])
        res = linprog(c, A_ub=A_ub, b_ub=np.zeros(2 * N), A_eq=np.hstack([Aeq, np.zeros((Aeq.shape[0], 1))]),
                      b_eq=rhs, bounds=[(None, None)] * N + [(0, None)], method="highs")
        return bool(res.status == 0 and res.x[-1] <= 1.0 + tol)


def matzono_times_Z(C: np.ndarray, Gm: np.ndarray, Z: np.ndarray) -> np.ndarray:
    """MatrixZonotope x Zonotope on the stored Z = [c, G] matrices (App. A.4):
    Z_new = [C Z, G_1 Z, ..., G_N Z]  => centre C c, generators [C G, G_1 c, G_1 G, G_2 c, ...]
    ([R] for the column order, [E] for the set; zero columns are retained).
    Z may carry trailing coefficient axes (the affine zonotopes of program.py);
    the product acts on axis 0.  Reference call sites: tzddpc/tzddpc.py:175-176,181,185."""
    blocks = [np.tensordot(C, Z, axes=(1, 0))]
    for G in Gm:
        blocks.append(np.tensordot(G, Z, axes=(1, 0)))
    return np.concatenate(blocks, axis=1)


def concatenate_zonotope(W: Zonotope, N: int, conv: Conventions = DEFAULT) -> MatrixZonotope:
    """Matrix zonotope of N-step noise sequences (App. A.7 [R], tzddpc/tzddpc.py:81):
    centre = c_W tiled to n x N; one generator g_i e_j^T per (W-generator i, column j)."""
    n = W.dimension
    C = np.tile(W.center[:, None], (1, N))
    gens = []
    pairs = [(i, j) for i in range(W.num_generators) for j in range(N)] if conv.concat_gen_major \
        else [(i, j) for j in range(N) for i in range(W.num_generators)]
    for i, j in pairs:
        G = np.zeros((n, N))
        G[:, j] = W.generators[:, i]
        gens.append(G)
    return MatrixZonotope(C, np.asarray(gens).reshape(-1, n, N), conv)


def compute_LTI_matrix_zonotope(Xm: np.ndarray, Xp: np.ndarray, Um: np.ndarray, Mw: MatrixZonotope) -> MatrixZonotope:
    """M_Sigma = (X1 - M_w) pinv([X0; U0])  (App. A.7 [R] + north_star; tzddpc/tzddpc.py:83).
    Arguments are (T-1) x dim as at the call site and are transposed here."""
    X0, X1, U0 = Xm.T, Xp.T, Um.T
    D = np.vstack([X0, U0])
    P = np.linalg.pinv(D)                                  # (T-1) x (n+m)
    X1W = (-1.0 * Mw) + X1                                 # centre X1 - C_w, generators -G_w
    return X1W * P
