"""Philox4x32-10 in numpy -- ORACLE (test infrastructure) restatement of the random stream of
tzddpc_b200/csrc/tz_rng.cu, so that CPU oracle and GPU consume identical draws.

The reference draws from numpy's legacy global generator through pyzonotope
(`W.sample()`, examples/2.pulley_sim.py:92; `W.compute_vertices()[randint]`, examples/utils.py:37); that stream is
not reproducible without pyzonotope, so the new path defines its own counter-based one:

    key = (seed lo, seed hi);  counter = (scenario lo, scenario hi, t, (purpose << 16) | block)
    draw j of (scenario, t, purpose) = half (j % 2) of block (j // 2);
    uniform  beta = 2 * (u64 >> 11) * 2**-53 - 1;   vertex  beta = +1 if the top bit is set else -1
    purposes: 0 closed-loop noise, 1 data-set inputs, 2 data-set noise, 3 data-set initial state

Pinned by the published known-answer vectors of Random123 (tests/test_oracle_philox.py).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(key, counter):
    """key: (..., 2), counter: (..., 4) arrays of 32-bit values -> (..., 4) uint32."""
    k = np.asarray(key, dtype=np.uint64) & MASK
    c = np.asarray(counter, dtype=np.uint64) & MASK
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def draws(seed: int, scenario, t, purpose: int, count: int, vertex: bool) -> np.ndarray:
    """`count` draws for every (scenario, t) pair (broadcast): returns array of shape broadcast(scenario, t) + (count,)."""
    scenario = np.asarray(scenario, dtype=np.uint64)
    t = np.asarray(t, dtype=np.uint64)
    shape = np.broadcast(scenario, t).shape
    scenario, t = np.broadcast_to(scenario, shape), np.broadcast_to(t, shape)
    out = np.empty(shape + (count,), dtype=np.float64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    for blk in range((count + 1) // 2):
        ctr = np.stack([scenario & MASK, scenario >> np.uint64(32), t & MASK,
                        np.full(shape, (purpose << 16) | blk, dtype=np.uint64)], axis=-1)
        o = philox4x32_10(np.broadcast_to(key, shape + (2,)), ctr).astype(np.uint64)
        for half in range(2):
            j = 2 * blk + half
            if j >= count:
                break
            bits = (o[..., 3] << np.uint64(32)) | o[..., 2] if half else (o[..., 1] << np.uint64(32)) | o[..., 0]
            if vertex:
                out[..., j] = np.where(bits >> np.uint64(63), 1.0, -1.0)
            else:
                out[..., j] = 2.0 * ((bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0
    return out


def sample_noise(WZ: np.ndarray, S: int, vertex: bool, seed: int, scenario_offset: int, t: int) -> np.ndarray:
    """(S, n): w = c_W + G_W beta  (W.sample(), examples/2.pulley_sim.py:92 / a random vertex, examples/utils.py:37)."""
    gW = WZ.shape[1] - 1
    beta = draws(seed, scenario_offset + np.arange(S), t, 0, gW, vertex)
    return WZ[:, 0][None] + beta @ WZ[:, 1:].T


def generate_trajectories(A, B, X0Z, UZ, WZ, S: int, T: int, seed: int, scenario_offset: int = 0):
    """examples/utils.py:6-45 for S data sets of one trajectory each, quirk Q9 included.  Returns U (S,T,m), X (S,T,n)."""
    n, m = B.shape
    sid = scenario_offset + np.arange(S)
    x = X0Z[:, 0][None] + draws(seed, sid, 0, 3, X0Z.shape[1] - 1, False) @ X0Z[:, 1:].T
    U = np.zeros((S, T, m))
    X = np.zeros((S, T, n))
    for t in range(T):
        u = UZ[:, 0][None] + draws(seed, sid, t, 1, UZ.shape[1] - 1, False) @ UZ[:, 1:].T
        U[:, t] = u
        if t + 1 < T:
            w = WZ[:, 0][None] + draws(seed, sid, t, 2, WZ.shape[1] - 1, True) @ WZ[:, 1:].T
            x = x @ A.T + u @ B.T + w
            X[:, t + 1] = x
    return U, X
