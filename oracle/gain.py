"""ORACLE (test infrastructure) -- numpy restatement of the batched gain synthesis `tz_gain_synthesis`
(tzddpc_b200/csrc/tz_gain.cu), the counterpart of the reference's `compute_theta` (tzddpc/utils.py:58-103).

**Parity unpinned, and K differs from the reference by construction** (opt-in): the reference needs
cvxpy + DCCP + MOSEK (tzddpc/utils.py:5-6,37), absent here, and its gain is "any feasible point" of an LMI
(utils.py:43-56), i.e. solver-dependent.  What is kept is the structure of compute_theta:

    An, Bn = A0, B0
    repeat (utils.py:77-94):
        K      = a stabilising gain of (An, Bn)            [reference: LMI feasibility; here: LQR gain of the DARE, Q = R = I]
        An, Bn = argmax ||A + B K||_F over M_Sigma         [reference: DCCP from `initial_points` random starts with
                 with independent beta_A, beta_B (:19-35)   MOSEK; here: the same convex-concave iteration in closed
                                                            form -- the linearised problem is maximised by
                                                            beta = sign(gradient) -- from the centre and from
                                                            Philox-seeded random starts]
        lambda_max = max(rho(An + Bn K), rho(A0 + B0 K))
    until lambda_max < 1 or |lambda_max - previous| < tolerance or max_iterations
    is_gain_robust (utils.py:105-129): N = ceil(ln(1/confidence) / ln(1/(1-accuracy))) samples of M_Sigma, all rho < 1

M_Sigma has the rank-one generators -g_k P[j,:], so A + B K = F0 - sum_k g_k r_k' with
r_k = sum_j betaA_kj P[j,:n] + betaB_kj (P[j,n:] K): nothing of size (T-1) x n x (n+m) is ever formed.
Spectral radii are computed by repeated squaring (Gelfand's formula), the form the CUDA kernel uses; it is pinned
against numpy's eigenvalues in tests/test_oracle_gain.py.
"""
from __future__ import annotations

import numpy as np

from . import philox

SQUARINGS = 30
PURPOSE_ADVERSARY, PURPOSE_ROBUST = 4, 5


def spectral_radius(M: np.ndarray, squarings: int = SQUARINGS) -> float:
    """rho(M) = lim ||M^(2^k)||_F^(1/2^k) (tzddpc/utils.py:8-11 uses eigvals)."""
    X = np.array(M, dtype=np.float64)
    logr, w = 0.0, 1.0
    for _ in range(squarings):
        s = np.sqrt((X * X).sum())
        if not np.isfinite(s):
            return np.inf
        if s == 0.0:
            return 0.0
        logr += w * np.log(s)
        X = X / s
        X = X @ X
        w *= 0.5
    s = np.sqrt((X * X).sum())
    if s == 0.0:
        return 0.0
    return float(np.exp(logr + w * np.log(s)))


def lqr_gain(A: np.ndarray, B: np.ndarray, max_iter: int = 60, tol: float = 1e-15):
    """K = -(I + B'PB)^-1 B'PA with P the stabilising solution of the DARE (Q = I, R = I), by the structure-preserving
    doubling algorithm:  W = (I + G H)^-1;  A <- A W A;  G <- G + A W G A';  H <- H + A' H W A;  H -> P."""
    n, m = B.shape
    Ak, G, H = A.copy(), B @ B.T, np.eye(n)
    ok = True
    for _ in range(max_iter):
        if not (np.all(np.isfinite(G)) and np.all(np.isfinite(H)) and np.abs(H).max() < 1e150):
            ok = False
            break
        W = np.linalg.inv(np.eye(n) + G @ H)
        AW = Ak @ W
        A1 = AW @ Ak
        G1 = G + AW @ G @ Ak.T
        H1 = H + Ak.T @ H @ W @ Ak
        with np.errstate(over="ignore", invalid="ignore"):
            dn, hn = np.sqrt(((H1 - H) ** 2).sum()), np.sqrt((H1 ** 2).sum())
        Ak, G, H = A1, G1, H1
        if not np.isfinite(hn):
            ok = False
            break
        if dn <= tol * hn:
            break
    P = H
    K = -np.linalg.solve(np.eye(m) + B.T @ P @ B, B.T @ P @ A)
    return K, ok and bool(np.all(np.isfinite(K)))


def _neg_sign(x):
    return np.where(x > 0.0, -1.0, 1.0)


def adversary(A0, B0, Pinv, GW, K, num_init: int, seed: int, dataset: int, max_ccp: int = 50):
    """argmax ||A + B K||_F over M_Sigma with independent beta_A, beta_B (tzddpc/utils.py:13-41), convex-concave iteration.
    Start 0 is the centre (beta = 0), starts 1 .. num_init-1 are uniform draws.  Returns An, Bn, ||An + Bn K||_F^2."""
    n, m = B0.shape
    gW, Tm = GW.shape[1], Pinv.shape[0]
    PA, PB = Pinv[:, :n], Pinv[:, n:]
    PBK = PB @ K
    F0 = A0 + B0 @ K
    best, best_b = -np.inf, None
    for s in range(num_init):
        if s == 0:
            bA, bB = np.zeros((gW, Tm)), np.zeros((gW, Tm))
        else:
            d = philox.draws(seed, dataset, s, PURPOSE_ADVERSARY, 2 * gW * Tm, False)
            bA, bB = d[:gW * Tm].reshape(gW, Tm), d[gW * Tm:].reshape(gW, Tm)
        for _ in range(max_ccp):
            F = F0 - GW @ (bA @ PA + bB @ PBK)
            Y = GW.T @ F                                   # row k: g_k' F
            nA, nB = _neg_sign(Y @ PA.T), _neg_sign(Y @ PBK.T)
            same = np.array_equal(nA, bA) and np.array_equal(nB, bB)
            bA, bB = nA, nB
            if same:
                break
        F = F0 - GW @ (bA @ PA + bB @ PBK)
        f = float((F * F).sum())
        if f > best:
            best, best_b = f, (bA, bB)
    bA, bB = best_b
    return A0 - GW @ (bA @ PA), B0 - GW @ (bB @ PB), best


def robust_check(A0, B0, Pinv, GW, K, accuracy: float, confidence: float, seed: int, dataset: int):
    """is_gain_robust (tzddpc/utils.py:105-129): uniform samples of M_Sigma (one beta per generator, shared by its A and B
    parts as MatrixZonotope.sample does).  Returns (robust, largest sampled spectral radius, N)."""
    n = A0.shape[0]
    gW, Tm = GW.shape[1], Pinv.shape[0]
    N = int(np.ceil(np.log(1 / confidence) / np.log(1 / (1 - accuracy))))
    Q = Pinv[:, :n] + Pinv[:, n:] @ K
    F0 = A0 + B0 @ K
    beta = philox.draws(seed, dataset, np.arange(N), PURPOSE_ROBUST, gW * Tm, False).reshape(N, gW, Tm)
    worst = 0.0
    for t in range(N):
        worst = max(worst, spectral_radius(F0 - GW @ (beta[t] @ Q)))
    return worst < 1.0, worst, N


def gain_synthesis(AB, Pinv, WZ, tol=1e-5, max_iter=20, num_init=10, accuracy=1e-2, confidence=1e-5, seed=25, dataset=0):
    """compute_theta for one data set.  AB: n x (n+m) centre of M_Sigma, Pinv: (T-1) x (n+m), WZ: n x (1+gW).
    Returns dict(K, dA, dB, rho0, rho_adv, rho_mc, robust, iters, ok)."""
    n = AB.shape[0]
    A0, B0, GW = AB[:, :n], AB[:, n:], WZ[:, 1:]
    An, Bn = A0.copy(), B0.copy()
    prev, it, ok = 0.0, 0, True
    while True:
        K, ok = lqr_gain(An, Bn)
        if not ok:
            break
        An, Bn, _ = adversary(A0, B0, Pinv, GW, K, num_init, seed, dataset)
        rho_adv, rho0 = spectral_radius(An + Bn @ K), spectral_radius(A0 + B0 @ K)
        lam = max(rho_adv, rho0)
        if abs(lam - prev) < tol or lam < 1.0:
            break
        it += 1
        prev = lam
        if it >= max_iter:
            break
    if not ok:
        return dict(K=K, dA=An - A0, dB=Bn - B0, rho0=np.nan, rho_adv=np.nan, rho_mc=np.nan, robust=False, iters=it, ok=False)
    robust, worst, _ = robust_check(A0, B0, Pinv, GW, K, accuracy, confidence, seed, dataset)
    return dict(K=K, dA=An - A0, dB=Bn - B0, rho0=rho0, rho_adv=rho_adv, rho_mc=worst, robust=robust, iters=it, ok=True)
