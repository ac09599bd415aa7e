"""Dense convex QP/LP solver for the ORACLE (test infrastructure, parity unpinned).

The reference hands its convex program to whatever conic interior-point solver
cvxpy picks by default (`tzddpc/tzddpc.py:367`, `problem_full.solve(**kw)`;
ECOS/Clarabel era, version un-pinned).  That stack is absent here, so the oracle
solves the same program -- assembled in `oracle/program.py` -- with the
textbook Mehrotra predictor-corrector method below, to ~1e-10, and the tests
cross-check it against scipy's HiGHS (LP cases) and SLSQP (QP case).

    minimise 0.5 x'Px + q'x   subject to   G x <= h
"""
from __future__ import annotations

import numpy as np


def solve_qp_ipm(P, q, G, h, tol: float = 1e-11, max_iter: int = 100, reg: float = 1e-11):
    """Returns (x, z, info) with z >= 0 the multipliers of G x <= h.
    info = {'status': 'optimal'|'max_iter', 'iters', 'res_p', 'res_d', 'gap'}."""
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    n, m = q.shape[0], h.shape[0]
    if m == 0:
        x = np.linalg.solve(P + reg * np.eye(n), -q)
        return x, np.zeros(0), {"status": "optimal", "iters": 0, "res_p": 0.0, "res_d": 0.0, "gap": 0.0}
    # row equilibration (keeps the multipliers recoverable: z_orig = z / rn)
    rn = np.maximum(np.max(np.abs(G), axis=1), 1e-300)
    rn = np.where(np.max(np.abs(G), axis=1) == 0.0, 1.0, rn)
    Gs, hs = G / rn[:, None], h / rn
    cs = max(1.0, float(np.max(np.abs(q), initial=0.0)), float(np.max(np.abs(P), initial=0.0)))
    Ps, qs = P / cs, q / cs

    x = np.zeros(n)
    s = np.maximum(hs - Gs @ x, 1.0)
    z = np.ones(m)
    I = np.eye(n)
    nq, nh = 1.0 + np.max(np.abs(qs), initial=0.0), 1.0 + np.max(np.abs(hs), initial=0.0)
    status = "max_iter"
    it = 0
    best = None
    for it in range(1, max_iter + 1):
        rd = Ps @ x + qs + Gs.T @ z
        rp = Gs @ x + s - hs
        mu = float(s @ z) / m
        res_d = np.max(np.abs(rd)) / nq
        res_p = np.max(np.abs(rp)) / nh
        merit = max(res_d, res_p, mu)
        if best is None or merit < best[0]:
            best = (merit, x.copy(), z.copy(), res_p, res_d, mu)
        if res_d <= tol and res_p <= tol and mu <= tol:
            status = "optimal"
            break
        w = z / s
        H = Ps + (Gs.T * w) @ Gs + reg * I
        try:
            L = np.linalg.cholesky(H)
        except np.linalg.LinAlgError:
            L = np.linalg.cholesky(H + 1e-8 * max(1.0, np.max(np.abs(H))) * I)

        def kkt(rc):
            rhs = -rd + Gs.T @ ((rc - z * rp) / s)
            dx = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds = -rp - Gs @ dx
            dz = (-rc - z * ds) / s
            return dx, ds, dz

        def step(ds, dz):
            a = 1.0
            neg = ds < 0
            if np.any(neg):
                a = min(a, float(np.min(-s[neg] / ds[neg])))
            neg = dz < 0
            if np.any(neg):
                a = min(a, float(np.min(-z[neg] / dz[neg])))
            return a

        dxa, dsa, dza = kkt(s * z)                              # predictor (sigma = 0)
        aa = step(dsa, dza)
        mu_aff = float((s + aa * dsa) @ (z + aa * dza)) / m
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dx, ds, dz = kkt(s * z + dsa * dza - sigma * mu)        # corrector
        a = min(1.0, 0.995 * step(ds, dz))
        x = x + a * dx
        s = s + a * ds
        z = z + a * dz
    if status != "optimal" and best is not None:
        _, x, z, res_p, res_d, mu = best
        if max(res_p, res_d, mu) <= 1e-8:
            status = "optimal"
    return x, z * cs / rn, {"status": status, "iters": it, "res_p": float(res_p), "res_d": float(res_d), "gap": float(mu)}
