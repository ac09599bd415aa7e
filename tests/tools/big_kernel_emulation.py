"""numpy emulation of csrc/tz_big.cu's algorithm (big_step_kernel: ADMM + active-set certificate) on a compiled program of
the complexity sweep, next to the oracle's interior-point answer.  TEST INFRASTRUCTURE (imports the oracle through
tests/common.py); no GPU needed.  Used to separate algorithmic behaviour from CUDA bugs when the kernel was written, and to
measure how the exits divide between the certificate and ADMM's residual test:

    python tests/tools/big_kernel_emulation.py 10 1        # horizon 10, k0 = 1  (add -v for the certificate trace)
"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import common
from tzddpc_b200 import configs
from tzddpc_b200.program import compile_program

def row_code(z, lo, hi, wabs, kk):
    c = np.zeros(len(z), dtype=int)
    for i in range(len(z)):
        if z[i] <= lo[i]: c[i] = 1 if lo[i] < hi[i] else 6
        elif z[i] >= hi[i]: c[i] = 2
        elif wabs[i] > 0: c[i] = 3 if z[i] == kk[i] else (4 if z[i] > kk[i] else 5)
    return c

def solve(prog, xbar0, e0, sp, verbose=False):
    order = np.argsort(-(prog.wabs > 0).astype(np.int64), kind="stable")
    c = float(prog.c); D = prog.D; E = prog.E[order]
    P = c * D[:, None] * prog.P * D[None, :]
    A = E[:, None] * prog.A[order] * D[None, :]
    l0, u0, kink0 = E * prog.l0[order], E * prog.u0[order], E * prog.kink0[order]
    wabs = np.where(prog.wabs[order] > 0, c * prog.wabs[order] / E, 0.0)
    R = E[:, None] * prog.R[order]
    q0 = c * D * prog.q0; Qp = c * D[:, None] * prog.Qp
    p = np.concatenate([xbar0, e0])
    w = np.concatenate([[1.0], p, np.abs(prog.Bt @ p + prog.gam)])
    lo, hi, kk = l0 + R @ w, u0 + R @ w, kink0 + R @ w
    q = q0 + Qp @ p
    c0 = prog.cc @ w + p @ prog.CC2 @ p
    if prog.Rchk.shape[0]:
        r = prog.Rchk @ w
        if (r > 1e-9 * np.maximum(1.0, np.abs(prog.Rchk[:, 0]))).any(): return 2, None, np.inf, 0
    nz, nc = A.shape[1], A.shape[0]
    def certify(code, x, y, n_iter, rounds=6):
        code = code.copy()
        for rd in range(rounds):
            ok, xk, lam, bad = certify1(code, x, y, n_iter)
            if ok: 
                if verbose: print("   refined ok after", rd)
                return ok, xk, lam
            # refine: violated rows join the set at the violated bound; wrong-sign rows leave it
            changed = False
            for kind, i in bad:
                if kind == "feas":
                    ax = A[i] @ xk
                    code[i] = 2 if ax > hi[i] else (1 if lo[i] < hi[i] else 6); changed = True
                elif kind in ("sgn1", "sgn2"):
                    code[i] = (4 if A[i] @ xk > kk[i] else 5) if wabs[i] > 0 else 0; changed = True
                elif kind in ("side4", "side5"):
                    ax = A[i] @ xk
                    code[i] = 3; changed = True
                elif kind == "kink":
                    code[i] = 4 if lam[i] < 0 else 5; changed = True
                elif kind in ("stat", "eq", "side?"):
                    return False, xk, lam
            if not changed: return False, xk, lam
            y = lam
        return False, xk, lam
    def certify1(code, x, y, n_iter):
        delta, mu = 1e-9, 1e6
        qt = q.copy()
        for i in range(nc):
            sg = 0.0
            if wabs[i] > 0:
                if code[i] == 4: sg = wabs[i]
                elif code[i] == 5: sg = -wabs[i]
                elif code[i] in (1, 2, 6):
                    b = hi[i] if code[i] == 2 else lo[i]
                    sg = wabs[i] if b > kk[i] else (-wabs[i] if b < kk[i] else 0.0)
            qt += sg * A[i]
        ia = np.isin(code, (1, 2, 3, 6))
        t = np.where(ia, mu, 0.0)
        tgt = np.where(code == 2, hi, np.where(code == 3, kk, lo)); tgt = np.where(ia, tgt, 0.0)
        lam = np.where(ia, y, 0.0)
        K = P + delta * np.eye(nz) + A.T @ (t[:, None] * A)
        xk = x.copy()
        for it in range(n_iter):
            xprev = xk
            rhs = delta * xk - qt + A[ia].T @ (mu * tgt[ia] - lam[ia])
            xk = np.linalg.solve(K, rhs)
            lam = np.where(ia, lam + mu * (A @ xk - tgt), lam)
        ax = A @ xk
        scale = max(1.0, np.abs(ax).max()); lscale = max(1.0, np.abs(lam).max()); qs = max(np.abs(qt).max(), lscale)
        ptol, ltol, etol = 1e-9 * scale, 1e-9 * lscale, 1e-8 * scale
        bad = []
        for i in range(nc):
            if lo[i] - ax[i] > ptol or ax[i] - hi[i] > ptol: bad.append(("feas", i))
            if ia[i]:
                rl = wabs[i] if (wabs[i] > 0 and code[i] in (1, 2, 6) and (hi[i] if code[i] == 2 else lo[i]) == kk[i]) else 0.0
                if abs(ax[i] - tgt[i]) > etol: bad.append(("eq", i))
                if code[i] == 2 and lam[i] < -ltol - rl: bad.append(("sgn2", i))
                if code[i] == 1 and lam[i] > ltol + rl: bad.append(("sgn1", i))
                if code[i] == 3 and abs(lam[i]) > wabs[i] * (1 + 1e-9): bad.append(("kink", i))
            elif wabs[i] > 0:
                if code[i] == 4 and not ax[i] > kk[i]: bad.append(("side4", i))
                elif code[i] == 5 and not ax[i] < kk[i]: bad.append(("side5", i))
                elif code[i] not in (4, 5): bad.append(("side?", i))
        if not (delta * np.abs(xprev - xk) <= 1e-9 * qs).all(): bad.append(("stat", float(np.abs(xprev - xk).max())))
        if verbose: print("   certify bad:", bad[:6], len(bad))
        return len(bad) == 0, xk, np.where(ia, lam, 0.0), bad
    # singleton presolve omitted
    rho = np.full(nc, sp["rho"]); y = np.zeros(nc); z = np.minimum(np.maximum(0.0, lo), hi); ysave = np.zeros(nc)
    x = np.zeros(nz)
    alpha, sigma = sp["alpha"], sp["sigma"]
    rho_act, rho_in = sp["rho"] * sp["rho_active"], sp["rho"] * sp["rho_inactive"]
    K = P + sigma * np.eye(nz) + A.T @ (rho[:, None] * A)
    next_upd, gap, next_cert, cert_gap = 2, 2, sp["cert_first"], 2
    until = sp["check_every"]; have_prev = False; switched = False
    status = 1
    for it in range(1, sp["max_iter"] + 1):
        rhs = sigma * x - q + A.T @ (rho * z - y)
        xt = np.linalg.solve(K, rhs)
        ax = A @ xt
        zr = alpha * ax + (1 - alpha) * z
        u = zr + y / rho
        pr = u.copy()
        kr = wabs > 0
        d = u - kk
        pr[kr] = kk[kr] + np.copysign(np.maximum(np.abs(d[kr]) - wabs[kr] / rho[kr], 0.0), d[kr])
        zn = np.minimum(np.maximum(pr, lo), hi)
        y = rho * (u - zn); z = zn
        x = alpha * xt + (1 - alpha) * x
        if it == next_cert:
            next_cert += cert_gap; cert_gap = (cert_gap * 3 + 1) // 2
            code = row_code(z, lo, hi, wabs, kk)
            if verbose: print("it", it, "cert try; active", int(np.isin(code, (1, 2, 3, 6)).sum()))
            ok, xc, yc = certify(code, x, y, sp["polish"] or 3)
            if ok: return 0, xc, (0.5 * xc @ P @ xc + q @ xc + (wabs * np.abs(A @ xc - kk)).sum()) / c + c0, it
        until -= 1
        if until == 0 or it == sp["max_iter"]:
            until = sp["check_every"]
            ax = A @ x
            rp = np.abs(ax - z).max(); pn = max(np.abs(ax).max(), np.abs(z).max())
            dy = y - ysave; ysave = y.copy()
            rd = np.abs(P @ x + q + A.T @ y).max()
            ep = sp["eps_abs"] + sp["eps_rel"] * pn
            ed = sp["eps_abs"] + sp["eps_rel"] * max(np.abs(P @ x).max(), np.abs(A.T @ y).max(), np.abs(q).max())
            if verbose and it % 64 == 0: print("it", it, "rp", rp, "rd", rd)
            if not np.isfinite(rp + rd): return 3, None, np.nan, it
            if rp <= ep and rd <= ed: status = 0; break
            have_prev = True
        if it == next_upd:
            gap = (gap * 3 + 1) // 2; next_upd = it + gap
            code = row_code(z, lo, hi, wabs, kk)
            rn = np.where(np.isin(code, (1, 2, 3, 6)), rho_act, rho_in)
            if (rn != rho).any() or not switched:
                rho = rn; switched = True
                K = P + sigma * np.eye(nz) + A.T @ (rho[:, None] * A)
    code = row_code(z, lo, hi, wabs, kk)
    ok, xc, yc = certify(code, x, y, sp["polish"])
    if ok: x = xc; status = 0
    return status, x, (0.5 * x @ P @ x + q @ x + (wabs * np.abs(A @ x - kk)).sum()) / c + c0, it


if __name__ == "__main__":
    horizon, k0 = int(sys.argv[1]), int(sys.argv[2])
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
    prog = common.make_compiled(cfg, o, horizon=horizon, k0=k0)
    print("nz", prog.nz, "nc", prog.nc, "npar", prog.npar, "na", prog.na, "nchk", prog.Rchk.shape[0], "nkink", int((prog.wabs>0).sum()))
    sp = dict(rho=0.1, rho_active=100.0, rho_inactive=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-6, eps_rel=1e-6, max_iter=4000, check_every=8, polish=3, cert_first=3)
    rng = np.random.default_rng(horizon * 7 + (k0 or 0))
    Xi = o.zonotopes.X.interval
    x0 = np.asarray(cfg.X0[0], dtype=np.float64)
    pts = [(x0, np.zeros(cfg.n))]
    for _ in range(7):
        pts.append((Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.3, 0.7, cfg.n), rng.uniform(-0.01, 0.01, cfg.n)))
    for xb, e in pts:
        r = o.solve_status(xb, e)
        st, xs, cost, it = solve(prog, xb, e, sp, verbose="-v" in sys.argv)
        print("oracle", r.status, r.cost, "| emul", st, cost, "iters", it)
