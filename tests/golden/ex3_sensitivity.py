"""Sensitivity study of examples/3.5dimsystem_sim.py in the ORACLE's reading of the reference (test infrastructure).

The reference script loops 200 closed-loop steps with no `try` (examples/3.5dimsystem_sim.py:73-89); in the oracle's
reading of tzddpc/tzddpc.py:119-128,155-207 the run ends infeasible after ~60 steps ('Problem is unbounded',
tzddpc/tzddpc.py:374-375).  Nothing in /root/reference can tell whether the shipped script really dies there -- it ships
no result file for example 3, ends in `pdb.set_trace()` (:97-98) and builds with `horizon = 10` unused (:55) -- so this
script measures what the verdict depends on:

  * every switch of oracle.Conventions (none can matter: tzddpc/ only ever reduces to order 1, where every generator is
    boxed whatever the metric / ordering -- the sweep shows it);
  * the feedback gain K (an input of the path; the reference's own SDP + DCCP/MOSEK gain is not reproducible): LQR with
    R = 1 (what tests and bench use), R = 0.1, R = 10, and the repo's gain synthesis (oracle/gain.py);
  * 20 data-set / noise seeds;
  * the width of the model boxes: M_K and M_Delta generators scaled by kappa <= 1.  kappa = 1 is the [R] reading
    "reduce(1) boxes every generator: d = (sum|g|)(sum_j|P_j|)'" (SURVEY.md App. A.6); a library whose order-1 reduction
    over-approximates less would behave like some kappa < 1.

    python tests/golden/ex3_sensitivity.py            # writes tests/golden/ex3_sensitivity.json (~10 min on 8 cores)

Reported per variant: first infeasible step (200 = the run survives), the tightening delta_1 of state row 1 at step 0,
and xbar[1] when the run ends."""
from __future__ import annotations

import json
import os
import sys
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle                                   # noqa: E402
from oracle import gain as ogain                # noqa: E402
from tzddpc_b200 import configs                 # noqa: E402

STEPS = 200
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ex3_sensitivity.json")


def lqr(A, B, r):
    from scipy.linalg import solve_discrete_are
    n, m = B.shape
    P = solve_discrete_are(A, B, np.eye(n), r * np.eye(m))
    return -np.linalg.solve(r * np.eye(m) + B.T @ P @ B, B.T @ P @ A)


def build(seed: int, gain: str = "lqr", kappa: float = 1.0, conv: oracle.Conventions = None):
    cfg = configs.fivedim()
    rng = np.random.default_rng(cfg.seed + 101 * seed)
    u, x = configs.generate_dataset(cfg, rng)
    Z = oracle.Zonotope
    zon = oracle.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    o = oracle.OracleTZDDPC(oracle.Data(u, x), conv or oracle.Conventions())
    o.build_zonotopes(zon)
    C = o.Mdata.center
    A0, B0 = C[:, :cfg.n], C[:, cfg.n:]
    if gain == "lqr":
        K = lqr(A0, B0, 1.0)
    elif gain.startswith("lqr_r"):
        K = lqr(A0, B0, float(gain[5:]))
    elif gain == "synthesis":
        K = ogain.lqr_gain(A0, B0)[0]
    else:
        raise ValueError(gain)
    o.build_zonotopes_theta(zon, K)
    if kappa != 1.0:
        o.MdataK = oracle.MatrixZonotope(o.MdataK.center, kappa * o.MdataK.generators)
        o.Mdelta = oracle.MatrixZonotope(o.Mdelta.center, kappa * o.Mdelta.generators)
    o.build_problem(cfg.horizon, oracle.StageCost(**cfg.cost), oracle.BoxConstraint(**cfg.box))
    return cfg, o


def run(seed: int, gain: str = "lqr", kappa: float = 1.0, conv=None, steps: int = STEPS):
    cfg, o = build(seed, gain, kappa, conv)
    rng = np.random.default_rng(7000 + seed)
    cW, GW = cfg.W
    noise = cW[None] + rng.uniform(-1, 1, size=(steps, GW.shape[1])) @ GW.T
    x0 = np.asarray(cfg.X0[0], dtype=np.float64)
    # delta_1 of state row 1 at step 0: half-width of the interval hull of Ze[1] at the optimum
    r0 = o.solve_status(x0, np.zeros(cfg.n))
    delta1 = float(np.abs(r0.Ze1[1, 1:]).sum()) if r0.status != 2 else float("nan")
    # what an UNREDUCED M_Delta (reduce(1) a no-op) would give for the same z0 = [xbar0; v0]: its generators are the rank-one
    # -g P[j,:], so sum_i |G_i z0| = |g| sum_j |P_j . z0| against the boxed |g| (sum_j |P_j|) |z0|
    D = np.hstack([o.dataset.Xm, o.dataset.Um]).T
    Pinv = np.linalg.pinv(D)
    z0 = np.r_[x0, r0.v[0]] if r0.status != 2 else np.r_[x0, 0.0]
    kappa_unreduced = float(np.abs(Pinv @ z0).sum() / (np.abs(Pinv).sum(axis=0) @ np.abs(z0)))
    r = o.closed_loop(cfg.A, cfg.B, x0, noise)
    bad = np.flatnonzero(r["status"] == 2)
    first = int(bad[0]) if len(bad) else steps
    last = first if first < steps else steps
    return {"seed": seed, "gain": gain, "kappa": kappa, "first_infeasible_step": first, "delta1_row1_step0": delta1,
            "xbar1_at_end": float(r["xbar"][last, 1]), "v0_at_end": float(r["v0"][max(last - 1, 0), 0]),
            "rho_closed_loop": float(max(abs(np.linalg.eigvals(o.MdataK.center)))),
            "kappa_equivalent_of_unreduced_model_step0": kappa_unreduced}


def _job(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    kind, seed, gain, kappa, conv_kw = args
    conv = replace(oracle.Conventions(), **conv_kw) if conv_kw else None
    out = run(seed, gain, kappa, conv)
    out["variant"] = kind
    if conv_kw:
        out["conventions"] = conv_kw
    return out


def jobs(seeds):
    js = []
    for s in seeds:
        for g in ("lqr", "lqr_r0.1", "lqr_r10", "synthesis"):
            js.append(("gain", s, g, 1.0, None))
        for k in (0.9, 0.8, 0.7, 0.6, 0.5, 0.25, 0.1, 0.0):
            js.append(("kappa", s, "lqr", k, None))
    for kw in ({"girard_metric": "l2"}, {"girard_metric": "l1"}, {"vec_order": "F"}, {"concat_gen_major": False},
               {"drop_zero_generators_on_reduce": False}, {"keep_zero_box_rows": False}):
        js.append(("convention", 0, "lqr", 1.0, kw))
    return js


def summarise(rows):
    out = {}
    for kind, key in (("gain", "gain"), ("kappa", "kappa")):
        for val in sorted({r[key] for r in rows if r["variant"] == kind}, key=str):
            sel = [r for r in rows if r["variant"] == kind and r[key] == val]
            f = np.array([r["first_infeasible_step"] for r in sel])
            out[f"{kind}={val}"] = {"runs": len(sel), "alive_200": int((f >= STEPS).sum()), "first_infeasible_min": int(f.min()),
                                    "first_infeasible_median": float(np.median(f)), "first_infeasible_max": int(f.max()),
                                    "delta1_median": float(np.nanmedian([r["delta1_row1_step0"] for r in sel]))}
    ku = [r["kappa_equivalent_of_unreduced_model_step0"] for r in rows if r["variant"] == "gain" and r["gain"] == "lqr"]
    out["kappa_equivalent_of_unreduced_model_step0"] = {"min": float(np.min(ku)), "median": float(np.median(ku)), "max": float(np.max(ku))}
    conv = [r for r in rows if r["variant"] == "convention"]
    base = [r for r in rows if r["variant"] == "gain" and r["gain"] == "lqr" and r["seed"] == 0]
    out["conventions"] = {json.dumps(r["conventions"]): {"first_infeasible_step": r["first_infeasible_step"],
                                                         "delta1": r["delta1_row1_step0"]} for r in conv}
    if base:
        out["conventions"]["default"] = {"first_infeasible_step": base[0]["first_infeasible_step"], "delta1": base[0]["delta1_row1_step0"]}
    return out


if __name__ == "__main__":
    import multiprocessing as mp
    seeds = range(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
    js = jobs(seeds)
    with mp.get_context("spawn").Pool(min(len(os.sched_getaffinity(0)), 8)) as pool:
        rows = pool.map(_job, js, chunksize=1)
    doc = {"steps": STEPS, "workload": "examples/3.5dimsystem_sim.py (fivedim), oracle closed loop", "summary": summarise(rows), "rows": rows}
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc["summary"], indent=1))
