#!/usr/bin/env python
"""Secondary benchmark (NOT the headline bench.py): the zonotope kernels on the stress configuration of BASELINE.json
configs[4] -- synthetic 5-dim system, horizon 32, Girard order cap 10 / 20 / 30:

    python bench_stress.py [scenarios]          (default 32,768 = one GPU's shard of 262,144 over 8 GPUs)

  rollout_order*     tz_tube_rollout: product + Minkowski sum + Girard reduction + hull per step, zonotope resident in
                     shared memory; scenario-steps/s and the HBM bandwidth the un-fused chain would need for the same rate
  standalone_order*  tz_reach_step / tz_girard_reduce / tz_interval_hull on blocks of the same size: algorithmic bytes
                     (read + written once) over the CUDA-event time, and the fraction of the measured HBM peak."""
import sys, os, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tzddpc_b200 import ops  # noqa

def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

n, m = 5, 1
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rng = np.random.default_rng(0)
dK = rng.uniform(0.001, 0.02, size=(n, n)); dD = rng.uniform(0.001, 0.02, size=(n, n + m))
GK = np.zeros((n * n, n, n)); GD = np.zeros((n * (n + m), n, n + m))
for r in range(n):
    for c in range(n): GK[r * n + c, r, c] = dK[r, c]
    for c in range(n + m): GD[r * (n + m) + c, r, c] = dD[r, c]
A = rng.normal(size=(n, n)); A *= 0.85 / np.abs(np.linalg.eigvals(A)).max()
W = np.hstack([np.zeros((n, 1)), 0.1 * np.ones((n, 1))])
f = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
out = {}
for order in (10, 20, 30):
    steps = 32
    gcap = n * (order - 1) + n
    Z0 = torch.zeros((S, n, 2), dtype=torch.float64, device="cuda"); Z0[:, :, 0] = torch.rand((S, n), dtype=torch.float64, device="cuda") - 0.5
    XU = torch.rand((S, steps, n + m), dtype=torch.float64, device="cuda") * 4 - 2
    args = (f(A), f(GK), f(GD), Z0, XU, f(W), float(order), 0, gcap)
    ms = ev(lambda: torch.ops.tzddpc.tube_rollout(*args), reps=2)
    gpre = 26 * gcap + 25 + 30 + 1
    out[f"rollout_order{order}"] = {"ms": ms, "scenario_steps_per_s": S * steps / ms * 1e3, "pre_reduce_generators": gpre,
                                    "equiv_unfused_GBps": S * steps * 8 * n * (2 * gpre + 2 * gcap) / ms / 1e6}
    # stand-alone kernels at the same size
    Sz = min(S, 8192)
    Zin = torch.rand((Sz, n, 1 + gcap), dtype=torch.float64, device="cuda") - 0.5
    dA_, dGK_ = f(A), f(GK)          # (uploaded once: the timed region holds the kernel only)
    t_reach = ev(lambda: torch.ops.tzddpc.reach_step(dA_, dGK_, Zin, None))
    pre = torch.ops.tzddpc.reach_step(dA_, dGK_, Zin, None)
    bytes_reach = 8 * n * Sz * ((1 + gcap) + pre.shape[2])
    t_gir = ev(lambda: torch.ops.tzddpc.girard_reduce(pre, float(order), 0, gcap))
    bytes_gir = 8 * n * Sz * (pre.shape[2] + 1 + gcap)
    t_hull = ev(lambda: torch.ops.tzddpc.interval_hull(pre))
    bytes_hull = 8 * n * Sz * pre.shape[2]
    out[f"standalone_order{order}"] = {"reach_GBps": bytes_reach / t_reach / 1e6, "girard_GBps": bytes_gir / t_gir / 1e6,
                                       "hull_GBps": bytes_hull / t_hull / 1e6, "frac_reach": bytes_reach / t_reach / 1e6 / peak,
                                       "frac_girard": bytes_gir / t_gir / 1e6 / peak, "frac_hull": bytes_hull / t_hull / 1e6 / peak,
                                       "generators_in": int(pre.shape[2] - 1), "zonotopes": Sz}
print(json.dumps(out, indent=1))
