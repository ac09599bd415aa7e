#!/usr/bin/env python
"""Benchmark of the TZDDPC hot path: closed-loop steps/sec over batched scenarios.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload fivedim|pulley|double_integrator]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is ONE fused closed-loop step (solve + tube + plant/nominal/error update,
examples/3.5dimsystem_sim.py:73-89) over the whole scenario batch of a GPU.  Default workload:
BASELINE.json configs[3], the 5-dim system with 65,536 scenarios (noise realisations) per GPU
("weak" scaling: every rank simulates its own 65,536 scenarios; no collective in the loop,
one NCCL all-reduce of the closed-loop statistics after it).

Prints ONE JSON line (rank 0).  `value` = scenario-steps/s with all state resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (pinned host arrays in, every output of
`solve` back on the host each step); `roofline` = algorithmic HBM bytes over the
CUDA-event duration of the fused kernel; `cpu_baseline` = the CPU oracle port timed here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "closed-loop TZDDPC steps/sec over batched scenarios"
UNIT = "scenario-steps/s"


def algorithmic_bytes_per_scenario_step(n, m, N, g1):
    """read x, xbar, e; write x+, xbar+, e+, v, xbar trajectory, cost, Ze[1].Z, status."""
    return 8 * (6 * n + N * m + (N + 1) * n + 1 + n * (1 + g1)) + 4


def measured_traffic(bucket: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the committed
    `ncu --set full` capture (profiles/step_kernel_traffic.json, written by profiles/summarise.py), or None."""
    p = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        d = json.load(open(p))
        return d.get(bucket, d.get("default"))
    except Exception:
        return None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU side: the oracle port (the reference itself cannot be imported: cvxpy/pyzonotope are absent)
# ----------------------------------------------------------------------------------------------
def _oracle_setup(workload):
    import oracle
    from tzddpc_b200 import configs
    cfg = configs.CONFIGS[workload]()
    rng = np.random.default_rng(cfg.seed)
    u, x = configs.generate_dataset(cfg, rng)
    Z = oracle.Zonotope
    zon = oracle.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    o.build_zonotopes(zon)
    C = o.Mdata.center
    K = configs.lqr_gain(C[:, :cfg.n], C[:, cfg.n:])
    o.build_zonotopes_theta(zon, K)
    box = oracle.BoxConstraint(**cfg.box) if cfg.box else None
    o.build_problem(cfg.horizon, oracle.StageCost(**cfg.cost), box)
    return cfg, o


def _oracle_worker(args):
    """Closed loop of `scen` scenarios for warmup+steps steps; returns seconds spent in the timed steps."""
    workload, scen, warmup, steps, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cfg, o = _oracle_setup(workload)
    rng = np.random.default_rng(seed)
    cW, GW = cfg.W
    n, K = cfg.n, o.theta.K
    x = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (scen, 1))
    xbar, e = x.copy(), np.zeros_like(x)
    t_timed = 0.0
    for t in range(warmup + steps):
        w = cW[None] + rng.uniform(-1, 1, size=(scen, GW.shape[1])) @ GW.T
        t0 = time.perf_counter()
        for s in range(scen):
            r = o.solve_status(xbar[s], e[s], check_feasibility=False)
            if r.status == 2:
                continue
            u = K @ e[s] + r.v[0]
            x[s] = cfg.A @ x[s] + cfg.B @ u + w[s]
            xbar[s] = r.xbar[1]
            e[s] = x[s] - xbar[s]
        if t >= warmup:
            t_timed += time.perf_counter() - t0
    return t_timed


def cpu_oracle_throughput(workload, cores, scen_per_core, warmup, steps):
    """scenario-steps/s of the oracle port on `cores` host processes (one per core)."""
    jobs = [(workload, scen_per_core, warmup, steps, 1000 + i) for i in range(cores)]
    if cores == 1:
        times = [_oracle_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(cores) as pool:
            times = pool.map(_oracle_worker, jobs)
    return cores * scen_per_core * steps / max(times), max(times)


def workload_name(cfg, S):
    return (f"{cfg.name}: n={cfg.n} m={cfg.m} T={cfg.T} horizon={cfg.horizon}, {S} scenarios per GPU (noise realisations, "
            f"shared data set), closed loop")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tzddpc_b200 import configs
    cfg = configs.CONFIGS[args.workload]()
    cores = len(os.sched_getaffinity(0))
    # a "step" here is one closed-loop step over a bounded sample of 2 scenarios per core
    scen_per_core = 2
    val, secs = cpu_oracle_throughput(args.workload, cores, scen_per_core, args.warmup, args.steps)
    sample = f"{cores * scen_per_core} scenarios x {args.steps} closed-loop steps, one process per core"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, args.scenarios), "scenarios_per_gpu": args.scenarios,
                       "note": "CPU oracle port of the reference path (the reference itself needs cvxpy / pyzonotope / "
                               "pydatadrivenreachability, absent here); each step is a bounded sample of the workload: "
                               + sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled from a thread every ~2 ms
    (the timed region of the default run lasts tens of milliseconds, too short for `nvidia-smi -lms`)."""

    def __init__(self, device_index):
        self.rows, self.ok, self._stop = [], False, False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if vis:
                ent = vis.split(",")[device_index].strip()
                if ent.startswith("GPU-"):
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(ent)
                    idx = None
                else:
                    idx = int(ent)
            if idx is not None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.rows.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                              nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
        except Exception:       # noqa: BLE001
            pass

    def _poll(self):
        while not self._stop:
            self.sample()
            time.sleep(0.002)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop = True
        self.th.join(timeout=1.0)
        nv = self.nv
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in rows))
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(sm)}


def batch1_latency(tz, ops, torch, dev, steps_per_graph=12, replays=100):
    from tzddpc_b200 import _abi, configs
    cfg = configs.double_integrator()
    rng = np.random.default_rng(cfg.seed)
    u_data, x_data = configs.generate_dataset(cfg, rng)
    ctl = tz.TZDDPC(tz.Data(u_data, x_data), device=dev)
    ctl.verbose = False
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    ctl.build_zonotopes(zon)
    n, m = cfg.n, cfg.m
    K = configs.lqr_gain(ctl.Mdata.center[:, :n], ctl.Mdata.center[:, n:])
    ctl.build_zonotopes_theta(zon, K=K)
    ctl.build_problem(cfg.horizon, tz.StageCost(**cfg.cost), tz.BoxConstraint())
    prog = ctl._program
    g1, nv, N = prog.compiled.g1, prog.compiled.nv, cfg.horizon
    f64 = dict(dtype=torch.float64, device=dev)
    S = 1
    x0 = torch.tensor(cfg.X0[0], **f64)[:, None].contiguous()
    x, xbar, e = x0.clone(), x0.clone(), torch.zeros((n, S), **f64)
    GW = torch.tensor(cfg.W[1], **f64)
    gen = torch.Generator(device=dev)
    gen.manual_seed(cfg.seed)
    beta = torch.sign(torch.rand((steps_per_graph, GW.shape[1], S), generator=gen, **f64) - 0.5)
    noise = torch.einsum("rg,tgs->trs", GW, beta).contiguous()       # a random vertex of W per step (examples/1.double_integrator_sim.py:85)
    At, Bt = torch.tensor(cfg.A, **f64), torch.tensor(cfg.B, **f64)
    status = torch.zeros((steps_per_graph, S), dtype=torch.int32, device=dev)
    cost = torch.empty((steps_per_graph, S), **f64)
    v = torch.empty((steps_per_graph, nv, S), **f64)
    traj = torch.empty((steps_per_graph, (N + 1) * n, S), **f64)
    ze1 = torch.empty((steps_per_graph, n * (1 + g1), S), **f64)
    warm = torch.zeros((prog.warm_rows, S), **f64)
    po = tz.SolverOptions(warm_start=2).pack()
    h = prog.handle.value

    def run():
        x.copy_(x0); xbar.copy_(x0); e.zero_(); warm.zero_()      # every replay is the example's 12-step run from X0
        for t in range(steps_per_graph):
            ops.closed_loop_step(h, x, xbar, e, noise[t], x0, At, Bt, status[t], cost[t], v[t], traj[t], ze1[t], None, None, warm,
                                 None, po)

    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        run()                                   # warm-up on the capture stream
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        run()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        graph.replay()
    b.record()
    torch.cuda.synchronize(dev)
    us = 1e3 * a.elapsed_time(b) / (replays * steps_per_graph)
    # the same through plain launches of the torch op (host in the loop)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(4):
        run()
    torch.cuda.synchronize(dev)
    us_eager = 1e6 * (time.perf_counter() - t0) / (4 * steps_per_graph)
    return {"us_per_step": us, "us_per_step_eager_launches": us_eager, "scenarios": 1,
            "workload": "double_integrator: n=2 m=1 T=100 horizon=2 (examples/1.double_integrator_sim.py), closed loop",
            "how": f"CUDA graph of the example's {steps_per_graph} closed-loop steps from X0, {replays} replays, CUDA events", "status_ok": bool((status == 0).all().item()),
            "kernel_bucket": prog.bucket}


def datasets_leg(args, tz, ops, torch, dev, cfg, noise, nring):
    """The default workload with the data-set axis: D data sets (seeds cfg.seed + 101 d) of the same plant, one identified
    model / gain / program each (the reference: one TZDDPC object per data set), S / D noise realisations per data set."""
    from tzddpc_b200 import _abi, configs
    D, S = args.datasets, args.scenarios
    per = S // D
    assert per % 16 == 0 and per * D == S, "scenarios must split into D blocks of a multiple of 16"
    n, m, N = cfg.n, cfg.m, cfg.horizon
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    t0 = time.perf_counter()
    data = [tz.Data(*configs.generate_dataset(cfg, np.random.default_rng(cfg.seed + 101 * d))) for d in range(D)]
    # one tz_identify + one tz_gain_synthesis launch for all data sets, then the host canonicalisation per data set
    ens = tz.TZDDPCEnsemble.from_datasets(data, zon, N, tz.StageCost(**cfg.cost),
                                          tz.BoxConstraint(**cfg.box) if cfg.box else tz.BoxConstraint(), scenarios_per_dataset=per,
                                          device=dev)
    torch.cuda.synchronize(dev)
    setup_s = time.perf_counter() - t0
    prog = ens._program
    g1, nv = prog.compiled.g1, prog.compiled.nv
    f64 = dict(dtype=torch.float64, device=dev)
    x0 = torch.tensor(cfg.X0[0], **f64)
    x = x0[:, None].repeat(1, S).contiguous()
    xbar, xr, e = x.clone(), x.clone(), torch.zeros((n, S), **f64)
    At, Bt = torch.tensor(cfg.A, **f64), torch.tensor(cfg.B, **f64)
    ze1 = torch.empty((n * (1 + g1), S), **f64)
    traj = torch.empty(((N + 1) * n, S), **f64)
    vbuf, cost = torch.empty((nv, S), **f64), torch.empty(S, **f64)
    status = torch.zeros(S, dtype=torch.int32, device=dev)
    warm = torch.zeros((prog.warm_rows, S), **f64)
    Kd = min(args.steps, 100)
    stats = torch.zeros((5 + Kd, _abi.TZ_NSTATS), **f64)
    po = tz.SolverOptions(warm_start=2, check_every=args.check_every, eps_abs=args.eps, eps_rel=args.eps, polish=args.polish).pack()
    h = prog.handle.value

    def step(t):
        ops.closed_loop_step_set(h, x, xbar, e, noise[t % nring], xr, At, Bt, status, cost, vbuf, traj, ze1, None, None, warm,
                                 stats[t], po)

    for t in range(5):
        step(t)
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for k in range(Kd):
            step(5 + k)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b) / Kd
    tot = stats[5:].sum(0).cpu().numpy()
    return {"datasets": D, "scenarios_per_dataset": per, "ms_per_step": ms, "value": S / (ms * 1e-3), "unit": UNIT, "steps": Kd,
            "setup_s": setup_s, "gain": "tz_gain_synthesis (one launch, a robust LQR gain per data set)",
            "gain_iterations_max": int(ens.theta_info["iterations"].max()), "rho_max": float(ens.theta_info["rho"].max()), "api": "TZDDPCEnsemble -> tz_closed_loop_step_set (one launch per step, one program per data set)",
            "status_ok_frac": float(1.0 - (tot[3] + tot[4] + tot[6]) / max(tot[7], 1.0)), "iters_mean": float(tot[5] / max(tot[7], 1.0))}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import tzddpc_b200 as tz
    from tzddpc_b200 import _abi, configs, ops, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = configs.CONFIGS[args.workload]()
    S, K_steps, W_steps = args.scenarios, args.steps, args.warmup
    n, m, N = cfg.n, cfg.m, cfg.horizon

    # ---- setup (not timed): identify on the GPU, canonicalise, upload the program
    rng = np.random.default_rng(cfg.seed)
    u_data, x_data = configs.generate_dataset(cfg, rng)
    ctl = tz.TZDDPC(tz.Data(u_data, x_data), device=dev)
    ctl.verbose = False
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    ctl.build_zonotopes(zon)
    Kgain = configs.lqr_gain(ctl.Mdata.center[:, :n], ctl.Mdata.center[:, n:])
    ctl.build_zonotopes_theta(zon, K=Kgain)
    ctl.build_problem(N, tz.StageCost(**cfg.cost), tz.BoxConstraint(**cfg.box) if cfg.box else tz.BoxConstraint())
    prog = ctl._program
    g1 = prog.compiled.g1
    nent, nt, nv = n * (1 + g1), (N + 1) * n, prog.compiled.nv
    opts = tz.SolverOptions(warm_start=int(args.warm_start), check_every=args.check_every, eps_abs=args.eps, eps_rel=args.eps,
                            polish=args.polish)
    po = opts.pack()

    f64 = dict(dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(cfg.seed + 7919 * rank)
    cW, GW = torch.tensor(cfg.W[0], **f64), torch.tensor(cfg.W[1], **f64)
    total = W_steps + K_steps
    nring = min(total, 64)          # noise realisations are drawn for 64 steps and cycled (bounded memory for long runs)
    beta = torch.rand((nring, GW.shape[1], S), generator=gen, **f64) * 2 - 1
    if cfg.noise == "vertex":
        beta = torch.sign(beta)
    noise = (cW[None, :, None] + torch.einsum("rg,tgs->trs", GW, beta)).contiguous()      # (total, n, S)
    del beta
    x0 = torch.tensor(cfg.X0[0], **f64)
    x = x0[:, None].repeat(1, S).contiguous()
    xbar = x.clone()
    xrestart = x.clone()         # an infeasible scenario (the reference raises: end of that run) starts a new run from x0
    e = torch.zeros((n, S), **f64)
    At, Bt = torch.tensor(cfg.A, **f64), torch.tensor(cfg.B, **f64)
    ring = 4
    ze1 = torch.empty((ring, nent, S), **f64)
    traj = torch.empty((ring, nt, S), **f64)
    vbuf = torch.empty((ring, nv, S), **f64)
    cost = torch.empty((ring, S), **f64)
    status = torch.zeros((ring, S), dtype=torch.int32, device=dev)
    iters = torch.zeros((ring, S), dtype=torch.int32, device=dev)
    stats = torch.zeros((total, _abi.TZ_NSTATS), **f64)
    warm = torch.zeros((prog.warm_rows, S), **f64) if args.warm_start else None
    h = prog.handle.value

    def step(t):
        r = t % ring
        ops.closed_loop_step(h, x, xbar, e, noise[t % nring], xrestart, At, Bt, status[r], cost[r], vbuf[r],
                             None if args.ablate >= 2 else traj[r], None if args.ablate >= 1 else ze1[r], None,
                             iters[r], warm, stats[t], po)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for t in range(W_steps):
        step(t)
    barrier()
    # The K timed steps are captured in ONE CUDA graph (K kernel nodes, no host work between steps), so that the number
    # does not depend on how fast this process's Python loop launches (8 ranks share the host's cores).  --dump-steps
    # (diagnostics) times eager launches step by step instead.
    use_graph = not args.no_graph and not args.dump_steps
    graph = None
    if use_graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for k in range(K_steps):
                step(W_steps + k)
        barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range((1 if use_graph else K_steps) + 1)]
    t_wall0 = time.perf_counter()
    ev[0].record()
    if use_graph:
        graph.replay()
        ev[1].record()
    else:
        for k in range(K_steps):
            step(W_steps + k)
            ev[k + 1].record()
    if sampler is not None:          # the work is asynchronous: sample the clocks while the GPU goes through it
        while not ev[-1].query():
            sampler.sample()
            time.sleep(0.0005)
    barrier()
    t_wall1 = time.perf_counter()
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [elapsed_ms / K_steps] if use_graph else [ev[k].elapsed_time(ev[k + 1]) for k in range(K_steps)]
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    elapsed_ms = shard.max_over_ranks(elapsed_ms, dev)
    # final statistics: the only collective of the path, one all-reduce after the loop
    tot_stats = shard.reduce_statistics(stats[W_steps:].clone()).sum(0).cpu().numpy()
    cnt = max(tot_stats[7], 1.0)
    if args.dump_steps and rank == 0:
        np.savez(args.dump_steps, kern_ms=np.asarray(kern_ms), stats=stats[W_steps:].cpu().numpy())

    value = world * S * K_steps / (elapsed_ms * 1e-3)
    bstep = algorithmic_bytes_per_scenario_step(n, m, N, g1)
    peak, peak_src = measured_hbm_peak()
    avg_kernel_ms = float(np.mean(kern_ms))
    achieved = bstep * S / (avg_kernel_ms * 1e-3) / 1e9
    traffic = measured_traffic(prog.bucket)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_steps, "warmup": W_steps,
            "ms_per_step": elapsed_ms / K_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, S), "scenarios_per_gpu": S, "parallelism": f"scenario-dp{world}",
                       "l2": f"per-step HBM traffic {bstep * S / 1e6:.0f} MB > 126 MB L2 (inputs larger than L2)",
                       "solver": {"warm_start": {0: "cold", 1: "previous (x, y)", 2: "active-set hint of the previous step (KKT-certified)"}[int(args.warm_start)],
                                  "eps": opts.eps_abs, "polish": True, "certificate": "active-set KKT"},
                       "kernel_bucket": prog.bucket,
                       "episodes": "a scenario whose step is infeasible (the reference raises: end of that run) starts a new "
                                   "run from x0; the 5-dim example reaches that point every ~62 steps",
                       "noise": f"{nring} pre-drawn realisations per scenario, cycled"},
            "gpu_launches": K_steps, "launch_mode": "one CUDA graph of the K steps" if use_graph else "eager",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_scenario_step": bstep,
                         "kernel_ms_avg": avg_kernel_ms, "kernel": "tz::step_kernel_param"},
            "solver_stats": {"iters_mean": float(tot_stats[5] / cnt), "iters_max_last_steps": int(iters.max().item()),
                             "status_ok_frac": float(1.0 - (tot_stats[3] + tot_stats[4] + tot_stats[6]) / cnt),
                             "infeasible": int(tot_stats[3]), "maxiter": int(tot_stats[4]), "nonfinite": int(tot_stats[6]),
                             "mean_norm_x": float(tot_stats[0] / cnt)},
            "clocks": clocks}

    # ---- e2e: host buffers through the C-ABI host entry point, every output of the step back on the host.
    # Headline `e2e`: the tube crosses the bus PACKED (TzSolverOpts.tube_packed: the n_nz entries of Ze[1].Z that are not
    # structurally zero; the binding scatters them into the dense matrix when `.Z.value` is read, as the reference's
    # `.Z.value` evaluates its expression on access).  `e2e_dense_tube`: the same call with the dense n x (1+g1) matrix.
    if not args.no_e2e:
        Ke = max(3, min(args.e2e_steps, K_steps))
        pin = lambda *shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()     # noqa: E731
        hx, hxb, he = pin(n, S), pin(n, S), pin(n, S)
        hnoise = pin(Ke + 2, n, S)
        hnoise.copy_(noise[torch.arange(Ke + 2, device=dev) % nring].cpu())
        nnz = len(prog.tube_pattern)
        hcost, hv, htraj, hze = pin(S), pin(nv, S), pin(nt, S), pin(nent, S)
        hstat = pin(S, dt=torch.int32)
        scratch = torch.zeros(_abi.lib().tz_closed_loop_step_host_scratch_bytes(h, S) // 8 + 8, **f64)
        import ctypes as C
        Ah, Bh = np.ascontiguousarray(cfg.A), np.ascontiguousarray(cfg.B)

        def e2e_leg(packed):
            o_ = ops._opts(tz.SolverOptions(warm_start=int(args.warm_start), check_every=args.check_every, eps_abs=args.eps, eps_rel=args.eps,
                                            polish=args.polish, tube_packed=int(packed)).pack())
            hx.copy_(x0[:, None].cpu().expand(n, S)); hxb.copy_(hx); he.zero_()
            scratch.zero_()

            def host_step(t):
                rc = _abi.lib().tz_closed_loop_step_host(
                    C.c_void_p(h), C.byref(o_), S, C.c_void_p(hx.data_ptr()), C.c_void_p(hxb.data_ptr()), C.c_void_p(he.data_ptr()),
                    C.c_void_p(hnoise[t].data_ptr()), C.c_void_p(Ah.ctypes.data), C.c_void_p(Bh.ctypes.data),
                    C.c_void_p(hcost.data_ptr()), C.c_void_p(hv.data_ptr()), C.c_void_p(htraj.data_ptr()), C.c_void_p(hze.data_ptr()),
                    C.c_void_p(hstat.data_ptr()), C.c_void_p(scratch.data_ptr()), args.e2e_chunks)
                _abi.check(rc, "tz_closed_loop_step_host")

            for t in range(2):
                host_step(t)
            barrier()
            t0 = time.perf_counter()
            for t in range(Ke):
                host_step(2 + t)
            barrier()
            dt = shard.max_over_ranks(time.perf_counter() - t0, dev)
            rows = nnz if packed else nent
            return {"value": world * S * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": int(4 * n * S * 8 + 8 * (n * n + n * m)),
                    "d2h_bytes_per_step": int(S * (8 * (3 * n + 1 + nv + nt + rows) + 4)), "steps": Ke,
                    "ms_per_step": 1e3 * dt / Ke, "api": "tz_closed_loop_step_host (pinned host buffers)",
                    "tube": (f"packed: {nnz} of {nent} entries of Ze[1].Z per scenario (the others are zero for every "
                             f"(xbar0, e0)); pattern from tz_program_tube_pattern") if packed else "dense n x (1+g1)",
                    "solver": "active-set hints carried between calls in the caller's device scratch (warm_start)",
                    "status_ok_frac": float((hstat == 0).double().mean().item())}

        line["e2e"] = e2e_leg(True)
        line["e2e_dense_tube"] = e2e_leg(False)

    # ---- data-set axis (north_star: scenarios = noise realisations x initial states x data sets): D data sets, one
    # program each, the whole batch in one launch per step (tz_closed_loop_step_set); rank 0 at N = 1 only
    if rank == 0 and world == 1 and args.datasets > 0:
        try:
            line["datasets_axis"] = datasets_leg(args, tz, ops, torch, dev, cfg, noise, nring)
        except Exception as exc:      # noqa: BLE001  (secondary leg: never fail the headline line)
            line["datasets_axis"] = {"error": repr(exc)[:300]}

    # ---- batch-1 latency (BASELINE.json metric: "us/step at batch 1"): examples/1.double_integrator_sim.py as shipped, one
    # scenario, the closed loop captured in a CUDA graph (50 fused steps per replay) so that no host work sits between steps
    if rank == 0 and world == 1 and not args.no_batch1:
        try:
            line["batch1"] = batch1_latency(tz, ops, torch, dev)
        except Exception as exc:      # noqa: BLE001  (diagnostic leg: never fail the headline line)
            line["batch1"] = {"error": repr(exc)[:200]}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only): the oracle port, one process per core
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0)) if args.cpu_cores <= 0 else args.cpu_cores
        scen, cpu_steps = 1, args.cpu_steps
        val, secs = cpu_oracle_throughput(args.workload, cores, scen, 2, cpu_steps)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cores * scen} scenarios (one per core) x {cpu_steps} closed-loop steps of the same "
                                          f"workload ({secs:.1f} s), numpy oracle of the reference path",
                                "per_core": val / cores}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="fivedim", choices=["fivedim", "pulley", "double_integrator"])
    ap.add_argument("--scenarios", type=int, default=65536, help="scenarios per GPU")
    ap.add_argument("--warm-start", type=int, default=2, help="0 cold, 1 previous (x, y), 2 active-set hint (default)")
    ap.add_argument("--check-every", type=int, default=8)
    ap.add_argument("--eps", type=float, default=1e-6)
    ap.add_argument("--polish", type=int, default=3, help="augmented-Lagrangian iterations of the certificate / polish")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-chunks", type=int, default=2)
    ap.add_argument("--datasets", type=int, default=64, help="data sets of the data-set-axis leg (0 = skip)")
    ap.add_argument("--cpu-steps", type=int, default=1500)
    ap.add_argument("--cpu-cores", type=int, default=0, help="processes of the CPU baseline (0 = all host cores)")
    ap.add_argument("--dump-steps", default="", help="write per-step kernel ms and solver statistics to this .npz (diagnostics)")
    ap.add_argument("--ablate", type=int, default=0, help="diagnostics only: 1 = do not write Ze[1].Z, 2 = nor the trajectory")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph of the K steps")
    ap.add_argument("--no-batch1", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
