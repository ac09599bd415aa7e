"""tz_closed_loop_run -- K closed-loop steps in one launch (small batches: the loop of examples/2.pulley_sim.py:81-96 without
a launch per step) -- against K calls of tz_closed_loop_step on the same inputs: every output of every step bit-equal, for
the dense and the packed tube, cold and with active-set hints, odd and even batches, with restarts."""
import numpy as np
import pytest
import torch

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,S,steps", [("double_integrator", 1, 12), ("pulley", 5, 20), ("pulley", 40, 9), ("fivedim", 16, 70),
                                          ("fivedim", 333, 12), ("double_integrator", 1024, 6)])
def test_fused_run_equals_the_step_loop(cuda_lib, name, S, steps):
    import tzddpc_b200 as tz
    from tzddpc_b200 import ops
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    prog = t._program
    n, m, N, g1, nv = cfg.n, cfg.m, cfg.horizon, prog.compiled.g1, prog.compiled.nv
    nent, nt, nnz = n * (1 + g1), (N + 1) * n, len(prog.tube_pattern)
    dev = t.device
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    rng = np.random.default_rng(6)
    noise = torch.tensor(np.ascontiguousarray(np.transpose(common.noise_for(cfg, steps, S, rng), (0, 2, 1))), **f64)     # (steps, n, S)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64)[:, None], (1, S))
    At = torch.tensor(np.ascontiguousarray(cfg.A, dtype=np.float64), **f64)
    Bt = torch.tensor(np.ascontiguousarray(np.asarray(cfg.B, dtype=np.float64).reshape(n, m)), **f64)
    h = prog.handle.value
    restarted = 0
    for opts in (tz.SolverOptions(warm_start=2), tz.SolverOptions(warm_start=2, tube_packed=1), tz.SolverOptions(), tz.SolverOptions(warm_start=1)):
        rows = nnz if opts.tube_packed else nent
        # ---- reference: one launch per step
        x1, xb1, e1, xr = (torch.tensor(x0, **f64), torch.tensor(x0, **f64), torch.zeros((n, S), **f64), torch.tensor(x0, **f64))
        st1, it1 = torch.zeros((steps, S), **i32), torch.zeros((steps, S), **i32)
        c1, v1, tr1 = torch.empty((steps, S), **f64), torch.empty((steps, nv, S), **f64), torch.empty((steps, nt, S), **f64)
        z1, u1, xh1 = torch.empty((steps, rows, S), **f64), torch.empty((steps, m, S), **f64), torch.empty((steps, n, S), **f64)
        xbh1, eh1 = torch.empty((steps, n, S), **f64), torch.empty((steps, n, S), **f64)
        s1 = torch.zeros((steps, 8), **f64)
        w1 = torch.zeros((prog.warm_rows, S), **f64) if opts.warm_start else None
        for k in range(steps):
            ops.closed_loop_step(h, x1, xb1, e1, noise[k], xr, At, Bt, st1[k], c1[k], v1[k], tr1[k], z1[k], u1[k], it1[k], w1, s1[k], opts.pack())
            xh1[k].copy_(x1); xbh1[k].copy_(xb1); eh1[k].copy_(e1)
        # ---- fused
        x2, xb2, e2 = torch.tensor(x0, **f64), torch.tensor(x0, **f64), torch.zeros((n, S), **f64)
        st2, it2 = torch.zeros((steps, S), **i32), torch.zeros((steps, S), **i32)
        c2, v2, tr2 = torch.empty((steps, S), **f64), torch.empty((steps, nv, S), **f64), torch.empty((steps, nt, S), **f64)
        z2, u2, xh2 = torch.empty((steps, rows, S), **f64), torch.empty((steps, m, S), **f64), torch.empty((steps, n, S), **f64)
        s2 = torch.zeros((steps, 8), **f64)
        w2 = torch.zeros((prog.warm_rows, S), **f64) if opts.warm_start else None
        xbh2, eh2 = torch.empty((steps, n, S), **f64), torch.empty((steps, n, S), **f64)
        ops.closed_loop_run(h, steps, x2, xb2, e2, noise, xr, At, Bt, st2, c2, v2, tr2, z2, u2, xh2, xbh2, eh2, it2, w2, s2, opts.pack())
        torch.cuda.synchronize()
        tag = f"{name} S={S} warm={opts.warm_start} packed={opts.tube_packed}"
        for a, b, nm in ((st1, st2, "status"), (it1, it2, "iters"), (xh1, xh2, "x history"), (xbh1, xbh2, "xbar history"), (eh1, eh2, "e history"), (x1, x2, "x"), (xb1, xb2, "xbar"), (e1, e2, "e")):
            assert torch.equal(a, b), f"{tag}: {nm}"
        good = (st1 == 0)
        for a, b, nm in ((c1, c2, "cost"), (v1, v2, "v"), (tr1, tr2, "traj"), (z1, z2, "tube"), (u1, u2, "u")):
            aa, bb = a.cpu().numpy(), b.cpu().numpy()
            g = good.cpu().numpy()
            if aa.ndim == 3:
                g = np.broadcast_to(g[:, None, :], aa.shape)
            np.testing.assert_array_equal(aa[g], bb[g], err_msg=f"{tag}: {nm}")
        np.testing.assert_allclose(s1.cpu().numpy(), s2.cpu().numpy(), rtol=1e-12, atol=1e-12)
        restarted += int((st1 == 2).sum().item())
    if name == "fivedim" and steps >= 70:
        assert restarted > 0, "expected infeasible steps (restarts) in the window"


@pytest.mark.parametrize("name,S", [("pulley", 7), ("fivedim", 64), ("double_integrator", 1)])
def test_simulate_uses_the_fused_run_for_small_batches_and_gets_the_same_answer(cuda_lib, name, S):
    """TZDDPC.simulate below FUSED_RUN_MAX_BATCH scenarios is one tz_closed_loop_run launch; with the switch off it is the step
    loop.  Every returned array is bit-equal."""
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    steps = min(cfg.steps, 30)
    rng = np.random.default_rng(12)
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    assert t._fused_run_ok(S, steps)
    for opts in (tz.SolverOptions(warm_start=2), tz.SolverOptions(tube_packed=1)):
        fused = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, restart=True, options=opts)
        t.FUSED_RUN_MAX_BATCH = 0
        try:
            assert not t._fused_run_ok(S, steps)
            loop = t.simulate(cfg.A, cfg.B, x0, noise, keep_tubes=True, restart=True, options=opts)
        finally:
            del t.FUSED_RUN_MAX_BATCH
        for k in ("x", "xbar", "e", "u", "v", "cost", "status", "iters", "tubes"):
            np.testing.assert_array_equal(fused[k], loop[k], err_msg=f"{name} {k}")
        np.testing.assert_allclose(fused["stats"], loop["stats"], rtol=1e-12, atol=1e-12)
