"""tz_closed_loop_step_host -- the host-buffer entry point bench.py times for `e2e` -- against the device-buffer step
(tz_closed_loop_step) on the same inputs: every output bit-equal, for the dense and the packed tube, cold and with the
active-set hints carried in the caller's scratch, for several chunk counts and a ragged batch."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,S", [("fivedim", 1000), ("pulley", 70), ("double_integrator", 16)])
def test_host_call_equals_device_call(cuda_lib, name, S):
    import tzddpc_b200 as tz
    from tzddpc_b200 import _abi, ops
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    prog = t._program
    n, m, N, g1, nv = cfg.n, cfg.m, cfg.horizon, prog.compiled.g1, prog.compiled.nv
    nent, nt, nnz = n * (1 + g1), (N + 1) * n, len(prog.tube_pattern)
    steps = 5
    rng = np.random.default_rng(4)
    noise = np.ascontiguousarray(np.transpose(common.noise_for(cfg, steps, S, rng), (0, 2, 1)))      # (steps, n, S)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64)[:, None], (1, S))
    dev = t.device
    h = prog.handle.value
    lib = _abi.lib()
    A, B = np.ascontiguousarray(cfg.A, dtype=np.float64), np.ascontiguousarray(np.asarray(cfg.B, dtype=np.float64).reshape(n, m))
    f64 = dict(dtype=torch.float64, device=dev)
    for packed, warm_start, chunks in [(0, 0, 1), (1, 0, 3), (0, 2, 2), (1, 2, 4)]:
        opts = tz.SolverOptions(warm_start=warm_start, tube_packed=packed)
        rows = nnz if packed else nent
        # device path
        dx, dxb, de = torch.tensor(x0, **f64), torch.tensor(x0, **f64), torch.zeros((n, S), **f64)
        warm = torch.zeros((prog.warm_rows, S), **f64) if warm_start else None
        At, Bt = torch.tensor(A, **f64), torch.tensor(B, **f64)
        # host path
        hx, hxb, he = x0.copy(), x0.copy(), np.zeros((n, S))
        hcost, hv, htraj, hze, hst = np.zeros(S), np.zeros((nv, S)), np.zeros((nt, S)), np.zeros((rows, S)), np.zeros(S, dtype=np.int32)
        scratch = torch.zeros(lib.tz_closed_loop_step_host_scratch_bytes(h, S) // 8 + 8, **f64)
        o_ = ops._opts(opts.pack())
        p = lambda a: C.c_void_p(a.ctypes.data)      # noqa: E731
        for k in range(steps):
            status = torch.zeros(S, dtype=torch.int32, device=dev)
            cost, v, traj, ze = torch.empty(S, **f64), torch.empty((nv, S), **f64), torch.empty((nt, S), **f64), torch.empty((rows, S), **f64)
            ops.closed_loop_step(h, dx, dxb, de, torch.tensor(noise[k], **f64), None, At, Bt, status, cost, v, traj, ze, None, None,
                                 warm, None, opts.pack())
            rc = lib.tz_closed_loop_step_host(C.c_void_p(h), C.byref(o_), S, p(hx), p(hxb), p(he), p(noise[k]), p(A), p(B), p(hcost),
                                              p(hv), p(htraj), p(hze), p(hst), C.c_void_p(scratch.data_ptr()), chunks)
            _abi.check(rc, "tz_closed_loop_step_host")
            tag = f"{name} packed={packed} warm={warm_start} chunks={chunks} step {k}"
            np.testing.assert_array_equal(hst, status.cpu().numpy(), err_msg=tag)
            for a, b, nm in ((hx, dx, "x"), (hxb, dxb, "xbar"), (he, de, "e"), (hcost, cost, "cost"), (hv, v, "v"), (htraj, traj, "traj"),
                             (hze, ze, "tube")):
                np.testing.assert_array_equal(a, b.cpu().numpy(), err_msg=f"{tag}: {nm}")
        assert (hst == 0).mean() > 0.5


@pytest.mark.parametrize("name,S,chunks", [("fivedim", 1000, 3), ("pulley", 70, 1), ("fivedim", 4096, 4)])
def test_resident_host_run_equals_device_loop(cuda_lib, name, S, chunks):
    """tz_closed_loop_run_host: the state stays in the caller's device scratch between calls; per call the noise goes up and
    x+, u, cost, status and the packed tube come down.  Bit-equal to the device loop (with restart from x_restart)."""
    import tzddpc_b200 as tz
    from tzddpc_b200 import _abi, ops
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    prog = t._program
    n, m, N, g1, nv = cfg.n, cfg.m, cfg.horizon, prog.compiled.g1, prog.compiled.nv
    nt, nnz = (N + 1) * n, len(prog.tube_pattern)
    steps = 80 if name == "fivedim" and S <= 1000 else 8            # (80 steps: through the first infeasibility wave of the 5-dim example)
    rng = np.random.default_rng(9)
    noise = np.ascontiguousarray(np.transpose(common.noise_for(cfg, steps, S, rng), (0, 2, 1)))
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64)[:, None], (1, S))
    dev = t.device
    h = prog.handle.value
    lib = _abi.lib()
    A, B = np.ascontiguousarray(cfg.A, dtype=np.float64), np.ascontiguousarray(np.asarray(cfg.B, dtype=np.float64).reshape(n, m))
    f64 = dict(dtype=torch.float64, device=dev)
    opts = tz.SolverOptions(warm_start=2, tube_packed=1)
    dx, dxb, de, dxr = torch.tensor(x0, **f64), torch.tensor(x0, **f64), torch.zeros((n, S), **f64), torch.tensor(x0, **f64)
    warm = torch.zeros((prog.warm_rows, S), **f64)
    At, Bt = torch.tensor(A, **f64), torch.tensor(B, **f64)
    hx, hxb, he = x0.copy(), x0.copy(), np.zeros((n, S))
    hcost, hze, hu, hst = np.zeros(S), np.zeros((nnz, S)), np.zeros((m, S)), np.zeros(S, dtype=np.int32)
    scratch = torch.zeros(lib.tz_closed_loop_step_host_scratch_bytes(h, S) // 8 + 8, **f64)
    o_ = ops._opts(opts.pack())
    p = lambda a: C.c_void_p(a.ctypes.data)      # noqa: E731
    restarted = 0
    for k in range(steps):
        status = torch.zeros(S, dtype=torch.int32, device=dev)
        cost, ze, uu = torch.empty(S, **f64), torch.empty((nnz, S), **f64), torch.empty((m, S), **f64)
        ops.closed_loop_step(h, dx, dxb, de, torch.tensor(noise[k], **f64), dxr, At, Bt, status, cost, None, None, ze, uu, None,
                             warm, None, opts.pack())
        last = k == steps - 1
        rc = lib.tz_closed_loop_run_host(C.c_void_p(h), C.byref(o_), S, (1 if k == 0 else 0) | (2 if last else 0), p(hx), p(hxb), p(he),
                                         p(x0), p(noise[k]), p(A), p(B), p(hcost), None, None, p(hze), p(hu), p(hst),
                                         C.c_void_p(scratch.data_ptr()), chunks)
        _abi.check(rc, "tz_closed_loop_run_host")
        st = status.cpu().numpy()
        restarted += int((st == 2).sum())
        tag = f"{name} step {k}"
        np.testing.assert_array_equal(hst, st, err_msg=tag)
        good = st == 0
        np.testing.assert_array_equal(hx, dx.cpu().numpy(), err_msg=tag + ": x")
        np.testing.assert_array_equal(hcost, cost.cpu().numpy(), err_msg=tag + ": cost")
        np.testing.assert_array_equal(hze[:, good], ze.cpu().numpy()[:, good], err_msg=tag + ": tube")
        np.testing.assert_array_equal(hu[:, good], uu.cpu().numpy()[:, good], err_msg=tag + ": u")
    np.testing.assert_array_equal(hxb, dxb.cpu().numpy())
    np.testing.assert_array_equal(he, de.cpu().numpy())
    if steps >= 80:
        assert restarted > 0, "expected infeasible steps (restart from x_restart) in the window"
