"""Out-of-bounds writes of the step kernels, checked with guard bands (compute-sanitizer is closed on this GPU pool, see
DESIGN.md section 9): every output and in/out buffer of tz_closed_loop_step / tz_solve is a window inside a larger
allocation pre-filled with a sentinel bit pattern, and after every call the bytes on both sides of every window must be
untouched.  Covers fast_step_kernel + step_kernel in list mode (dense and packed tube, partial tiles, deferred tiles,
restarts), step_kernel cold / warm, the larger register buckets and the generic large-program kernel (tz_big.cu)."""
import numpy as np
import pytest
import torch

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu

PAD = 512                       # elements on either side of a window
SENT64 = 0x7FF8DEADBEEF1234     # a NaN payload no kernel writes
SENT32 = 0x5EAD1234


class Guarded:
    """Tensors carved out of sentinel-filled allocations; check() asserts that no byte outside the windows changed."""

    def __init__(self, dev):
        self.dev, self.items = dev, []

    def make(self, shape, dtype=torch.float64, init=None):
        n = int(np.prod(shape))
        if dtype == torch.float64:
            raw = torch.full((n + 2 * PAD,), SENT64, dtype=torch.int64, device=self.dev)
            win = raw[PAD:PAD + n].view(torch.float64).view(shape)
        else:
            raw = torch.full((n + 2 * PAD,), SENT32, dtype=torch.int32, device=self.dev)
            win = raw[PAD:PAD + n].view(shape)
        if init is not None:
            win.copy_(torch.as_tensor(init, dtype=dtype, device=self.dev).reshape(shape))
        else:
            win.zero_()
        self.items.append((raw, n, SENT64 if dtype == torch.float64 else SENT32))
        return win

    def check(self, tag):
        for k, (raw, n, sent) in enumerate(self.items):
            lo, hi = raw[:PAD], raw[PAD + n:]
            assert bool((lo == sent).all()) and bool((hi == sent).all()), f"{tag}: buffer {k} written outside its window"


@pytest.mark.parametrize("name,S,steps", [("fivedim", 70, 70), ("fivedim", 333, 6), ("fivedim", 4099, 6), ("pulley", 33, 5),
                                          ("pulley", 1000, 5), ("double_integrator", 300, 5)])
def test_step_kernels_stay_inside_their_buffers(cuda_lib, name, S, steps):
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
    rng = np.random.default_rng(1)
    noise = np.ascontiguousarray(np.transpose(common.noise_for(cfg, steps, S, rng), (0, 2, 1)))
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64)[:, None], (1, S))
    At = torch.tensor(np.ascontiguousarray(cfg.A, dtype=np.float64), **f64)
    Bt = torch.tensor(np.ascontiguousarray(np.asarray(cfg.B, dtype=np.float64).reshape(n, m)), **f64)
    variants = [tz.SolverOptions(warm_start=2, hot_path=1), tz.SolverOptions(warm_start=2, hot_path=1, tube_packed=1),
                tz.SolverOptions(warm_start=2, hot_path=0), tz.SolverOptions(warm_start=1), tz.SolverOptions(tube_packed=1)]
    for vi, opts in enumerate(variants):
        g = Guarded(dev)
        rows = nnz if opts.tube_packed else nent
        dx, dxb, de, dxr = g.make((n, S), init=x0), g.make((n, S), init=x0), g.make((n, S)), g.make((n, S), init=x0)
        status, iters = g.make((S,), torch.int32), g.make((S,), torch.int32)
        cost, v, traj, ze, uu = g.make((S,)), g.make((nv, S)), g.make((nt, S)), g.make((rows, S)), g.make((m, S))
        warm = g.make((prog.warm_rows, S)) if opts.warm_start else None
        stats = g.make((8,))
        k_max = steps if (opts.warm_start == 2 and opts.hot_path) else min(steps, 4)
        for k in range(k_max):
            ops.closed_loop_step(prog.handle.value, dx, dxb, de, torch.tensor(noise[k], **f64), dxr, At, Bt, status, cost, v, traj, ze, uu,
                                 iters, warm, stats, opts.pack())
            torch.cuda.synchronize()
            g.check(f"{name} S={S} variant {vi} step {k}")
        assert bool(torch.isfinite(dx).all())


@pytest.mark.parametrize("horizon,k0,S", [(3, None, 37), (4, 2, 5), (5, None, 9), (6, 1, 7), (10, 1, 3)])
def test_larger_programs_stay_inside_their_buffers(cuda_lib, horizon, k0, S):
    """Buckets B1-B3 and the generic large-program kernel (tz_big.cu), closed-loop form with every optional output."""
    import tzddpc_b200 as tz
    from tzddpc_b200 import ops
    cfg = configs.sweep()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
    t = common.make_product(cfg, u, x, K, horizon=horizon, k0=k0)
    prog = t._program
    n, m, g1, nv = cfg.n, cfg.m, prog.compiled.g1, prog.compiled.nv
    nent, nt, nnz = n * (1 + g1), (horizon + 1) * n, len(prog.tube_pattern)
    dev = t.device
    f64 = dict(dtype=torch.float64, device=dev)
    rng = np.random.default_rng(2)
    noise = np.ascontiguousarray(np.transpose(common.noise_for(cfg, 3, S, rng), (0, 2, 1)))
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64)[:, None], (1, S))
    At = torch.tensor(np.ascontiguousarray(cfg.A, dtype=np.float64), **f64)
    Bt = torch.tensor(np.ascontiguousarray(np.asarray(cfg.B, dtype=np.float64).reshape(n, m)), **f64)
    for packed in (0, 1):
        opts = tz.SolverOptions(tube_packed=packed)
        g = Guarded(dev)
        rows = nnz if packed else nent
        dx, dxb, de, dxr = g.make((n, S), init=x0), g.make((n, S), init=x0), g.make((n, S)), g.make((n, S), init=x0)
        status, iters = g.make((S,), torch.int32), g.make((S,), torch.int32)
        cost, v, traj, ze, uu = g.make((S,)), g.make((nv, S)), g.make((nt, S)), g.make((rows, S)), g.make((m, S))
        stats = g.make((8,))
        for k in range(3):
            ops.closed_loop_step(prog.handle.value, dx, dxb, de, torch.tensor(noise[k], **f64), dxr, At, Bt, status, cost, v, traj, ze, uu,
                                 iters, None, stats, opts.pack())
            torch.cuda.synchronize()
            g.check(f"horizon {horizon} k0 {k0} packed {packed} step {k}")
        assert bool(torch.isfinite(dx).all())
