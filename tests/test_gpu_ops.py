"""GPU parity of the stand-alone kernels (through torch.ops.tzddpc.* -> C ABI) against the oracle:
interval hull, MatrixZonotope x Zonotope (+W), Girard reduction, identification, explicit-instance ADMM;
edge cases: no generators, zero generators, ties, no-op reductions, ragged batch sizes."""
import os

import numpy as np
import pytest

import oracle
from oracle.zono import girard_reduce_generators, matzono_times_Z
from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def T(cuda_lib):
    import torch
    assert torch.cuda.is_available()
    from tzddpc_b200 import ops  # noqa: F401  (registers torch.ops.tzddpc)
    return torch


def _gpu(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("S,n,g", [(1, 1, 0), (3, 2, 1), (33, 5, 113), (7, 4, 619), (130, 5, 31), (2, 8, 2700)])
def test_interval_hull(T, S, n, g):
    rng = np.random.default_rng(S * 1000 + g)
    Z = rng.normal(size=(S, n, 1 + g)) * rng.choice([0.0, 1e-3, 1.0, 1e3], size=(S, 1, 1 + g))
    lo, hi = T.ops.tzddpc.interval_hull(_gpu(T, Z))
    for s in range(S):
        iv = oracle.Zonotope(Z[s, :, 0], Z[s, :, 1:]).interval
        np.testing.assert_allclose(lo[s].cpu().numpy(), iv.left_limit, rtol=common.GEN_RTOL, atol=1e-300)
        np.testing.assert_allclose(hi[s].cpu().numpy(), iv.right_limit, rtol=common.GEN_RTOL, atol=1e-300)


@pytest.mark.parametrize("S,n,p,N,g,gW,per", [(5, 2, 2, 4, 1, 0, False), (9, 5, 5, 25, 1, 0, False), (9, 5, 6, 30, 1, 1, False),
                                             (4, 4, 4, 16, 75, 1, False), (6, 3, 4, 5, 10, 2, True), (1, 5, 5, 0, 40, 3, False),
                                             (3, 5, 5, 25, 300, 1, False)])
def test_reach_step(T, S, n, p, N, g, gW, per):
    rng = np.random.default_rng(N * 31 + g)
    C = rng.normal(size=(S, n, p) if per else (n, p))
    Gm = rng.normal(size=(S, N, n, p) if per else (N, n, p))
    Z = rng.normal(size=(S, p, 1 + g))
    W = rng.normal(size=(n, 1 + gW)) if gW else None
    out = T.ops.tzddpc.reach_step(_gpu(T, C), _gpu(T, Gm), _gpu(T, Z), None if W is None else _gpu(T, W)).cpu().numpy()
    assert out.shape == (S, n, (N + 1) * (1 + g) + gW)
    for s in range(S):
        ref = matzono_times_Z(C[s] if per else C, Gm[s] if per else Gm, Z[s])
        if W is not None:       # Minkowski sum with W (tzddpc/tzddpc.py:176)
            ref = np.hstack([ref, W[:, 1:]])
            ref[:, 0] += W[:, 0]
        np.testing.assert_allclose(out[s], ref, rtol=common.GEN_RTOL, atol=1e-13)


@pytest.mark.parametrize("metric", ["l1-linf", "l1", "l2"])
@pytest.mark.parametrize("n,g,order", [(2, 24, 3), (5, 113, 1), (5, 113, 2.5), (5, 1413, 10), (4, 10, 5), (3, 40, 1.5), (5, 2763, 20)])
def test_girard_reduce(T, metric, n, g, order):
    rng = np.random.default_rng(g + int(order * 10))
    S = 6
    Z = rng.normal(size=(S, n, 1 + g)) * rng.uniform(0.01, 1.0, size=(S, 1, 1 + g))
    Z[1, :, 3:3 + min(5, g - 3)] = 0.0                 # all-zero generators are dropped first
    Z[2, :, 1 + g // 2:] = Z[2, :, 1:1 + (g - g // 2)]  # exact ties: lowest index is reduced first
    Z[3, :, 1:] = np.round(Z[3, :, 1:], 1)             # many equal metrics
    conv = oracle.Conventions(girard_metric=metric)
    cap = max(g, int(np.ceil(n * order)) + n)
    out, gout = T.ops.tzddpc.girard_reduce(_gpu(T, Z), float(order), {"l1-linf": 0, "l1": 1, "l2": 2}[metric], cap)
    out, gout = out.cpu().numpy(), gout.cpu().numpy()
    for s in range(S):
        ref = girard_reduce_generators(Z[s, :, 1:], order, conv)
        assert gout[s] == ref.shape[1], (s, gout[s], ref.shape)
        np.testing.assert_array_equal(out[s, :, 0], Z[s, :, 0])
        kept = ref.shape[1] - (n if ref.shape[1] < np.count_nonzero(np.any(Z[s, :, 1:] != 0, axis=0)) else 0)
        np.testing.assert_array_equal(out[s, :, 1:1 + kept], ref[:, :kept])           # kept generators: bit-exact copies
        np.testing.assert_allclose(out[s, :, 1 + kept:1 + gout[s]], ref[:, kept:], rtol=common.GEN_RTOL, atol=1e-300)
        assert not np.any(out[s, :, 1 + gout[s]:])


@pytest.mark.parametrize("g", [40, 400])          # warp-per-zonotope kernel / CTA kernel
def test_girard_reduce_row_carried_by_the_kept_generators(T, g):
    """Row 0 is non-zero only in the three generators of largest metric (kept at order 2, n = 3): its box entry is exactly 0
    (the warp kernel computes the box as total minus kept row sums: they must cancel exactly here)."""
    rng = np.random.default_rng(g)
    n, order = 3, 2.0
    Z = np.zeros((2, n, 1 + g))
    Z[:, 1:, 1:] = 0.01 * rng.normal(size=(2, n - 1, g))
    big = [5, g // 2, g - 2]
    for j in big:
        Z[:, :, 1 + j] = [[7.0 + j, 1.0 + 0.1 * j, -2.0 - 0.01 * j]] * 2
    Z[1, 1, 1:] += 0.3                                  # second zonotope: another row with heavy cancellation (not exact)
    out, gout = T.ops.tzddpc.girard_reduce(_gpu(T, Z), order, 0, 2 * n)
    out, gout = out.cpu().numpy(), gout.cpu().numpy()
    for s in range(2):
        ref = girard_reduce_generators(Z[s, :, 1:], order)
        assert gout[s] == ref.shape[1] == 2 * n
        np.testing.assert_array_equal(out[s, :, 1:1 + n], ref[:, :n])
        np.testing.assert_allclose(out[s, :, 1 + n:], ref[:, n:], rtol=common.GEN_RTOL, atol=1e-300)
        assert out[s, 0, 1 + n] == 0.0


def test_girard_reduce_matrix_zonotope_order_one(T):
    """MatrixZonotope.reduce(1) of tzddpc/tzddpc.py:126-128 on the vectorised generators (dimension n(n+m) = 30)."""
    cfg = configs.fivedim()
    u, x = common.dataset(cfg)
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    o.build_zonotopes(common.oracle_zonotopes(cfg))
    Gv = np.stack([G.flatten() for G in o.Mdata.generators], axis=1)            # 30 x 399
    Z = np.hstack([o.Mdata.center.reshape(-1, 1), Gv])[None]
    out, gout = T.ops.tzddpc.girard_reduce(_gpu(T, Z), 1.0, 0, Gv.shape[1])
    ref = girard_reduce_generators(Gv, 1)
    assert int(gout[0]) == ref.shape[1] == 30
    np.testing.assert_allclose(out[0, :, 1:31].cpu().numpy(), ref, rtol=common.GEN_RTOL, atol=1e-300)


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_identify_batched_datasets(T, name):
    """Per-scenario data sets (the 'datasets' scenario axis of BASELINE.json): S different trajectories at once."""
    cfg = configs.CONFIGS[name]()
    S = 5
    rng = np.random.default_rng(99)
    U, X = configs.generate_dataset(cfg, rng, num_trajectories=S)
    U, X = U.reshape(S, cfg.T, cfg.m), X.reshape(S, cfg.T, cfg.n)
    Ks = rng.normal(size=(S, cfg.m, cfg.n)) * 0.3
    WZ = np.hstack([cfg.W[0][:, None], cfg.W[1]])
    AB, dAB, dK, Pinv, status = T.ops.tzddpc.identify(_gpu(T, X), _gpu(T, U), _gpu(T, WZ), _gpu(T, Ks), True)
    assert (status.cpu().numpy() == 0).all()
    for s in range(S):
        o = oracle.OracleTZDDPC(oracle.Data(U[s], X[s]))
        z = common.oracle_zonotopes(cfg)
        o.build_zonotopes_theta(z, Ks[s])
        np.testing.assert_allclose(AB[s].cpu().numpy(), o.Mdata.center, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dAB[s].cpu().numpy(), np.abs(o.Mdelta.generators).sum(0), rtol=common.GEN_RTOL, atol=1e-13)
        np.testing.assert_allclose(dK[s].cpu().numpy(), np.abs(o.MdataK.generators).sum(0), rtol=common.GEN_RTOL, atol=1e-13)
        D = np.vstack([X[s, :-1].T, U[s, :-1].T])
        Pref = np.linalg.pinv(D)
        np.testing.assert_allclose(Pinv[s].cpu().numpy(), Pref, rtol=common.GEN_RTOL, atol=common.GEN_RTOL * np.abs(Pref).max())


@pytest.mark.parametrize("scale", [1e-3, 1e-5])
def test_identify_ill_conditioned_data(T, scale):
    """cond([X0; U0]) up to ~1e6 (one state channel recorded in other units): the Householder route keeps the error at
    cond * eps where normal equations lose cond^2 * eps (round 1: 1e-8 .. 1e-7 already on the well-conditioned examples)."""
    cfg = configs.fivedim()
    rng = np.random.default_rng(5)
    U, X = configs.generate_dataset(cfg, rng)
    X = X.copy()
    X[:, 0] *= scale
    D = np.vstack([X[:-1].T, U[:-1].T])
    cond = np.linalg.cond(D)
    assert cond > 0.05 / scale
    WZ = np.hstack([cfg.W[0][:, None], cfg.W[1]])
    Ks = rng.normal(size=(1, cfg.m, cfg.n)) * 0.3
    AB, dAB, dK, Pinv, status = T.ops.tzddpc.identify(_gpu(T, X[None]), _gpu(T, U[None]), _gpu(T, WZ), _gpu(T, Ks), True)
    assert int(status[0].item()) == 0
    Pref = np.linalg.pinv(D)
    tol = max(common.GEN_RTOL, 50 * cond * np.finfo(np.float64).eps)
    # column-wise: the column of the rescaled channel is 1/scale times larger than the others
    err = np.abs(Pinv[0].cpu().numpy() - Pref).max(axis=0) / np.abs(Pref).max(axis=0)
    assert err.max() <= tol, (err, tol, cond)
    ABref = (X[1:].T - cfg.W[0][:, None]) @ Pref
    errAB = np.abs(AB[0].cpu().numpy() - ABref).max(axis=0) / np.abs(ABref).max(axis=0)
    assert errAB.max() <= tol, (errAB, tol, cond)
    gw = np.abs(cfg.W[1]).sum(axis=1)
    np.testing.assert_allclose(dAB[0].cpu().numpy(), np.outer(gw, np.abs(Pref).sum(axis=0)), rtol=tol)


def test_identify_is_bitwise_repeatable(T):
    """The same data set identified 40 times (one launch, and again launch by launch) gives the same bits: every CTA-wide
    sum of tz_identify is combined in a fixed order.  bench.py's N-GPU == 1-GPU check builds one controller per rank."""
    cfg = configs.fivedim()
    rng = np.random.default_rng(cfg.seed)
    U, X = configs.generate_dataset(cfg, rng)
    WZ = np.hstack([cfg.W[0][:, None], cfg.W[1]])
    Ks = rng.normal(size=(1, cfg.m, cfg.n)) * 0.3
    R = 40
    outs = T.ops.tzddpc.identify(_gpu(T, np.repeat(X[None], R, 0)), _gpu(T, np.repeat(U[None], R, 0)), _gpu(T, WZ),
                                 _gpu(T, np.repeat(Ks, R, 0)), True)
    for o in outs[:4]:
        assert bool((o == o[:1]).all()), "identify differs between CTAs of one launch"
    for _ in range(5):
        again = T.ops.tzddpc.identify(_gpu(T, X[None]), _gpu(T, U[None]), _gpu(T, WZ), _gpu(T, Ks), True)
        for a, o in zip(again[:4], outs[:4]):
            assert bool((a[0] == o[0]).all()), "identify differs between launches"


def test_identify_flags_rank_deficient_data(T):
    cfg = configs.double_integrator()
    X = np.zeros((2, 20, 2)); U = np.zeros((2, 20, 1))
    rng = np.random.default_rng(0)
    U[1], X[1] = configs.generate_dataset(cfg, rng)[0][:20], configs.generate_dataset(cfg, rng)[1][:20]
    WZ = np.hstack([cfg.W[0][:, None], cfg.W[1]])
    *_, status = T.ops.tzddpc.identify(_gpu(T, X), _gpu(T, U), _gpu(T, WZ), None, False)
    assert status.cpu().numpy().tolist() == [3, 0]


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_solve_matches_committed_golden(T, name):
    """CUDA `solve` against tests/golden/oracle_<name>.npz (oracle outputs committed by make_golden.py)."""
    cfg = configs.CONFIGS[name]()
    fx = np.load(os.path.join(GOLD, f"oracle_{name}.npz"))
    t = common.make_product(cfg, fx["u_data"], fx["x_data"], fx["K"])
    np.testing.assert_allclose(t.Mdata.center, fx["AB"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(t.Mdelta.generators, fx["GD"], rtol=common.GEN_RTOL, atol=1e-13)
    np.testing.assert_allclose(t.MdataK.generators, fx["GK"], rtol=common.GEN_RTOL, atol=1e-13)
    cost, v, xbar, tube, status = t.solve(fx["xbar0"], fx["e0"])
    Z = tube.Z.value
    wmax = t._program.compiled.wmax
    assert np.array_equal(status == 2, fx["status"] == 2)
    ok = fx["status"] == 0
    assert (status[ok] == 0).all() and np.isinf(cost[~ok]).all()
    assert common.cost_close(cost[ok], fx["cost"][ok], wmax).all()
    np.testing.assert_allclose(v[ok, 0], fx["v"][ok, 0], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(xbar[ok, :2], fx["xbar"][ok, :2], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(Z[ok], fx["ze1"][ok], rtol=1e-6, atol=1e-6)


def test_qp_solve_explicit_instances(T):
    """tz_qp_solve: the ADMM stage alone on explicit (q, l, u) batches, against the oracle's interior-point solver."""
    cfg = configs.double_integrator()
    fx = np.load(os.path.join(GOLD, "oracle_double_integrator.npz"))
    t = common.make_product(cfg, fx["u_data"], fx["x_data"], fx["K"])
    prog = t._program.compiled
    rng = np.random.default_rng(5)
    S = 40
    idx = np.flatnonzero(fx["status"] == 0)[:S]
    S = len(idx)
    q = np.zeros((prog.nz, S)); l = np.zeros((prog.nc, S)); u = np.zeros((prog.nc, S))
    for k, i in enumerate(idx):
        p = np.r_[fx["xbar0"][i], fx["e0"][i]]
        w = np.r_[1.0, p, np.abs(prog.Bt @ p + prog.gam)]
        r = prog.R @ w
        q[:, k] = prog.q0 + prog.Qp @ p + 0.01 * rng.normal(size=prog.nz)
        l[:, k], u[:, k] = prog.l0 + r, prog.u0 + r
    z, y, status, iters = T.ops.tzddpc.qp_solve(t._program.handle.value, _gpu(T, q), _gpu(T, l), _gpu(T, u), t.solver_options.pack())
    z, status = z.cpu().numpy(), status.cpu().numpy()
    assert (status == 0).all()
    for k in range(S):
        G = np.vstack([prog.A[np.isfinite(u[:, k])], -prog.A[np.isfinite(l[:, k])]])
        h = np.r_[u[np.isfinite(u[:, k]), k], -l[np.isfinite(l[:, k]), k]]
        zz, _, info = oracle.solve_qp_ipm(prog.P, q[:, k], G, h)
        f = lambda a: 0.5 * a @ prog.P @ a + q[:, k] @ a          # noqa: E731
        assert abs(f(z[:, k]) - f(zz)) <= 1e-6 * max(1.0, abs(f(zz)))
        assert np.all(G @ z[:, k] <= h + 1e-6 * np.maximum(1.0, np.abs(h)))
