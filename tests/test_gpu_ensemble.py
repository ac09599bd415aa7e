"""Data-set axis and packed tube (GPU).

* `TZDDPCEnsemble` -- D data sets, one program each, one launch per closed-loop step (tz_closed_loop_step_set) -- must give,
  for every data set's slice of the batch, exactly what that data set's own controller gives (tz_closed_loop_step), and
  the oracle's answer for the oracle built on that data set (the reference: one TZDDPC object per data set,
  tzddpc/tzddpc.py:20-28,67-85,132-241).
* packed tube (`SolverOptions.tube_packed`): scattering the n_nz rows by `tube_pattern` must reproduce the dense Ze[1].Z
  bit for bit (examples/2.pulley_sim.py:96).
"""
import numpy as np
import pytest
import torch

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


def _controllers(cfg, D, with_oracle=False):
    ctls, oracles = [], []
    for d in range(D):
        u, x = common.dataset(cfg, seed=cfg.seed + 101 * d)
        o, K = common.make_oracle(cfg, u, x) if with_oracle else (None, None)
        if K is None:
            import oracle
            oo = oracle.OracleTZDDPC(oracle.Data(u, x))
            oo.build_zonotopes(common.oracle_zonotopes(cfg))
            C = oo.Mdata.center
            K = configs.lqr_gain(C[:, :cfg.n], C[:, cfg.n:])
        ctls.append(common.make_product(cfg, u, x, K))
        oracles.append(o)
    return ctls, oracles


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_ensemble_closed_loop_equals_per_dataset_controllers(cuda_lib, name):
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS[name]()
    D, per, steps = 5, 48, 12
    ctls, _ = _controllers(cfg, D)
    ens = tz.TZDDPCEnsemble(ctls, per)
    assert ens.num_scenarios == D * per and ens.num_datasets == D
    rng = np.random.default_rng(3)
    noise = common.noise_for(cfg, steps, D * per, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (D * per, 1))
    for ws in (0, 2):
        opts = tz.SolverOptions(warm_start=ws)
        r = ens.simulate(cfg.A, cfg.B, x0, noise=noise, keep_tubes=True, options=opts, restart=True)
        for d, c in enumerate(ctls):
            sl = slice(d * per, (d + 1) * per)
            rd = c.simulate(cfg.A, cfg.B, x0[sl], noise=noise[:, sl], keep_tubes=True, options=opts, restart=True)
            for k in ("x", "xbar", "e", "u", "v", "cost", "status", "tubes"):
                np.testing.assert_array_equal(r[k][:, sl], rd[k], err_msg=f"{k}, data set {d}, warm_start {ws}")
        # the data sets differ, so do their answers
        assert not np.array_equal(r["v"][:, :per], r["v"][:, per:2 * per])
        # statistics are accumulated over all programs of the launch
        np.testing.assert_allclose(r["stats"][:, 7], D * per)


def test_ensemble_solve_matches_each_datasets_oracle(cuda_lib):
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS["pulley"]()
    D, per = 3, 16
    ctls, oracles = _controllers(cfg, D, with_oracle=True)
    ens = tz.TZDDPCEnsemble(ctls, per)
    rng = np.random.default_rng(11)
    Xi = oracles[0].zonotopes.X.interval
    xb = Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.2, 0.8, (D * per, cfg.n))
    ee = rng.uniform(-0.2, 0.2, (D * per, cfg.n))
    dev = ens.device
    r = ens.solve_batch(torch.tensor(xb.T.copy(), device=dev), torch.tensor(ee.T.copy(), device=dev))
    cost, v, traj, status = r.cost.cpu().numpy(), r.v.cpu().numpy(), r.xbar.cpu().numpy(), r.status.cpu().numpy()
    wmax = ctls[0]._program.compiled.wmax
    checked = 0
    for s in range(D * per):
        ro = oracles[s // per].solve_status(xb[s], ee[s])
        if ro.status == 2:
            assert status[s] == 2
            continue
        assert status[s] == 0
        assert common.cost_close(cost[s], ro.cost, wmax), (s, cost[s], ro.cost)
        np.testing.assert_allclose(v[0, s], ro.v[0, 0], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(traj[cfg.n:2 * cfg.n, s], ro.xbar[1], rtol=1e-6, atol=1e-6)
        checked += 1
    assert checked >= D * per // 2


def test_program_set_with_more_programs_than_ctas_and_ragged_last(cuda_lib):
    """600 set entries (cycling through 3 programs) > one wave of CTAs: a CTA serves several programs in turn; the last
    entry is a partial tile."""
    import ctypes as C
    import tzddpc_b200 as tz
    from tzddpc_b200 import _abi, ops
    cfg = configs.CONFIGS["fivedim"]()
    ctls, _ = _controllers(cfg, 3)
    nprog, per, last = 600, 16, 6
    progs = [ctls[j % 3]._program for j in range(nprog)]
    begin = np.concatenate([np.arange(nprog) * per, [(nprog - 1) * per + last]]).astype(np.int64)
    pset = _abi.ProgramSet(progs, begin)
    S = int(begin[-1])
    dev = ctls[0].device
    rng = np.random.default_rng(5)
    x0 = np.asarray(cfg.X0[0], dtype=np.float64)
    xb = torch.tensor((x0[:, None] + 0.05 * rng.standard_normal((cfg.n, S))), device=dev)
    ee = torch.tensor(0.02 * rng.standard_normal((cfg.n, S)), device=dev)
    o = tz.SolverOptions()
    cost, v, traj, ze1, status, iters = ops.solve_set(pset.handle.value, ctls[0]._dims, xb, ee, None, True, o.pack())
    for j in range(3):
        idx = np.concatenate([np.arange(begin[k], begin[k + 1]) for k in range(j, nprog, 3)])
        it = torch.as_tensor(idx, device=dev)
        r = ctls[j].solve_batch(xb[:, it].contiguous(), ee[:, it].contiguous())
        assert torch.equal(r.status, status[it])
        assert torch.equal(r.cost, cost[it]) and torch.equal(r.v, v[:, it]) and torch.equal(r.xbar, traj[:, it])
        assert torch.equal(r.tube._ze1, ze1[:, it])
    assert int((status == 0).sum()) > S // 2
    # error paths: misaligned start, structure mismatch is covered by identical programs here; wrong batch size
    with pytest.raises(RuntimeError, match="multiple of 16"):
        _abi.ProgramSet(progs[:2], np.array([0, 8, 24], dtype=np.int64))
    with pytest.raises(RuntimeError, match="created for"):
        ops.solve_set(pset.handle.value, ctls[0]._dims, xb[:, :32].contiguous(), ee[:, :32].contiguous(), None, True, o.pack())


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_packed_tube_scatters_to_the_dense_tube(cuda_lib, name):
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    prog = t._program
    n, g1 = cfg.n, prog.compiled.g1
    pat = prog.tube_pattern
    assert len(pat) < n * (1 + g1) and len(np.unique(pat)) == len(pat) and pat.max() < n * (1 + g1)
    S, steps = 70, 8          # a ragged batch (partial tiles)
    rng = np.random.default_rng(9)
    noise = common.noise_for(cfg, steps, S, rng)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    rd = t.simulate(cfg.A, cfg.B, x0, noise=noise, keep_tubes=True, options=tz.SolverOptions(warm_start=2))
    rp = t.simulate(cfg.A, cfg.B, x0, noise=noise, keep_tubes=True, options=tz.SolverOptions(warm_start=2, tube_packed=1))
    for k in ("x", "xbar", "e", "u", "v", "cost", "status"):
        np.testing.assert_array_equal(rd[k], rp[k])
    ok = rd["status"] == 0
    np.testing.assert_array_equal(rd["tubes"][ok], rp["tubes"][ok])
    # every entry outside the pattern is zero in the dense tube, whatever the scenario
    mask = np.ones(n * (1 + g1), dtype=bool)
    mask[pat] = False
    assert not rd["tubes"][ok].reshape(-1, n * (1 + g1))[:, mask].any()
    # solve(): TubeHandle expands the packed buffer on access
    t.solver_options = tz.SolverOptions(tube_packed=1)
    cost, v, xbar, tube = t.solve(x0[0], np.zeros(n))
    t.solver_options = tz.SolverOptions()
    cost2, v2, xbar2, tube2 = t.solve(x0[0], np.zeros(n))
    np.testing.assert_array_equal(tube.Z.value, tube2.Z.value)
    assert tube.Z.value.shape == (n, 1 + g1)
    np.testing.assert_array_equal(tube.device_tensor.cpu().numpy(), tube2.device_tensor.cpu().numpy())


def test_from_datasets_batches_identification_and_gain_synthesis(cuda_lib):
    """from_datasets (one identify / gain-synthesis launch for all data sets) builds the controllers the per-data-set route builds."""
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS["fivedim"]()
    D, per = 4, 32
    data = [tz.Data(*common.dataset(cfg, seed=cfg.seed + 101 * d)) for d in range(D)]
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    ens = tz.TZDDPCEnsemble.from_datasets(data, zon, cfg.horizon, tz.StageCost(**cfg.cost), tz.BoxConstraint(**cfg.box),
                                          scenarios_per_dataset=per, num_initial_points=3, accuracy=0.05, confidence=1e-2)
    assert ens.theta_info["robust"].all() and ens.K.shape == (D, cfg.m, cfg.n)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (D * per, 1))
    r = ens.simulate(cfg.A, cfg.B, x0, steps=6, seed=3, restart=True)
    for d in range(D):
        c = tz.TZDDPC(data[d])
        c.verbose = False
        c.build_zonotopes_theta(zon, K=ens.K[d])
        c.build_problem(cfg.horizon, tz.StageCost(**cfg.cost), tz.BoxConstraint(**cfg.box))
        for a, b in ((c.Mdata, ens.controllers[d].Mdata), (c.MdataK, ens.controllers[d].MdataK), (c.Mdelta, ens.controllers[d].Mdelta)):
            np.testing.assert_allclose(a.center, b.center, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(a.generators, b.generators, rtol=1e-12, atol=1e-14)
        sl = slice(d * per, (d + 1) * per)
        rd = c.simulate(cfg.A, cfg.B, x0[sl], steps=6, seed=3, restart=True, scenario_offset=d * per)
        for k in ("x", "u", "cost", "status"):
            np.testing.assert_allclose(r[k][:, sl], rd[k], rtol=1e-9, atol=1e-9, err_msg=k)


@pytest.mark.parametrize("horizon,k0", [(3, None), (4, None), (5, None), (4, 2)])
def test_program_sets_on_the_larger_buckets(cuda_lib, horizon, k0):
    """Longer horizons / the simplified variant use the kernel buckets B1-B3: a program set over three data sets of the
    complexity-sweep system equals the three single-program solves, packed and dense tube alike."""
    import tzddpc_b200 as tz
    from tzddpc_b200 import _abi
    cfg = configs.sweep()
    ctls = []
    for d in range(3):
        u, x = common.dataset(cfg, seed=cfg.seed + 31 * d)
        o, K = common.make_oracle(cfg, u, x, horizon=horizon, k0=k0)
        ctls.append(common.make_product(cfg, u, x, K, horizon=horizon, k0=k0))
    assert not ctls[0]._program.bucket.startswith("B0")
    counts = [16, 32, 9]
    ens = tz.TZDDPCEnsemble(ctls, counts)
    S = sum(counts)
    rng = np.random.default_rng(horizon)
    Xi = o.zonotopes.X.interval
    xb = Xi.left_limit + (Xi.right_limit - Xi.left_limit) * rng.uniform(0.3, 0.7, (S, cfg.n))
    ee = rng.uniform(-0.01, 0.01, (S, cfg.n))
    dev = ens.device
    xbt, eet = torch.tensor(xb.T.copy(), device=dev), torch.tensor(ee.T.copy(), device=dev)
    for packed in (0, 1):
        opts = tz.SolverOptions(tube_packed=packed)
        r = ens.solve_batch(xbt, eet, options=opts)
        b = 0
        for c, cnt in zip(ctls, counts):
            rd = c.solve_batch(xbt[:, b:b + cnt].contiguous(), eet[:, b:b + cnt].contiguous(), options=opts)
            assert torch.equal(rd.status, r.status[b:b + cnt])
            for name in ("cost", "v", "xbar"):
                assert torch.equal(torch.nan_to_num(getattr(rd, name), nan=-7.0), torch.nan_to_num(getattr(r, name)[..., b:b + cnt], nan=-7.0)), name
            np.testing.assert_array_equal(rd.tube.Z.value, r.tube.Z.value[b:b + cnt])
            b += cnt
        assert int((r.status == 0).sum()) >= S // 2


def test_program_set_with_an_empty_data_set(cuda_lib):
    """A data set without scenarios in the middle of the set is skipped; its neighbours are unaffected."""
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS["pulley"]()
    ctls, _ = _controllers(cfg, 3)
    counts = [16, 0, 40]
    ens = tz.TZDDPCEnsemble(ctls, counts)
    S = sum(counts)
    rng = np.random.default_rng(2)
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    noise = common.noise_for(cfg, 4, S, rng)
    r = ens.simulate(cfg.A, cfg.B, x0, noise=noise, keep_tubes=True)
    for c, sl in ((ctls[0], slice(0, 16)), (ctls[2], slice(16, 56))):
        rd = c.simulate(cfg.A, cfg.B, x0[sl], noise=noise[:, sl], keep_tubes=True)
        for k in ("x", "u", "cost", "status", "tubes"):
            np.testing.assert_array_equal(r[k][:, sl], rd[k], err_msg=k)


@pytest.mark.parametrize("per", [8192, 8208])       # 512 tiles per data set (rounds never straddle) / 513 (they do)
def test_program_set_over_several_rounds(cuda_lib, per):
    """More tiles than one wave of warps (1,776): every CTA works through several rounds, re-stages programs between them
    and prefetches the next round's inputs -- still bit-equal to the per-data-set controllers, step after step."""
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS["fivedim"]()
    D, steps = 5, 4
    ctls, _ = _controllers(cfg, D)
    ens = tz.TZDDPCEnsemble(ctls, per)
    S = D * per
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    x0 += 0.01 * np.random.default_rng(8).standard_normal(x0.shape)
    opts = tz.SolverOptions(warm_start=2)
    r = ens.simulate(cfg.A, cfg.B, x0, steps=steps, seed=11, options=opts, restart=True, keep_tubes=False)
    for d, c in enumerate(ctls):
        sl = slice(d * per, (d + 1) * per)
        rd = c.simulate(cfg.A, cfg.B, x0[sl], steps=steps, seed=11, options=opts, restart=True, scenario_offset=d * per)
        for k in ("x", "xbar", "e", "u", "v", "cost", "status"):
            np.testing.assert_array_equal(r[k][:, sl], rd[k], err_msg=f"{k}, data set {d}")
    assert (r["status"] == 0).mean() > 0.9


@pytest.mark.parametrize("per", [16, 32, 64, 48])       # two program slots per warp / one program per 32- and 64-scenario CTA / mixed
def test_program_set_hot_path_with_and_without_helper_groups(cuda_lib, per, monkeypatch):
    """fast_step_kernel over a program set: the helper group of the small-block kernels (TZDDPC_FAST_SET_H, csrc/tz_fast.cuh)
    changes nothing in the results, and both equal the per-data-set controllers bit for bit."""
    import tzddpc_b200 as tz
    cfg = configs.CONFIGS["pulley"]()
    D, steps = 18, 8
    ctls, _ = _controllers(cfg, D)
    ens = tz.TZDDPCEnsemble(ctls, per)
    S = D * per
    x0 = np.tile(np.asarray(cfg.X0[0], dtype=np.float64), (S, 1))
    x0 += 0.01 * np.random.default_rng(4).standard_normal(x0.shape)
    opts = tz.SolverOptions(warm_start=2, hot_path=1)
    runs = {}
    for h in ("1", "2"):
        monkeypatch.setenv("TZDDPC_FAST_SET_H", h)
        runs[h] = ens.simulate(cfg.A, cfg.B, x0, steps=steps, seed=5, options=opts, restart=True, keep_tubes=True)
    for k in ("x", "xbar", "e", "u", "v", "cost", "status", "tubes"):
        np.testing.assert_array_equal(runs["1"][k], runs["2"][k], err_msg=k)
    monkeypatch.delenv("TZDDPC_FAST_SET_H")
    for d in (0, 7, D - 1):
        sl = slice(d * per, (d + 1) * per)
        rd = ctls[d].simulate(cfg.A, cfg.B, x0[sl], steps=steps, seed=5, options=opts, restart=True, scenario_offset=d * per, keep_tubes=True)
        for k in ("x", "u", "cost", "status", "tubes"):
            np.testing.assert_array_equal(runs["2"][k][:, sl], rd[k], err_msg=f"{k}, data set {d}")
