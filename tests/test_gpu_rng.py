"""The Philox stream on the GPU (tz_sample_noise, tz_generate_trajectories) against its numpy restatement: random bits are
integers, so the noise is bit-exact; the data sets agree to rounding; draws are shard-invariant."""
import numpy as np
import pytest

from oracle import philox
from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T(cuda_lib):
    import torch
    from tzddpc_b200 import ops  # noqa: F401
    return torch


def _f(T, a):
    return T.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("name,vertex", [("pulley", False), ("fivedim", False), ("double_integrator", True)])
def test_sample_noise_bit_exact_and_shard_invariant(T, name, vertex):
    cfg = configs.CONFIGS[name]()
    WZ = np.hstack([cfg.W[0][:, None], cfg.W[1]])
    S, seed = 1000, 25
    for t in (0, 3, 199):
        got = T.ops.tzddpc.sample_noise(_f(T, WZ), S, vertex, seed, 0, t).cpu().numpy().T
        want = philox.sample_noise(WZ, S, vertex, seed, 0, t)
        np.testing.assert_allclose(got, want, rtol=1e-15, atol=1e-17)
        # a shard that starts at scenario 600 sees the same draws
        part = T.ops.tzddpc.sample_noise(_f(T, WZ), 400, vertex, seed, 600, t).cpu().numpy().T
        np.testing.assert_array_equal(part, got[600:])
    iv = np.abs(cfg.W[1]).sum(axis=1)
    assert np.all(np.abs(got - cfg.W[0]) <= iv + 1e-15)


@pytest.mark.parametrize("name", ["double_integrator", "pulley", "fivedim"])
def test_generate_trajectories_matches_numpy(T, name):
    cfg = configs.CONFIGS[name]()
    Z = lambda z: np.hstack([z[0][:, None], z[1]])          # noqa: E731
    S, Tn, seed = 37, cfg.T, 7
    U, X = T.ops.tzddpc.generate_trajectories(_f(T, cfg.A), _f(T, cfg.B), _f(T, Z(cfg.X0)), _f(T, Z(cfg.U)), _f(T, Z(cfg.W)), S, Tn, seed, 0)
    Uo, Xo = philox.generate_trajectories(cfg.A, cfg.B, Z(cfg.X0), Z(cfg.U), Z(cfg.W), S, Tn, seed, 0)
    np.testing.assert_allclose(U.cpu().numpy(), Uo, rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(X.cpu().numpy(), Xo, rtol=1e-9, atol=1e-9)
    assert np.all(X.cpu().numpy()[:, 0] == 0.0)                 # quirk Q9: the first returned state row is the origin
    # the device-generated data sets feed tz_identify directly (the 'datasets' scenario axis): centre close to the plant
    WZ = _f(T, Z(cfg.W))
    AB, dAB, dK, _, status = T.ops.tzddpc.identify(X, U, WZ, None, False)
    assert (status.cpu().numpy() == 0).all()
    Xh, Uh = X.cpu().numpy(), U.cpu().numpy()
    for s_ in (0, S - 1):
        D = np.vstack([Xh[s_, :-1].T, Uh[s_, :-1].T])
        C = (Xh[s_, 1:].T - cfg.W[0][:, None]) @ np.linalg.pinv(D)          # (X1 - c_W) pinv([X0; U0]), tzddpc/tzddpc.py:83
        np.testing.assert_allclose(AB[s_].cpu().numpy(), C, rtol=1e-8, atol=1e-10)


def test_simulate_with_seed_is_reproducible_and_matches_oracle(T):
    cfg = configs.pulley()
    u, x = common.dataset(cfg)
    o, K = common.make_oracle(cfg, u, x)
    t = common.make_product(cfg, u, x, K)
    S, steps, seed = 16, 10, 99
    x0 = np.zeros((S, cfg.n))
    a = t.simulate(cfg.A, cfg.B, x0, steps=steps, seed=seed)
    b = t.simulate(cfg.A, cfg.B, x0[4:9], steps=steps, seed=seed, scenario_offset=4)
    np.testing.assert_array_equal(a["x"][:, 4:9], b["x"])       # shard invariance, bitwise
    WZ = np.hstack([cfg.W[0][:, None], cfg.W[1]])
    noise = np.stack([philox.sample_noise(WZ, S, False, seed, 0, k) for k in range(steps)])
    r = o.closed_loop(cfg.A, cfg.B, x0[3], noise[:, 3])
    np.testing.assert_allclose(a["x"][:, 3], r["x"], rtol=1e-6, atol=1e-6)
