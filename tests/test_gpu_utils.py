"""tzddpc_b200.utils: the reference's gain helpers (tzddpc/utils.py:8-129) as wrappers over the batched CUDA routines."""
import numpy as np
import pytest

from tests import common
from tzddpc_b200 import configs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctl(cuda_lib):
    import tzddpc_b200 as tz
    cfg = configs.pulley()
    u, x = common.dataset(cfg)
    t = tz.TZDDPC(tz.Data(u, x))
    t.verbose = False
    Z = tz.Zonotope
    zon = tz.SystemZonotopes(Z(*cfg.X0), Z(*cfg.U), Z(*cfg.X), Z(*cfg.W))
    t.build_zonotopes(zon)
    return cfg, t


def test_compute_control_gain_is_the_lqr_gain(ctl):
    import tzddpc_b200 as tz
    cfg, t = ctl
    C = t.Mdata.center
    A0, B0 = C[:, :cfg.n], C[:, cfg.n:]
    K = tz.compute_control_gain(A0, B0)
    np.testing.assert_allclose(K, configs.lqr_gain(A0, B0), rtol=1e-8, atol=1e-10)
    assert tz.spectral_radius(A0 + B0 @ K) < 1.0


def test_adversary_robustness_and_theta(ctl):
    import tzddpc_b200 as tz
    from oracle import gain as ogain
    cfg, t = ctl
    n = cfg.n
    C = t.Mdata.center
    A0, B0 = C[:, :n], C[:, n:]
    K = tz.compute_control_gain(A0, B0)
    An, Bn = tz.compute_A_B(t.Mdata, K, num_init=4)
    assert An.shape == (n, n) and Bn.shape == (n, cfg.m)
    # the adversary maximises ||A + B K||_F over the set: not below the centre's value
    assert np.linalg.norm(An + Bn @ K) >= np.linalg.norm(A0 + B0 @ K) - 1e-12
    assert tz.is_gain_robust(t.Mdata, K, 1e-2, 1e-5) is True
    assert tz.is_gain_robust(t.Mdata, 0.0 * K + 50.0, 1e-2, 1e-5) is False            # a destabilising gain
    th = tz.compute_theta(t.Mdata, A0, B0)
    assert isinstance(th, tz.Theta) and th.K.shape == (cfg.m, n)
    # same routine as TZDDPC.compute_theta() without K (oracle/gain.py is its CPU restatement)
    th2 = t.compute_theta()
    np.testing.assert_array_equal(th.K, th2.K)
    Pinv = np.linalg.pinv(np.hstack([t.dataset.Xm, t.dataset.Um]).T)
    ref = ogain.gain_synthesis(C, Pinv, t.zonotopes.W.Z)
    np.testing.assert_allclose(th.K, ref["K"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(th.deltaA, ref["dA"], rtol=1e-7, atol=1e-9)
    # compute_A_B for the final gain = the adversary of the oracle for that gain
    An2, Bn2, _ = ogain.adversary(A0, B0, Pinv, t.zonotopes.W.Z[:, 1:], th.K, 10, 25, 0)
    An3, Bn3 = tz.compute_A_B(t.Mdata, th.K, num_init=10)
    np.testing.assert_allclose(An3, An2, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(Bn3, Bn2, rtol=1e-8, atol=1e-10)


def test_hand_built_matrix_zonotope_is_refused(cuda_lib):
    import tzddpc_b200 as tz
    M = tz.MatrixZonotope(np.eye(2, 3), np.zeros((1, 2, 3)))
    with pytest.raises(NotImplementedError, match="rank-one structure"):
        tz.compute_A_B(M, np.zeros((1, 2)))
