"""Oracle unit tests: identities that hold whatever the [R] conventions of SURVEY.md App. A are, and the analytic
known answers of SURVEY.md 4.3 / 8(c)."""
import numpy as np
import pytest

import oracle
from oracle.zono import girard_reduce_generators, matzono_times_Z
from tests import common
from tzddpc_b200 import configs


def _rand_zono(rng, n, g):
    return oracle.Zonotope(rng.normal(size=n), rng.normal(size=(n, g)))


def test_interval_of_minkowski_sum_is_sum_of_intervals():
    rng = np.random.default_rng(0)
    a, b = _rand_zono(rng, 4, 7), _rand_zono(rng, 4, 3)
    s = (a + b).interval
    np.testing.assert_allclose(s.left_limit, a.interval.left_limit + b.interval.left_limit, rtol=1e-14)
    np.testing.assert_allclose(s.right_limit, a.interval.right_limit + b.interval.right_limit, rtol=1e-14)
    assert (a + b).num_generators == 10


def test_mul_is_left_multiplication():
    """`Z * K` must be m-dimensional (tzddpc/tzddpc.py:192 compares it with U.interval)."""
    rng = np.random.default_rng(1)
    z = _rand_zono(rng, 4, 5)
    K = rng.normal(size=(1, 4))
    zk = z * K
    assert zk.dimension == 1
    np.testing.assert_allclose(zk.Z, K @ z.Z)


@pytest.mark.parametrize("metric", ["l1-linf", "l1", "l2"])
@pytest.mark.parametrize("order", [1, 1.5, 2, 3])
def test_girard_reduce_contains_input_and_has_right_size(metric, order):
    rng = np.random.default_rng(2)
    n, g = 3, 40
    conv = oracle.Conventions(girard_metric=metric)
    z = oracle.Zonotope(rng.normal(size=n), rng.normal(size=(n, g)) * rng.uniform(0.01, 1, size=g), conv)
    r = z.reduce(order)
    assert r.num_generators == int(np.floor(n * (order - 1))) + n
    for _ in range(50):                      # support function of the reduced set dominates
        d = rng.normal(size=n)
        assert r.support(d) >= z.support(d) - 1e-12
    # boxing preserves the interval hull exactly up to rounding
    np.testing.assert_allclose(r.interval.right_limit, z.interval.right_limit, rtol=1e-12)


def test_girard_reduce_noop_and_zero_filter():
    G = np.array([[1.0, 0.0, 2.0], [0.5, 0.0, -1.0]])
    out = girard_reduce_generators(G, 2)             # 2 non-zero generators <= order * n
    np.testing.assert_array_equal(out, G[:, [0, 2]])


def test_order_one_reduce_is_the_interval_box():
    rng = np.random.default_rng(3)
    z = _rand_zono(rng, 5, 23)
    r = z.reduce(1)
    d = np.abs(z.generators).sum(axis=1)
    np.testing.assert_allclose(r.generators, np.diag(d), rtol=1e-13)


def test_girard_picks_smallest_metric_and_keeps_order():
    G = np.array([[3.0, 0.1, 2.0, 0.2, 1.0], [3.0, 0.1, -2.0, 0.3, 0.0]])
    # l1 - linf: [3, .1, 2, .2, 0]; order 2, n = 2 -> keep 2, box 3 smallest metric: columns 4, 1, 3
    out = girard_reduce_generators(G, 2)
    np.testing.assert_allclose(out[:, :2], G[:, [0, 2]])
    np.testing.assert_allclose(out[:, 2:], np.diag([1.3, 0.4]))


def test_matzono_product_layout_and_count():
    rng = np.random.default_rng(4)
    n, N, g = 3, 4, 5
    C, Gm, Z = rng.normal(size=(n, n)), rng.normal(size=(N, n, n)), rng.normal(size=(n, 1 + g))
    out = matzono_times_Z(C, Gm, Z)
    assert out.shape == (n, (N + 1) * (g + 1))                  # (N+1)(g+1) - 1 generators + centre
    np.testing.assert_allclose(out[:, :1 + g], C @ Z)
    np.testing.assert_allclose(out[:, (1 + g) * 2:(1 + g) * 3], Gm[1] @ Z)
    # the set is right: a sampled point M z lies inside (support function test)
    M = C + np.tensordot(rng.uniform(-1, 1, N), Gm, axes=(0, 0))
    zpt = Z[:, 0] + Z[:, 1:] @ rng.uniform(-1, 1, g)
    zo = oracle.Zonotope(out[:, 0], out[:, 1:])
    for _ in range(30):
        d = rng.normal(size=n)
        assert d @ (M @ zpt) <= zo.support(d) + 1e-12


def test_identification_noise_free_recovers_the_plant():
    """W = 0 data: centre of M_Sigma = [A B] exactly (least squares), no generators of non-zero size."""
    cfg = configs.pulley()
    rng = np.random.default_rng(5)
    n, m, T = cfg.n, cfg.m, 60
    u = rng.uniform(-1, 1, size=(T, m))
    x = np.zeros((T, n))
    x[0] = rng.normal(size=n)
    for t in range(1, T):
        x[t] = cfg.A @ x[t - 1] + cfg.B @ u[t - 1]
    W = oracle.Zonotope(np.zeros(n), 0.1 * np.ones((n, 1)))
    Mw = oracle.concatenate_zonotope(W, T - 1)
    M = oracle.compute_LTI_matrix_zonotope(x[:-1], x[1:], u[:-1], Mw)
    np.testing.assert_allclose(M.center, np.hstack([cfg.A, cfg.B]), atol=1e-9)
    assert M.num_generators == T - 1


@pytest.mark.parametrize("name,expect", [("double_integrator", (198, 6, 4, 24, 64)), ("pulley", (399, 20, 16, 75, 619)),
                                         ("fivedim", (399, 30, 25, 113, 1413))])
def test_generator_counts_match_survey_table(name, expect):
    """SURVEY.md section 8 table = what tzddpc/tzddpc.py:206 prints."""
    cfg = configs.CONFIGS[name]()
    u, x = common.dataset(cfg)
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    z = common.oracle_zonotopes(cfg)
    o.build_zonotopes(z)
    raw = o.Mdata.num_generators
    C = o.Mdata.center
    o.build_zonotopes_theta(z, configs.lqr_gain(C[:, :cfg.n], C[:, cfg.n:]))
    o.build_problem(2, oracle.StageCost(**cfg.cost), oracle.BoxConstraint(**cfg.box) if cfg.box else None)
    assert (raw, o.Mdelta.num_generators, o.MdataK.num_generators, *o.num_generators_log) == expect


def test_order_one_box_closed_form():
    """App. A.6: the order-1 box of the rank-one generators -g_k P[j,:] is (sum_k |g_k|)(sum_j |P[j,:]|)'."""
    cfg = configs.double_integrator()
    u, x = common.dataset(cfg)
    o = oracle.OracleTZDDPC(oracle.Data(u, x))
    z = common.oracle_zonotopes(cfg)
    o.build_zonotopes(z)
    D = np.vstack([x[:-1].T, u[:-1].T])
    P = np.linalg.pinv(D)
    gsum = np.abs(z.W.generators).sum(axis=1)
    d = np.outer(gsum, np.abs(P).sum(axis=0))
    C = o.Mdata.center
    K = configs.lqr_gain(C[:, :cfg.n], C[:, cfg.n:])
    o.build_zonotopes_theta(z, K)
    box = np.abs(o.Mdelta.generators).sum(axis=0)
    np.testing.assert_allclose(box, d, rtol=1e-10)
    dK = np.outer(gsum, np.abs(P @ np.vstack([np.eye(cfg.n), K])).sum(axis=0))
    np.testing.assert_allclose(np.abs(o.MdataK.generators).sum(axis=0), dK, rtol=1e-10)
    # every boxed generator has a single non-zero entry
    assert all(np.count_nonzero(G) <= 1 for G in o.Mdelta.generators)


def test_vertices_and_sample():
    W = oracle.Zonotope(np.zeros(2), 0.1 * np.array([[1.0, 0.5], [0.5, 1.0]]))
    V = W.compute_vertices()
    assert V.shape == (4, 2)
    rng = np.random.default_rng(0)
    s = W.sample(1000, rng)
    iv = W.interval
    assert np.all(s >= iv.left_limit - 1e-15) and np.all(s <= iv.right_limit + 1e-15)
