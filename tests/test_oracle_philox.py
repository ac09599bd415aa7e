"""The random stream shared by oracle and GPU: numpy Philox4x32-10 against the published known-answer vectors of
Random123 (kat_vectors: philox4x32 10) and against the C restatement exported by libtzddpc.so (host function, no GPU)."""
import ctypes as C

import numpy as np

from oracle import philox

KAT = [  # counter (4), key (2) -> output (4)
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_numpy_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.array(key), np.array(ctr))
        assert tuple(int(v) for v in got) == want


def test_library_host_philox_matches(cuda_lib):
    out = (C.c_uint32 * 4)()
    rng = np.random.default_rng(0)
    cases = [(c, k) for c, k, _ in KAT] + [(tuple(int(v) for v in rng.integers(0, 2**32, 4)), tuple(int(v) for v in rng.integers(0, 2**32, 2)))
                                          for _ in range(50)]
    for ctr, key in cases:
        cuda_lib.tz_philox4x32_10_host(key[0], key[1], ctr[0], ctr[1], ctr[2], ctr[3], out)
        want = philox.philox4x32_10(np.array(key), np.array(ctr))
        assert [int(v) for v in out] == [int(v) for v in want]


def test_draws_are_uniform_and_shard_invariant():
    b = philox.draws(25, np.arange(4096), 7, 0, 3, False)
    assert b.shape == (4096, 3) and b.min() >= -1.0 and b.max() < 1.0
    assert abs(b.mean()) < 0.03 and abs(b.std() - 1 / np.sqrt(3)) < 0.02
    # the draw of scenario i does not depend on the batch it is computed in
    np.testing.assert_array_equal(philox.draws(25, np.arange(1000, 1010), 7, 0, 3, False), b[1000:1010])
    v = philox.draws(25, np.arange(4096), 7, 2, 2, True)
    assert set(np.unique(v)) == {-1.0, 1.0} and abs(v.mean()) < 0.05
    assert not np.array_equal(philox.draws(25, np.arange(8), 7, 0, 1, False), philox.draws(25, np.arange(8), 8, 0, 1, False))
