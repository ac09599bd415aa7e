"""N > 1 host logic on CPU: two processes over gloo shard a scenario batch, each produces the per-step statistics of
its shard, one all-reduce gives the statistics of the whole batch (the only collective of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tzddpc_b200 import shard


def test_shard_bounds_partition():
    for S in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            b = [shard.shard_bounds(S, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == S
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [s1 - s0 for s0, s1 in b]
            assert max(sizes) - min(sizes) <= 1


def _fake_stats(x):
    """(steps, S, n) states -> (steps, 8) statistics rows as tz_closed_loop_step accumulates them."""
    nrm = np.linalg.norm(x, axis=2)
    st = np.zeros((x.shape[0], 8))
    st[:, shard.STAT_SUM_NORM] = nrm.sum(1)
    st[:, shard.STAT_SUM_NORM2] = (nrm ** 2).sum(1)
    st[:, shard.STAT_COUNT] = x.shape[1]
    return st


def _worker(rank, world, port, S, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = np.random.default_rng(123).normal(size=(5, S, 4))          # the same global batch on every rank
        s0, s1 = shard.shard_bounds(S, rank, world)
        st = torch.from_numpy(_fake_stats(x[:, s0:s1]))
        shard.reduce_statistics(st)
        t = shard.max_over_ranks(1.0 + rank)
        assert shard.world_info() == (rank, world)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.r_[st.numpy().ravel(), t])
    finally:
        dist.destroy_process_group()


def test_two_rank_statistics_reduce(tmp_path):
    world, S = 2, 37
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, S, str(tmp_path)), nprocs=world, join=True)
    x = np.random.default_rng(123).normal(size=(5, S, 4))
    want = _fake_stats(x)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        np.testing.assert_allclose(got[:-1].reshape(5, 8), want, rtol=1e-13)
        assert got[-1] == 2.0                       # max over ranks of (1 + rank)
    sm = shard.summarise(want)
    nrm = np.linalg.norm(x, axis=2)
    np.testing.assert_allclose(sm.mean_norm, nrm.mean(1), rtol=1e-12)
    np.testing.assert_allclose(sm.ci95, 1.96 * nrm.std(1) / np.sqrt(S), rtol=1e-9)
