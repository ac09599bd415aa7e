"""Scenario data-parallelism: one process per GPU, contiguous scenario shards, no collective in the
step loop, ONE all-reduce of the closed-loop statistics after the last step.

The reference runs its scenarios strictly one after another in one process
(examples/2.pulley_sim.py:62-103) and reduces them in a notebook
(examples/2.pulley_analyse_results.ipynb: E||x_t|| +- 1.96 sigma/sqrt(N)); nothing in `solve`
reads another scenario, so sharding needs no exchange.  Works with any torch.distributed
backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

# layout of the per-step statistics row accumulated by tz_closed_loop_step (include/tzddpc.h)
STAT_SUM_NORM, STAT_SUM_NORM2, STAT_SUM_COST, STAT_INFEASIBLE, STAT_MAXITER, STAT_SUM_ITERS, STAT_NONFINITE, STAT_COUNT = range(8)


def shard_bounds(num_scenarios: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [s0, s1) of rank `rank`: the first `num_scenarios % world` ranks get one extra scenario."""
    assert 0 <= rank < world and num_scenarios >= 0
    base, extra = divmod(num_scenarios, world)
    s0 = rank * base + min(rank, extra)
    return s0, s0 + base + (1 if rank < extra else 0)


def world_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def reduce_statistics(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the (steps, TZ_NSTATS) statistics of all ranks, in place; identity for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    """Timing rule of the benchmark: the slowest rank defines the step time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


@dataclass
class ClosedLoopSummary:
    """What examples/2.pulley_analyse_results.ipynb plots, from the reduced statistics."""
    mean_norm: np.ndarray        # E||x_t||_2 per step over the scenarios that solved
    ci95: np.ndarray             # 1.96 sigma / sqrt(N)
    mean_cost: np.ndarray
    infeasible: np.ndarray
    maxiter: np.ndarray
    mean_iters: np.ndarray
    count: np.ndarray


def summarise(stats) -> ClosedLoopSummary:
    s = np.asarray(stats.cpu() if isinstance(stats, torch.Tensor) else stats, dtype=np.float64).reshape(-1, 8)
    cnt = s[:, STAT_COUNT]
    ok = np.maximum(cnt - s[:, STAT_INFEASIBLE] - s[:, STAT_NONFINITE], 1.0)
    mean = s[:, STAT_SUM_NORM] / ok
    var = np.maximum(s[:, STAT_SUM_NORM2] / ok - mean ** 2, 0.0)
    return ClosedLoopSummary(mean, 1.96 * np.sqrt(var / ok), s[:, STAT_SUM_COST] / ok, s[:, STAT_INFEASIBLE],
                             s[:, STAT_MAXITER], s[:, STAT_SUM_ITERS] / np.maximum(cnt, 1.0), cnt)
