"""Env sharding and the one collective of the path: the reduction of return statistics.

Envs are independent (all state of the reference env lives in the env object,
`/root/reference/src/weather2alert/env.py:156-180`), so N GPUs step N contiguous shards
with replicated tables and no step-time communication.  What is worth reducing are the
rollout statistics the reference's SB3 callback logs
(`/root/reference/src/weather2alert/callbacks.py:63-74`: mean return, over-budget
frequency, alert frequency, ...); their sufficient statistics are a short float64 vector
that is all-reduced once per rollout (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

STAT_NAMES = ("episodes", "return_sum", "return_sq_sum", "steps", "alerts", "refused_at_budget",
              "alert_day_sum", "alert_day_sq_sum", "streaks", "streak_sum", "streak_sq_sum",
              "a50_sum", "a80_sum", "a100_sum", "a_count")
N_STATS = len(STAT_NAMES)


def shard_bounds(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_total` envs for `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} not in [0, {world})")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class RolloutStats:
    """Device-side accumulator of the return statistics of one rank's envs."""

    def __init__(self, num_envs: int, device):
        self.ep_return = torch.zeros(num_envs, dtype=torch.float64, device=device)
        self.totals = torch.zeros(N_STATS, dtype=torch.float64, device=device)
        # a device constant, so that `update` creates no tensor from host data (CUDA-graph capturable)
        self._n = torch.full((), float(num_envs), dtype=torch.float64, device=device)

    def update(self, actions, reward, terminated, at_budget):
        """One step: `actions` any integer/bool dtype, `at_budget` the pre-action flag."""
        a = actions.bool() if actions.dtype != torch.bool else actions
        atb = at_budget.bool()
        self.ep_return += reward.double()
        done = terminated.bool()
        fin = torch.where(done, self.ep_return, torch.zeros_like(self.ep_return))
        upd = torch.stack([done.sum().double(), fin.sum(), (fin * fin).sum(), self._n,
                           (a & ~atb).sum().double(), (a & atb).sum().double()])
        self.totals[:6] += upd  # the timing statistics (6..14) are produced by the rollout kernel only
        self.ep_return.masked_fill_(done, 0.0)

    def reduce(self, group=None) -> torch.Tensor:
        return reduce_stats(self.totals.clone(), group)


def reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a `[N_STATS]` float64 vector over all ranks (no-op without a process group)."""
    if stats.shape[-1] != N_STATS or stats.dtype != torch.float64:
        raise ValueError(f"expected float64[..., {N_STATS}]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats: torch.Tensor) -> dict:
    """The quantities `AlertLoggingCallback` reports (callbacks.py:63-74) from the reduced vector."""
    s = dict(zip(STAT_NAMES, (float(v) for v in stats.tolist())))
    n_ep, n_steps = max(s["episodes"], 1.0), max(s["steps"], 1.0)
    mean = s["return_sum"] / n_ep
    var = max(s["return_sq_sum"] / n_ep - mean * mean, 0.0)
    out = {**s, "mean_return": mean, "std_return": var ** 0.5,
           "alert_freq": s["alerts"] / n_steps, "over_budget_freq": s["refused_at_budget"] / n_steps}

    def mean_std(total, sq, count):
        if count <= 0:
            return float("nan"), float("nan")
        m = total / count
        return m, max(sq / count - m * m, 0.0) ** 0.5

    out["average_t_alerts"], out["stdev_t_alerts"] = mean_std(s["alert_day_sum"], s["alert_day_sq_sum"], s["alerts"])
    out["average_streak"], out["stdev_streak"] = mean_std(s["streak_sum"], s["streak_sq_sum"], s["streaks"])
    for k in ("a50", "a80", "a100"):
        out[f"alert_t_{k[1:]}%"] = s[f"{k}_sum"] / s["a_count"] if s["a_count"] > 0 else float("nan")
    return out


def rollout_hash(seed: int, envs, steps):
    """The counter-based 32-bit stream of the rollout kernels on the host: ``u[k, i] = hash(seed, envs[i],
    steps[k])`` (splitmix64 finaliser; same arithmetic as `rollout_hash` in csrc/w2a_kernels.cuh).
    Returns uint64 [len(steps), len(envs)] holding values below 2**32."""
    import numpy as np

    e = np.asarray(envs, dtype=np.uint64)[None, :]
    k = np.asarray(steps, dtype=np.uint64)[:, None]
    with np.errstate(over="ignore"):
        z = np.uint64(seed & (2**64 - 1)) + np.uint64(0x9E3779B97F4A7C15) * ((e << np.uint64(20)) ^ k ^ np.uint64(0x5851F42D4C957F2D))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z >> np.uint64(32)


def rollout_threshold(p):
    """Probability -> the 32-bit threshold the kernels compare the hash with (`p * 2**32`, saturated)."""
    import numpy as np

    thr = np.asarray(p, dtype=np.float64) * 4294967296.0
    return np.where(thr >= 4294967295.0, 4294967295.0, thr).astype(np.uint64)


def rollout_action_bits(seed: int, envs, steps, p_alert: float):
    """The action stream of `w2a_rollout(actions=NULL)` on the host: ``bits[k, i] = rollout_hash(seed,
    envs[i], steps[k]) < p_alert * 2**32``.  Returns bool [len(steps), len(envs)]."""
    return rollout_hash(seed, envs, steps) < rollout_threshold(p_alert)
