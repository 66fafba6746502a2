"""world_size-2 gloo tests (CPU) of the multi-rank host logic: sharding + the stat reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from weather2alert_b200.distributed import (N_STATS, RolloutStats, reduce_stats, shard_bounds, summarize)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_rollout(lo, hi, n_total, steps=40, seed=3):
    """Deterministic per-env streams so that shards of any split see the same data."""
    g = np.random.default_rng(seed)
    act = g.random((steps, n_total)) < 0.4
    rew = -g.random((steps, n_total))
    atb = g.random((steps, n_total)) < 0.2
    term = np.zeros((steps, n_total), bool)
    term[19] = True
    term[39] = True
    st = RolloutStats(hi - lo, "cpu")
    for t in range(steps):
        st.update(torch.from_numpy(act[t, lo:hi]), torch.from_numpy(rew[t, lo:hi]).float(),
                  torch.from_numpy(term[t, lo:hi]), torch.from_numpy(atb[t, lo:hi]))
    return st


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_total, rank, world)
        red = _fake_rollout(lo, hi, n_total).reduce()
        if rank == 0:
            out.put(red.tolist())
    finally:
        dist.destroy_process_group()


def test_two_rank_stat_reduction_equals_single_rank():
    n_total, world = 1001, 2
    whole = _fake_rollout(0, n_total, n_total).totals
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(np.array(got), whole.numpy(), rtol=1e-12)
    s = summarize(torch.tensor(got, dtype=torch.float64))
    assert s["episodes"] == 2 * n_total and s["steps"] == 40 * n_total
    assert 0.25 < s["alert_freq"] < 0.4 and 0.04 < s["over_budget_freq"] < 0.12


def test_reduce_stats_without_process_group_is_identity():
    v = torch.arange(N_STATS, dtype=torch.float64)
    assert torch.equal(reduce_stats(v.clone()), v)
    with pytest.raises(ValueError):
        reduce_stats(torch.zeros(3, dtype=torch.float64))
