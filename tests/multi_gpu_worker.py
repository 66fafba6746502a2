"""Worker of tests/test_gpu_multi.py: one process per GPU under torch.distributed.run (NCCL).

Every rank builds the same tables and the same GLOBAL index / action streams, steps its contiguous shard of the
envs, runs the fused rollout with the logistic policy on it and all-reduces the 15 sufficient statistics; the
per-env results are gathered on rank 0, which also runs ALL envs on its own GPU and compares (SURVEY appendix B,
last row: 1 vs N GPU shards, same index/action streams).  Result -> JSON file given as argv[1]."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run_shard(core, idx, acts, lo, hi, dev, theta):
    """Step envs [lo, hi) for len(acts) steps (TABLE variant), then a fresh summer through the policy rollout."""
    from weather2alert_b200 import _lib

    n = hi - lo
    d = {k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).to(dev) for k, v in idx.items()}
    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), dtype=torch.float32, device=dev)
    reward, term = torch.zeros(n, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    rem, atb = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    core.reset(state, obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], remaining=rem)
    ret = torch.zeros(n, dtype=torch.float64, device=dev)
    for a in acts:
        core.step(state, torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev), obs, reward, term, rem, atb,
                  variant=_lib.STEP_TABLE)
        ret += reward.double()
    stepped = dict(state=state.view(2, n, 8).clone(), obs=obs.clone(), ret=ret, rem=rem.clone())
    # fused rollout of a fresh summer with the deterministic logistic policy, statistics accumulated on the device
    core.reset(state, obs, d["loc"], d["year"], d["coef_col"], d["coef_idx"], remaining=rem)
    ep_ret = torch.zeros(n, dtype=torch.float64, device=dev)
    stats = torch.zeros(15, dtype=torch.float64, device=dev)
    core.rollout_policy(state, core.T, ep_ret, theta, stochastic=False, seed=5, stats=stats)
    return stepped, ep_ret, stats


def main():
    out_path = sys.argv[1]
    n_total = int(sys.argv[2])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.distributed import reduce_stats, shard_bounds

    store, weights = bench.make_tables(years=4, counties=60, samples=16)
    core = HeatAlertCore(store, weights, dev)
    rng = np.random.default_rng(2024)  # identical on every rank: the GLOBAL streams
    Sw = len(weights.fips_list)
    col = (np.arange(n_total) % Sw).astype(np.int32)
    idx = dict(loc=np.array([store.fips_index[f] for f in weights.fips_list], np.int32)[col],
               year=rng.integers(0, store.shape[1], n_total).astype(np.int32), coef_col=col,
               coef_idx=rng.integers(0, weights.n_samples, n_total).astype(np.int32))
    acts = (rng.random((25, n_total)) < 0.4).astype(np.uint8)
    theta = torch.zeros(30, dtype=torch.float64)
    theta[0], theta[17], theta[25], theta[28], theta[29] = 2.0, 0.004, 0.05, -0.3, -2.2
    theta = theta.to(dev)

    lo, hi = shard_bounds(n_total, rank, world)
    stepped, ep_ret, stats = run_shard(core, idx, acts, lo, hi, dev, theta)
    local_stats = stats.clone()
    reduced = reduce_stats(stats.clone())  # the one collective of the path: NCCL all-reduce of float64[15]

    # gather the shards on rank 0 (padded to the largest shard)
    m = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))

    def gather(t):
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[: t.shape[0]] = t
        bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, bufs, dst=0)
        if rank != 0:
            return None
        return torch.cat([b[: shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0]] for r, b in enumerate(bufs)])

    g_state = gather(stepped["state"].permute(1, 0, 2).contiguous())  # [n, 2, 8] so that envs are the leading axis
    g_obs, g_ret, g_rem, g_ep = gather(stepped["obs"]), gather(stepped["ret"]), gather(stepped["rem"]), gather(ep_ret)
    all_stats = [torch.zeros_like(local_stats) for _ in range(world)] if rank == 0 else None
    dist.gather(local_stats, all_stats, dst=0)
    if rank == 0:
        whole, ep_whole, stats_whole = run_shard(core, idx, acts, 0, n_total, dev, theta)
        same = lambda a, b: bool(torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8)))  # noqa: E731
        counts = [0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]  # integer-valued statistics: exact in float64
        sums = [1, 2]
        res = {
            "world": world, "n_total": n_total,
            "state_bytes_equal": same(g_state, whole["state"].permute(1, 0, 2)),
            "obs_bytes_equal": same(g_obs, whole["obs"]),
            "returns_equal": same(g_ret, whole["ret"]), "remaining_equal": same(g_rem, whole["rem"]),
            "rollout_returns_equal": same(g_ep, ep_whole),
            "reduced_equals_sum_of_ranks": bool(torch.allclose(reduced, torch.stack(all_stats).sum(0), rtol=1e-15, atol=0)),
            "counts_equal_single_gpu": bool(torch.equal(reduced[counts], stats_whole[counts])),
            "sums_rel_err": float(((reduced[sums] - stats_whole[sums]).abs() / stats_whole[sums].abs()).max()),
            "episodes": float(reduced[0]), "nccl": dist.get_backend(),
        }
        with open(out_path, "w") as fh:
            json.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
