#!/usr/bin/env python
"""Host-link ceiling of the box: plain pinned cudaMemcpyAsync device->host (and host->device) of the e2e path's
block size, 1..N ranks copying at once (one process per GPU, NUMA-bound like bench.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bench.py [--mib 126]

Rank 0 prints one JSON line: per-rank and aggregate GB/s with all N ranks copying at the same time (barrier before,
max time over ranks), plus the same copy split into the five arrays and 16 chunks `w2a_step_host` issues per step.
This is the denominator of `e2e.frac_of_link`: what the link gives without any of our code."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=126)  # 2^20 envs x 126 B of step outputs
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import bench

    bench.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = args.mib << 20
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.zero_()  # first touch on this rank's NUMA node

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return nbytes * args.iters / float(t.item()) / 1e9

    streams = [torch.cuda.Stream(dev) for _ in range(3)]

    def chunked():  # the copy pattern of w2a_step_host: 16 chunks x 5 arrays over 3 streams
        n = 1 << 20
        sizes = [n * 116, n * 4, n, n * 4, n]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        for c in range(16):
            with torch.cuda.stream(streams[c % 3]):
                for a in range(5):
                    per = sizes[a] // 16
                    lo = offs[a] + c * per
                    if lo + per <= nbytes:
                        h[lo:lo + per].copy_(d[lo:lo + per], non_blocking=True)
        for s in streams:
            s.synchronize()

    res = {"world": world, "mib": args.mib,
           "d2h_gbs_per_rank": timed(lambda: h.copy_(d, non_blocking=True)),
           "h2d_gbs_per_rank": timed(lambda: d.copy_(h, non_blocking=True)),
           "d2h_chunked_gbs_per_rank": timed(chunked)}
    res["d2h_gbs_aggregate"] = res["d2h_gbs_per_rank"] * world
    res["d2h_chunked_gbs_aggregate"] = res["d2h_chunked_gbs_per_rank"] * world
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
