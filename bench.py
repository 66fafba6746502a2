#!/usr/bin/env python
"""bench.py -- env-steps/s of the HeatAlertEnv step hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3_1m] [--impl ours|reference]

One "step" is one `env.step` over every env of the batch (all GPUs); the timed region is
EXACTLY K steps after W warm-up steps, CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Rank 0 prints ONE JSON line.

Workloads (BASELINE.json `configs`; DESIGN.md "Measurement" says why cfg3_1m is the default):
  cfg2_4096  4 096 envs/GPU, env i -> county i mod S            (configs[1]; launch-latency bound)
  cfg3_1m    2**20 envs/GPU, env i -> county i mod S            (configs[2] size; default)
  cfg4_2m    2**21 envs/GPU, similar_climate_counties=True      (configs[3] = 16M envs / 8 GPUs)
  cfg5_8m    2**23 envs/GPU                                     (configs[4] = 64M envs / 8 GPUs)
All use the linear-format checkpoint layout (S=746 counties, P=100 draws; `nn_*`
checkpoints have the identical layout and take the same path), 11
years x 153 days of synthetic weather, uniform random actions (p=0.5) resident in HBM.

`--impl reference` times the reference's own CPU implementation on the host cores: the
UNMODIFIED reference env from baseline/_ref (one forked worker per core) when that
install is present, else the C port in oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
S_COUNTIES, N_YEARS, P_SAMPLES = 746, 11, 100
WORKLOADS = {
    "cfg2_4096": dict(envs_per_gpu=4096, augment=False, desc="BASELINE configs[1]: 4096 envs, env i -> county i mod S"),
    "cfg3_1m": dict(envs_per_gpu=1 << 20, augment=False, desc="BASELINE configs[2] size: 2^20 envs/GPU, env i -> county i mod S"),
    "cfg4_2m": dict(envs_per_gpu=1 << 21, augment=True, desc="BASELINE configs[3]: 2^21 envs/GPU (16M on 8), similar_climate_counties"),
    "cfg5_8m": dict(envs_per_gpu=1 << 23, augment=False, desc="BASELINE configs[4]: 2^23 envs/GPU (64M on 8)"),
}
# Algorithmic bytes per env-step (DESIGN.md "Roofline"):
#  streaming: what must cross HBM every step -- action 1 + state 16 read, state 8 + obs 116 +
#             reward 4 + terminated 1 + remaining 4 + at_budget 1 written = 151
#  canonical: SURVEY.md 8(d) = 466, which also counts the feature row (96) and the 56
#             coefficients (224) that this design serves from L2 instead of HBM
BYTES_STREAMING = 1 + 16 + 8 + 116 + 4 + 1 + 4 + 1
BYTES_CANONICAL = 466


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_tables(years=N_YEARS, counties=S_COUNTIES, samples=P_SAMPLES):
    from weather2alert_b200 import synthetic

    fips, zones = synthetic.synthetic_counties(counties, seed=0)
    store = synthetic.make_store(fips, list(range(2006, 2006 + years)), seed=1, zones=zones)
    weights = synthetic.make_weights(fips, samples, seed=2)
    return store, weights


def index_streams(store, weights, n, rank, augment, locality="interleaved"):
    """env i -> county (global i) mod S; year / posterior sample / similar-list index drawn with
    numpy's default_rng(1234 + rank).  locality="sorted" assigns
    the SAME multiset of episodes to env slots in (county, sample, year) order."""
    from weather2alert_b200.sampling import EpisodeSampler

    rng = np.random.default_rng(1234 + rank)
    S = len(weights.fips_list)
    list_idx = (np.arange(n, dtype=np.int64) + rank * n) % S
    sp = EpisodeSampler(store, weights)
    s = sp.draw_bulk(rng, n, list_idx, augment)
    if locality == "sorted":
        order = np.lexsort((s["year"], s["coef_idx"], s["coef_col"], s["loc"]))
        s = {k: np.ascontiguousarray(v[order]) for k, v in s.items()}
    return s


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPU cores next to its GPU (NVML's ideal affinity), so that the pinned host
    buffers of the e2e path are first-touched on that NUMA node and PCIe DMA does not cross sockets."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            log(f"[gpu {index}] bound to {len(cpus)} cores of its NUMA node")
    except Exception as exc:  # pragma: no cover - best effort
        log(f"[gpu {index}] no NUMA binding: {exc}")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.stop_flag, self.ok = [], threading.Event(), False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # pragma: no cover
            log(f"clock sampling unavailable: {exc}")

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), sm, int(reasons)))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost",
                 0x100: "display_clock_setting"}
        bits = 0
        for _, _, r in inside:
            bits |= r
        reasons = [n for b, n in names.items() if bits & b]
        return {"sm_mhz": statistics.median([s[1] for s in inside]) if inside else None,
                "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(inside)}


# =================================================================================================
# our arm
# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    n = args.envs_per_gpu or wl["envs_per_gpu"]
    K, W = args.steps, args.warmup
    t_setup = time.time()
    store, weights = make_tables()
    core = HeatAlertCore(store, weights, dev)
    s = index_streams(store, weights, n, rank, wl["augment"], args.locality)
    idx = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
    T = int(store.n_days.max())
    variant = {"auto": _lib.STEP_AUTO, "thread": _lib.STEP_THREAD, "coop": _lib.STEP_COOP,
               "table": _lib.STEP_TABLE}[args.variant]
    if variant == _lib.STEP_AUTO:  # what HeatAlertVectorEnv does: the logit table for big non-augmented env sets
        variant = core.choose_variant(n, canonical=not wl["augment"])
    kernel_name = {_lib.STEP_TABLE: "logit_table_kernel + step_table_kernel", _lib.STEP_COOP: "step_coop_kernel",
                   _lib.STEP_THREAD: "step_thread_kernel"}.get(
        variant, "step_coop_kernel")

    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), dtype=torch.float32, device=dev)
    reward = torch.zeros(n, dtype=torch.float32, device=dev)
    term = torch.zeros(n, dtype=torch.uint8, device=dev)
    rem = torch.zeros(n, dtype=torch.int32, device=dev)
    atb = torch.zeros(n, dtype=torch.uint8, device=dev)
    # uniform random actions, p = 0.5, one row per day of the summer, resident in HBM
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = (torch.rand((T, n), generator=g, device=dev) < 0.5).to(torch.uint8)
    step = core.bind_step(state, obs, reward, term, rem, atb, variant=variant)

    def do_reset():
        core.reset(state, obs, idx["loc"], idx["year"], idx["coef_col"], idx["coef_idx"], remaining=rem)

    log(f"[rank {rank}] setup {time.time() - t_setup:.1f}s: {n} envs, tables "
        f"{(core.x.numel() * 8 + core.w.numel() * 4) / 1e6:.0f} MB")

    # ---- device-resident throughput: `value` ---------------------------------------------------
    def run_steps(k, t):
        """k steps starting at day t of the summer; a finished summer is reset in place (included)."""
        for _ in range(k):
            step(actions[t])
            t += 1
            if t == T:
                do_reset()
                t = 0
        return t

    # Replay one CUDA graph per summer (T step launches + the reset) instead of T+1 host launches:
    # small batches are launch-latency bound and even at 2^20 envs the Python-side
    # launch loop costs a few per cent.
    # graph replays cover whole summers; any other --steps/--warmup is honoured exactly with host launches
    use_graph = (not args.no_graph) and args.steps % T == 0 and args.warmup % T == 0
    graph = None
    if use_graph:
        do_reset()
        run_steps(3, 0)  # warm every lazy path (occupancy queries, attributes) before capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for tt in range(T):
                step(actions[tt])
            do_reset()

        def run_steps(k, t):  # noqa: F811 - graph version; whole summers only, remainder by launches
            assert t == 0
            while k >= T:
                graph.replay()
                k -= T
            for i in range(k):
                step(actions[i])
            return k


    do_reset()
    t = run_steps(W, 0)
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = core.launch_count
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    ev0.record()
    t = run_steps(K, t)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    clocks.stop_flag.set()
    launches = core.launch_count - launches0
    if use_graph:
        per_step = 2 if variant == _lib.STEP_TABLE else 1  # TABLE = table build + per-env kernel
        launches += (K // T) * (T * per_step + 1)  # kernels replayed from the graph are not seen by the C-side counter
    ms = ev0.elapsed_time(ev1)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = n * world * K / (ms_max * 1e-3)

    # return statistics of one full summer (untimed), then the only collective of the path:
    # one all-reduce of the sufficient statistics (NCCL over NVLink when world > 1)
    from weather2alert_b200.distributed import RolloutStats, summarize

    rs = RolloutStats(n, dev)
    do_reset()
    for tt in range(T):
        step(actions[tt])
        rs.update(actions[tt], reward, term, atb)
    stats = summarize(rs.reduce())

    # the same summer through the episode-fused rollout kernel (one launch for all T steps, no
    # observations): an extra, not the headline metric
    ep_ret = torch.zeros(n, dtype=torch.float64, device=dev)
    ro_stats = torch.zeros(15, dtype=torch.float64, device=dev)
    ro_ms = []
    for _ in range(3):
        do_reset()
        ep_ret.zero_()
        ro_stats.zero_()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record()
        core.rollout(state, T, ep_ret, actions=actions, stats=ro_stats)
        r1.record()
        torch.cuda.synchronize()
        ro_ms.append(r0.elapsed_time(r1))
    ro_t = torch.tensor([min(ro_ms[1:])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ro_t, op=dist.ReduceOp.MAX)
    from weather2alert_b200.distributed import reduce_stats

    ro_sum = summarize(reduce_stats(ro_stats))
    rollout_info = {"value": n * world * T / (float(ro_t.item()) * 1e-3), "unit": UNIT, "ms_per_rollout": float(ro_t.item()),
                    "steps_per_rollout": T, "kernel": "rollout_kernel (episode-fused, obs elided)",
                    "mean_return": ro_sum["mean_return"],
                    "same_returns_as_stepping": bool(abs(ro_sum["mean_return"] - stats["mean_return"]) <= 1e-6 * abs(stats["mean_return"]))}

    # the agent inside the launch: a logistic policy on the observation (w2a_rollout_policy), deterministic and
    # stochastic, same summer; timed like the open-loop rollout
    theta = torch.zeros(30, dtype=torch.float64)
    theta[0], theta[6], theta[25], theta[28], theta[29] = 2.0, 0.004, 0.05, -0.3, -2.2  # heat_qi, dos, budget left, alerts in 2 weeks
    theta = theta.to(dev)
    for mode, stochastic in (("policy_deterministic", False), ("policy_stochastic", True)):
        pms = []
        for _ in range(3):
            do_reset()
            ep_ret.zero_()
            ro_stats.zero_()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            r0.record()
            core.rollout_policy(state, T, ep_ret, theta, stochastic=stochastic, seed=99, stats=ro_stats)
            r1.record()
            torch.cuda.synchronize()
            pms.append(r0.elapsed_time(r1))
        p_t = torch.tensor([min(pms[1:])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(p_t, op=dist.ReduceOp.MAX)
        p_sum = summarize(reduce_stats(ro_stats))
        rollout_info[mode] = {"value": n * world * T / (float(p_t.item()) * 1e-3), "ms_per_rollout": float(p_t.item()),
                              "mean_return": p_sum["mean_return"], "alert_freq": p_sum.get("alert_freq"),
                              "kernel": "rollout_kernel<POLICY> (logistic policy on the 29-slot observation, in the launch)"}

    # ---- policy in the loop (BASELINE configs[4] style): a small torch MLP policy on the observations and
    # the step kernel, a whole summer + reset captured in ONE CUDA graph (HeatAlertVectorEnv.graphed_rollout)
    policy_info = None
    if not args.no_e2e:
        try:
            from weather2alert_b200.vector_env import HeatAlertVectorEnv

            venv = HeatAlertVectorEnv(n, core=core, seed=1234 + rank, similar_climate_counties=wl["augment"],
                                      sort_episodes=(args.locality == "sorted"))
            venv.reset()
            gp = torch.Generator(device=dev).manual_seed(5)
            pw1 = torch.randn(29, 64, generator=gp, device=dev) * 0.3
            pw2 = torch.randn(64, generator=gp, device=dev) * 0.3

            def policy(o):
                return (torch.relu(torch.nan_to_num(o) @ pw1) @ pw2) > 0.0

            roll = venv.graphed_rollout(policy, reset_at_end=True)
            roll.replay()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            p0.record()
            for _ in range(3):
                roll.replay()
            p1.record()
            torch.cuda.synchronize()
            p_t = torch.tensor([p0.elapsed_time(p1) / 3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(p_t, op=dist.ReduceOp.MAX)
            policy_info = {"value": n * world * roll.steps / (float(p_t.item()) * 1e-3), "unit": UNIT,
                           "ms_per_summer": float(p_t.item()), "steps_per_summer": roll.steps,
                           "policy": "torch MLP 29-64-1 (fp32 matmuls) on the observation tensor, deterministic threshold",
                           "mode": "policy forward + step kernel x 153 + reset in one CUDA graph"}
            del roll, venv
        except Exception as exc:  # an extra: never let it take the headline line down
            policy_info = {"error": repr(exc)[:300]}

    # ---- end to end through the host-buffer C entry points: `e2e` ---------------------------------
    e2e = None
    if not args.no_e2e:
        core.host_envs_create(n)
        pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()  # noqa: E731
        h_obs, h_rew, h_term = pin((n, 29), torch.float32), pin(n, torch.float32), pin(n, torch.uint8)
        h_rem, h_atb = pin(n, torch.int32), pin(n, torch.uint8)
        h_act = (torch.rand((8, n)) < 0.5).to(torch.uint8).pin_memory()
        core.reset_host(s["loc"], s["year"], s["coef_col"], s["coef_idx"], h_obs, remaining=h_rem)
        ke = max(3, min(K, args.e2e_steps))
        for i in range(3):
            core.step_host(h_act[i % 8], h_obs, h_rew, h_term, h_rem, h_atb)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0 = time.perf_counter()
        for i in range(ke):
            core.step_host(h_act[i % 8], h_obs, h_rew, h_term, h_rem, h_atb)
        torch.cuda.synchronize()
        e_s = time.perf_counter() - e0
        e_t = torch.tensor([e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": n * world * ke / float(e_t.item()), "unit": UNIT, "steps": ke,
               "h2d_bytes_per_step": n * world * 1,
               "d2h_bytes_per_step": n * world * (29 * 4 + 4 + 1 + 4 + 1),
               "api": "w2a_step_host (C ABI, pinned host buffers, 3-stream chunked pipeline)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # BASELINE configs[0] on the GPU: the single-env drop-in class, one summer, random policy
    single = None
    if not args.no_e2e:
        from weather2alert_b200 import HeatAlertEnv

        env1 = HeatAlertEnv(weights=weights, store=store, device=str(dev))
        loc1 = "06037" if "06037" in weights.fips_list else weights.fips_list[0]
        env1.reset(location=loc1, seed=1234)
        arng = np.random.default_rng(7)
        for _ in range(5):
            env1.step(int(arng.integers(2)))
        env1.reset(location=loc1, seed=1234)
        s0 = time.perf_counter()
        nst, done1 = 0, False
        while not done1:
            _, _, done1, _, _ = env1.step(int(arng.integers(2)))
            nst += 1
        single = {"ms_per_step": (time.perf_counter() - s0) / nst * 1e3, "steps": nst,
                  "api": "HeatAlertEnv.reset(location, seed=1234) / .step(a): batch of one through w2a_step_host"}
        env1.close()

    # BASELINE configs[1] beside the headline workload: 4 096 envs, one CUDA graph per summer, device-resident
    small = None
    if not args.no_e2e and args.workload != "cfg2_4096" and world == 1:
        n2 = WORKLOADS["cfg2_4096"]["envs_per_gpu"]
        i2 = {k: torch.from_numpy(v).to(dev) for k, v in index_streams(store, weights, n2, 0, False, "interleaved").items()}
        st2, ob2 = core.alloc_state(n2), torch.zeros((n2, 29), dtype=torch.float32, device=dev)
        rw2, tm2 = torch.zeros(n2, dtype=torch.float32, device=dev), torch.zeros(n2, dtype=torch.uint8, device=dev)
        step2 = core.bind_step(st2, ob2, rw2, tm2)
        act2 = actions[:, :n2].contiguous()

        def summer2():
            core.reset(st2, ob2, i2["loc"], i2["year"], i2["coef_col"], i2["coef_idx"])
            for tt in range(T):
                step2(act2[tt])

        summer2()
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            summer2()
        for _ in range(3):
            g2.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g2.replay()
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / (10 * T)
        ret2 = torch.zeros(n2, dtype=torch.float64, device=dev)
        r_ms = []
        for _ in range(3):
            core.reset(st2, ob2, i2["loc"], i2["year"], i2["coef_col"], i2["coef_idx"])
            e0.record()
            core.rollout(st2, T, ret2, actions=act2)
            e1.record()
            torch.cuda.synchronize()
            r_ms.append(e0.elapsed_time(e1))
        small = {"workload": "cfg2_4096", "value": n2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2,
                 "rollout_value": n2 * T / (min(r_ms[1:]) * 1e-3),
                 "note": "launch-latency bound (0.6 MB per step); the episode-fused rollout kernel is the fast path at this size"}

    peak, peak_src = load_peaks()
    sec_per_step = ms_max * 1e-3 / K
    ach = BYTES_STREAMING * n / sec_per_step / 1e9
    ach_c = BYTES_CANONICAL * n / sec_per_step / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "envs_per_gpu": n, "envs_total": n * world,
                   "counties": S_COUNTIES, "years": N_YEARS, "days": T, "posterior_samples": P_SAMPLES,
                   "weights": "linear-format synthetic (nn_* checkpoints share the layout and the path)",
                   "augment": wl["augment"],
                   "mode": "one CUDA graph per summer (153 steps + reset)" if use_graph else "per-step launches",
                   "locality": args.locality, "variant": args.variant, "kernel": kernel_name,
                   "parallelism": f"env-sharded x{world}, no step-time collective",
                   "l2": "per-step streaming traffic (151 B x envs) exceeds the 126 MB L2 for >= 2^20 envs; "
                         "tables (feature rows of the day + coefficients) are meant to stay L2-resident"},
        "gpu_launches": int(launches),
        "wall_s": w1 - w0,
        "clocks": clocks.summary(w0, w1),
        "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "bytes_per_env_step": BYTES_STREAMING, "peak_source": peak_src,
                     "kernel": kernel_name,
                     "note": "achieved = streaming bytes (HBM-mandatory) x envs / mean step time over the timed region"},
        "roofline_canonical_466B": {"achieved": ach_c, "frac": ach_c / peak,
                                    "note": "SURVEY 8(d) canonical 466 B/env-step; includes table gathers served from L2"},
        "return_stats": {**stats, "reduced_over": world},
        "rollout": rollout_info,
        "single_env": single,
        "configs_1_4096_envs": small,
        "policy_in_loop": policy_info,
    }
    traffic_file = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                tr = json.load(fh)
            if tr.get("workload") == args.workload and tr.get("envs") == n:
                out["roofline"]["traffic"] = tr["dram_bytes_per_launch"]
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_port(store, weights, args.cpu_seconds)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


# =================================================================================================
# CPU baselines
# =================================================================================================
def cpu_baseline_port(store, weights, budget_s=12.0):
    """oracle/c/heat_oracle.c (the reference algorithm restated in C, OpenMP over envs) on the host
    cores, on a bounded sample of the same workload."""
    from oracle import c_oracle

    if not c_oracle.available():
        import __graft_entry__ as g

        g.build_oracle()
    n = 1 << 16
    rng = np.random.default_rng(99)
    S = len(weights.fips_list)
    loc = (np.arange(n) % S).astype(np.int32)
    co = c_oracle.COracle(store.feat, store.sig_code, store.n_days, weights.coef, n)
    co.reset(loc, rng.integers(0, store.shape[1], n), loc, rng.integers(0, weights.n_samples, n))
    acts = (rng.random((152, n)) < 0.5).astype(np.uint8)
    for i in range(3):
        co.step(acts[i])
    t0, steps = time.perf_counter(), 0
    while steps < 149 and time.perf_counter() - t0 < budget_s:
        co.step(acts[3 + steps])
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "cores": co.threads, "kind": "port",
            "sample": f"{n} envs x {steps} steps of the same workload through oracle/c/heat_oracle.c (OpenMP)"}


_REF_BASE_ENV = None  # built once in the parent; forked workers share its frames copy-on-write


def _ref_worker(conn, seeds, locations):
    """One forked worker: a few reference envs, stepped on command."""
    import copy
    import warnings

    warnings.simplefilter("ignore")
    envs = []
    for sd, loc in zip(seeds, locations):
        env = copy.copy(_REF_BASE_ENV)  # shares the (read-only) merged frame and coefficient tensors
        env.reset(location=loc, seed=int(sd))
        envs.append(env)
    rng = np.random.default_rng(seeds[0])
    conn.send("ready")
    while True:
        k = conn.recv()
        if k is None:
            break
        for _ in range(k):
            for env in envs:
                _, _, done, _, _ = env.step(int(rng.integers(2)))
                if done:
                    env.reset(seed=int(rng.integers(10000)))
        conn.send(k * len(envs))


def run_reference(args):
    """The reference's own CPU implementation on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    import tempfile

    from oracle import ref_harness

    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    if ref_harness.reference_location() is None:
        # no install of the reference travelled here: fall back to the C port of its algorithm
        store, weights = make_tables()
        cb = cpu_baseline_port(store, weights, 20.0)
        out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
               "steps": K, "warmup": W, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"]},
               "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(out)
        return

    from weather2alert_b200 import synthetic

    # Same county list / checkpoint shape as our arm; 3 years of weather instead of 11 keeps
    # the pandas frames of every forked worker small -- the reference's step cost does not
    # depend on the table size (it works on one episode frame).
    t0 = time.time()
    fips, zones = synthetic.synthetic_counties(S_COUNTIES, seed=0)
    store = synthetic.make_store(fips, [2006, 2007, 2008], seed=1, zones=zones)
    weights = synthetic.make_weights(fips, P_SAMPLES, seed=2)
    fixture = tempfile.mkdtemp(prefix="w2a_ref_fixture_")
    synthetic.write_reference_files(store, weights, fixture)
    log(f"reference fixture written in {time.time() - t0:.1f}s")
    global _REF_BASE_ENV
    _REF_BASE_ENV = ref_harness.make_reference_env(fixture)
    # SURVEY 8(d) (i) one reference env, (ii) a serial loop over 8 reference envs in this process
    import copy

    arng = np.random.default_rng(7)
    one = copy.copy(_REF_BASE_ENV)
    loc1 = "06037" if "06037" in fips else fips[0]
    one.reset(location=loc1, seed=1234)
    t1 = time.perf_counter()
    n1 = 0
    for _ in range(min(100, max(10, K))):
        _, _, d1, _, _ = one.step(int(arng.integers(2)))
        n1 += 1
        if d1:
            one.reset(location=loc1, seed=1234)
    single_ms = (time.perf_counter() - t1) / n1 * 1e3
    eight = [copy.copy(_REF_BASE_ENV) for _ in range(8)]
    for j, e8 in enumerate(eight):
        e8.reset(location=fips[j % len(fips)], seed=100 + j)
    t8 = time.perf_counter()
    for _ in range(20):
        for e8 in eight:
            e8.step(int(arng.integers(2)))
    sync8 = 8 * 20 / (time.perf_counter() - t8)
    del one, eight
    n_workers = max(1, min(cores, args.ref_workers or cores))
    envs_per_worker = 2
    ctx = mp.get_context("fork")
    workers = []
    for w in range(n_workers):
        parent, child = ctx.Pipe()
        seeds = [1234 + w * envs_per_worker + j for j in range(envs_per_worker)]
        locs = [fips[(w * envs_per_worker + j) % len(fips)] for j in range(envs_per_worker)]
        p = ctx.Process(target=_ref_worker, args=(child, seeds, locs), daemon=True)
        p.start()
        workers.append((p, parent))
    for _, c in workers:
        assert c.recv() == "ready"
    log(f"{n_workers} reference workers x {envs_per_worker} envs ready in {time.time() - t0:.1f}s")

    def run(k):
        for _, c in workers:
            c.send(k)
        return sum(c.recv() for _, c in workers)

    run(max(1, W))
    t0 = time.perf_counter()
    total = run(K)
    dt = time.perf_counter() - t0
    for p, c in workers:
        c.send(None)
    value = total / dt
    sample = (f"{n_workers} forked workers x {envs_per_worker} unmodified reference envs (baseline/_ref, gymnasium stub, "
              f"local fixture {S_COUNTIES} counties x 3 years), {K} steps each, random actions")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
           "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "desc": wl["desc"], "envs_total": n_workers * envs_per_worker},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_workers, "kind": "reference", "sample": sample},
           "reference_single_env_ms_per_step": single_ms, "reference_sync_loop_8_envs_steps_per_s": sync8,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


_REAL_STDOUT = None


def protect_stdout():
    """Library banners (e.g. NCCL's version line) go to stderr; fd 1 carries only the JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3060)
    ap.add_argument("--warmup", type=int, default=153)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=list(WORKLOADS), default="cfg3_1m")
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--variant", choices=["auto", "thread", "coop", "table"], default="auto")
    ap.add_argument("--locality", choices=["interleaved", "sorted"], default="interleaved")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from the host instead of replaying a CUDA graph per summer")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-workers", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 200:
            args.steps = 200  # 3 ms per reference env-step: keep the arm to a couple of minutes
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
