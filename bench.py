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


NN_FIPS_FILE = os.path.join(REPO, "weather2alert_b200", "data", "nn_full_medicare_all_fips.txt")


def make_tables(years=N_YEARS, counties=S_COUNTIES, samples=P_SAMPLES, layout="linear"):
    """layout="linear": 746 synthetic counties, weights for all of them (weights/linear: S = 746).
    layout="nn": the store keeps its 746 counties but the checkpoint carries the 720-county `fips_list`
    of the reference's weights/nn_full_medicare_all/config.yaml (BASELINE configs[2]): coefficient column
    != store county, and 26 store counties have no column."""
    from weather2alert_b200 import synthetic

    if layout == "nn":
        with open(NN_FIPS_FILE) as fh:
            nn_fips = fh.read().split()
        fips, zones = synthetic.synthetic_counties(counties, seed=0, must_include=tuple(nn_fips))
        store = synthetic.make_store(fips, list(range(2006, 2006 + years)), seed=1, zones=zones)
        return store, synthetic.make_weights(nn_fips, samples, seed=2)
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
    list_idx = (np.arange(n, dtype=np.int64) + rank * n) % S  # index into the weights' fips_list
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
class EnvSet:
    """One device-resident env set on one GPU: buffers, the bound step call, and CUDA graphs of step
    segments so that ANY --steps/--warmup runs from graphs: a segment is `l` consecutive steps that do
    not cross the end of the summer, optionally followed by the reset of all envs; one graph per
    distinct (l, reset) is captured before the timed region and replayed inside it."""

    def __init__(self, core, idx, n, T, actions, variant, use_graph=True):
        import torch

        self.torch, self.core, self.idx, self.n, self.T, self.actions = torch, core, idx, n, T, actions
        dev = core.device
        self.state = core.alloc_state(n)
        self.obs = torch.zeros((n, 29), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.term = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.rem = torch.zeros(n, dtype=torch.int32, device=dev)
        self.atb = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.step = core.bind_step(self.state, self.obs, self.reward, self.term, self.rem, self.atb, variant=variant)
        self.use_graph, self.graphs, self.t, self.kernels = use_graph, {}, 0, 0

    def reset(self):
        i = self.idx
        self.core.reset(self.state, self.obs, i["loc"], i["year"], i["coef_col"], i["coef_idx"], remaining=self.rem)
        self.t = 0

    def segments(self, t, k):
        """(length, reset_after) pieces of k steps starting on day t; a finished summer is reset in place."""
        out = []
        while k > 0:
            seg = min(k, self.T - t)
            t += seg
            k -= seg
            out.append((seg, t == self.T))
            if t == self.T:
                t = 0
        return out, t

    def _launch(self, seg, reset_after):
        for i in range(seg):
            self.step(self.actions[i])
        if reset_after:
            self.reset()

    def prepare(self, runs):
        """Capture every graph the given sequence of run(k) calls will need, starting from the current day."""
        if not self.use_graph:
            return
        torch, t = self.torch, self.t
        need = []
        for k in runs:
            segs, t = self.segments(t, k)
            need += [sg for sg in segs if sg not in self.graphs and sg not in need]
        t_keep = self.t
        for sg in need:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch(*sg)
            self.graphs[sg] = g
        self.t = t_keep  # a capture executes nothing

    def run(self, k):
        segs, t_end = self.segments(self.t, k)
        for sg in segs:
            g = self.graphs.get(sg)
            if g is not None:
                g.replay()
            else:
                self._launch(*sg)
            self.kernels += sg[0] + (1 if sg[1] else 0)
        self.t = t_end


def parity_gate(core, store, weights, variant, n=1 << 16, steps=12, seed=4321):
    """before anything is timed, the tables the bench runs on (746 counties x 11 years
    x 100 samples) are stepped on the GPU and by the C oracle (oracle/c/heat_oracle.c, the checker) from the
    same index and action streams: observations, flags and counters bit-exact, rewards within 1e-5 relative."""
    import torch

    from oracle import c_oracle

    if not c_oracle.available():
        import __graft_entry__ as g

        g.build_oracle()
    dev = core.device
    rng = np.random.default_rng(seed)
    S = len(weights.fips_list)
    col = (np.arange(n) % S).astype(np.int32)
    loc = np.array([store.fips_index[f] for f in weights.fips_list], np.int32)[col]
    year = rng.integers(0, store.shape[1], n).astype(np.int32)
    p = rng.integers(0, weights.n_samples, n).astype(np.int32)
    co = c_oracle.COracle(store.feat, store.sig_code, store.n_days, weights.coef, n)
    o_ref = co.reset(loc, year, col, p)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    state = core.alloc_state(n)
    obs = torch.zeros((n, 29), dtype=torch.float32, device=dev)
    reward, term = torch.zeros(n, dtype=torch.float32, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    rem, atb = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
    core.reset(state, obs, d(loc), d(year), d(col), d(p), remaining=rem)

    def same(a, b):  # bit-for-bit on values, NaN == NaN
        return bool(np.array_equal(a, b, equal_nan=True))

    obs_ok = same(obs.cpu().numpy(), o_ref.astype(np.float32))
    flags_ok, max_rel = True, 0.0
    for i in range(steps):
        act = (rng.random(n) < (0.1, 0.5, 0.9)[i % 3]).astype(np.uint8)
        core.step(state, d(act), obs, reward, term, rem, atb, variant=variant)
        o_ref, r_ref, d_ref, rem_ref, atb_ref = co.step(act)
        obs_ok &= same(obs.cpu().numpy(), o_ref.astype(np.float32))
        flags_ok &= same(term.cpu().numpy(), d_ref) and same(rem.cpu().numpy(), rem_ref) and same(atb.cpu().numpy(), atb_ref)
        r = reward.cpu().numpy().astype(np.float64)
        err = np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1e-5)  # 1e-10 absolute floor of the 1e-5 relative bar
        max_rel = max(max_rel, float(err.max()))
    ok = bool(obs_ok and flags_ok and max_rel <= 1e-5)
    return {"ok": ok, "envs": n, "steps": steps, "obs_bit_exact": bool(obs_ok), "flags_and_counters_exact": bool(flags_ok),
            "max_reward_rel": max_rel, "reward_rtol": 1e-5,
            "checker": "oracle/c/heat_oracle.c on the bench's own tables, same index and action streams"}


def timed_steps(es, K, W, world, dist, target_ms=50.0, max_repeats=400):
    """W warm-up steps, then R back-to-back regions of EXACTLY K steps, each bracketed by CUDA events on the
    launching stream; barrier + synchronize on both sides of the whole sequence.  Returns the per-region times."""
    import torch

    dev = es.core.device
    es.reset()
    es.prepare([W, K])
    es.run(W)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    p0.record()
    es.run(K)  # pilot region (also warm-up): sizes the number of repeats
    p1.record()
    torch.cuda.synchronize()
    reps = int(min(max_repeats, max(1, -(-target_ms // max(p0.elapsed_time(p1), 1e-3)))))
    r_t = torch.tensor([reps], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
    reps = int(r_t.item())
    es.prepare([K] * reps)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    k0 = es.kernels
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    ev[0].record()
    for r in range(reps):
        es.run(K)
        ev[r + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    times = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
    return times, (es.kernels - k0) // reps, w0, w1


def run_ours(args):
    import torch
    import torch.distributed as dist

    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore
    from weather2alert_b200.distributed import RolloutStats, reduce_stats, summarize

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

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = WORKLOADS[args.workload]
    n = args.envs_per_gpu or wl["envs_per_gpu"]
    K, W = args.steps, args.warmup
    t_setup = time.time()
    store, weights = make_tables(layout=args.weights_layout)
    core = HeatAlertCore(store, weights, dev)
    torch.cuda.synchronize()
    s = index_streams(store, weights, n, rank, wl["augment"], args.locality)
    idx = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
    T = int(store.n_days.max())
    names = {"auto": _lib.STEP_AUTO, "thread": _lib.STEP_THREAD, "coop": _lib.STEP_COOP, "table": _lib.STEP_TABLE}

    def resolve(n_envs, augment):  # what HeatAlertVectorEnv does above the C ABI
        v = names[args.variant]
        return core.choose_variant(n_envs, canonical=not augment) if v == _lib.STEP_AUTO else v

    variant = resolve(n, wl["augment"])
    kernel_names = {_lib.STEP_TABLE: "step_table_kernel", _lib.STEP_COOP: "step_coop_kernel",
                    _lib.STEP_THREAD: "step_thread_kernel", _lib.STEP_AUTO: "step_coop_kernel"}
    kernel_name = kernel_names[variant]
    log(f"[rank {rank}] setup {time.time() - t_setup:.1f}s: {n} envs, tables "
        f"{(core.x.numel() * 8 + core.w.numel() * 4) / 1e6:.0f} MB + {core.table_bytes / 1e9:.2f} GB precomputed "
        f"(has_tables={core.has_tables}), variant {kernel_name}")

    # ---- parity gate: nothing is timed unless the path agrees with the oracle on these very tables ----
    parity = None
    if not args.no_parity:
        parity = parity_gate(core, store, weights, variant)
        log(f"[rank {rank}] parity gate: {parity}")
        ok_t = all_max(0.0 if parity["ok"] else 1.0)
        if ok_t != 0.0:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "error": "parity gate failed",
                      "parity": parity})
            raise SystemExit(3)

    # uniform random actions, p = 0.5, one row per day of the summer, resident in HBM
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = (torch.rand((T, n), generator=g, device=dev) < 0.5).to(torch.uint8)

    # ---- device-resident throughput: `value` ---------------------------------------------------
    # env slots in table order: tell the step (the hint HeatAlertVectorEnv passes for its locality-sorted resets)
    local_hint = _lib.STEP_LOCAL if (args.locality == "sorted" and variant == _lib.STEP_TABLE) else 0
    es = EnvSet(core, idx, n, T, actions, variant | local_hint, use_graph=not args.no_graph)
    es.reset()
    es.run(3)  # warm every lazy path (occupancy queries, attributes) before any capture
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    times, launches, w0, w1 = timed_steps(es, K, W, world, dist)
    clocks.stop_flag.set()
    ms_med = all_max(statistics.median(times))
    value = n * world * K / (ms_med * 1e-3)
    timing = {"repeats": len(times), "ms_region_median": ms_med, "ms_region_min": all_max(min(times)),
              "ms_region_max": all_max(max(times)), "timed_ms_total": float(sum(times)),
              "note": "R back-to-back regions of exactly K steps, one CUDA-event pair each; value and ms_per_step use "
                      "the median region (max over ranks)"}

    # ---- the same episodes in (county, sample, year) slot order (`sort_episodes=True` of the vector env) ----
    sorted_info = None
    if args.locality == "interleaved" and not args.no_extras:
        s2 = index_streams(store, weights, n, rank, wl["augment"], "sorted")
        es2 = EnvSet(core, {k: torch.from_numpy(v).to(dev) for k, v in s2.items()}, n, T, actions,
                     variant | (_lib.STEP_LOCAL if variant == _lib.STEP_TABLE else 0), use_graph=not args.no_graph)
        t2, _, _, _ = timed_steps(es2, K, W, world, dist)
        ms2 = all_max(statistics.median(t2))
        sorted_info = {"value": n * world * K / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / K,
                       "roofline_frac": BYTES_STREAMING * n / (ms2 * 1e-3 / K) / 1e9 / load_peaks()[0],
                       "note": "same multiset of episodes, env slots in (county, sample, year) order -- what "
                               "HeatAlertVectorEnv(sort_episodes=True), its default for device-sampled resets, produces"}
        del es2

    # ---- HeatAlertVectorEnv with its defaults (device-sampled episodes, sort_episodes on), plain .step() calls ----
    venv_info = None
    if not args.no_extras:
        from weather2alert_b200.vector_env import HeatAlertVectorEnv

        venv = HeatAlertVectorEnv(n, core=core, seed=4321 + rank, similar_climate_counties=wl["augment"])
        venv.reset()
        for tt in range(T + 5):  # a whole warm-up summer: the first autoreset of a process pays one-off allocator growth
            venv.step(actions[tt % T])
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        v0.record()
        for tt in range(T):  # crosses the end of the summer: the autoreset (draw + sort + reset kernel) is inside
            venv.step(actions[(tt + 5) % T])
        v1.record()
        torch.cuda.synchronize()
        vms = all_max(v0.elapsed_time(v1)) / T
        venv_info = {"value": n * world / (vms * 1e-3), "unit": UNIT, "ms_per_step": vms,
                     "roofline_frac": BYTES_STREAMING * n / (vms * 1e-3) / 1e9 / load_peaks()[0],
                     "kernel": kernel_names[venv._variant() & 0xff], "sort_episodes": venv.sort_episodes,
                     "note": "HeatAlertVectorEnv(num_envs, core=...) with default arguments: .step(actions) called from "
                             "Python 153 times (no CUDA graph), the autoreset of all envs included"}
        del venv

    stats, rollout_info, theta = {}, None, None
    if not args.no_extras:
        # return statistics of one full summer (untimed), then the only collective of the path:
        # one all-reduce of the sufficient statistics (NCCL over NVLink when world > 1)
        rs = RolloutStats(n, dev)
        es.reset()
        for tt in range(T):
            es.step(actions[tt])
            rs.update(actions[tt], es.reward, es.term, es.atb)
        stats = summarize(rs.reduce())
        if int(round(stats["episodes"])) != n * world:
            raise SystemExit(f"return statistics: {stats['episodes']} episodes reduced, expected {n * world}")
        state, reward, term, atb = es.state, es.reward, es.term, es.atb

        def do_reset():
            es.reset()

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
        ro_t = all_max(min(ro_ms[1:]))
        ro_sum = summarize(reduce_stats(ro_stats))
        rollout_info = {"value": n * world * T / (ro_t * 1e-3), "unit": UNIT, "ms_per_rollout": ro_t,
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
            p_t = all_max(min(pms[1:]))
            p_sum = summarize(reduce_stats(ro_stats))
            rollout_info[mode] = {"value": n * world * T / (p_t * 1e-3), "ms_per_rollout": p_t,
                                  "mean_return": p_sum["mean_return"], "alert_freq": p_sum.get("alert_freq"),
                                  "kernel": "rollout_kernel<POLICY> (logistic policy on the 29-slot observation, in the launch)"}

    # ---- policy in the loop (BASELINE configs[4] style): a small torch MLP policy on the observations and
    # the step kernel, a whole summer + reset captured in ONE CUDA graph (HeatAlertVectorEnv.graphed_rollout)
    policy_info = None
    if not args.no_extras:
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
            p_t = all_max(p0.elapsed_time(p1) / 3)
            policy_info = {"value": n * world * roll.steps / (p_t * 1e-3), "unit": UNIT,
                           "ms_per_summer": p_t, "steps_per_summer": roll.steps,
                           "policy": "torch MLP 29-64-1 (fp32 matmuls) on the observation tensor, deterministic threshold",
                           "mode": "policy forward + step kernel x 153 + reset in one CUDA graph"}
            del roll
            # the same network as ONE fused kernel per step (w2a_policy_mlp: tensor cores, float16 inputs, float32
            # accumulate), then with the step and the policy in ONE launch per iteration (w2a_step_mlp)
            for key, fuse, what, mode in (
                    ("fused_mlp_policy", False,
                     "the same 29-64-1 ReLU MLP as one w2a_policy_mlp launch per step (mma.sync f16 inputs, fp32 accumulate)",
                     "policy kernel + step kernel x 153 + reset in one CUDA graph"),
                    ("fused_step_and_policy", True,
                     "the same MLP evaluated inside the step launch on the fresh observation rows (w2a_step_mlp)",
                     "ONE kernel per iteration x 153 + reset in one CUDA graph; observations written, never read back")):
                venv.reset()
                fused = venv.mlp_policy(pw1.T.contiguous(), torch.zeros(64, device=dev), pw2, 0.0)
                fused.fuse_step = fuse
                roll = venv.graphed_rollout(fused, reset_at_end=True)
                roll.replay()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                p0.record()
                for _ in range(3):
                    roll.replay()
                p1.record()
                torch.cuda.synchronize()
                f_t = all_max(p0.elapsed_time(p1) / 3)
                policy_info[key] = {
                    "value": n * world * roll.steps / (f_t * 1e-3), "unit": UNIT, "ms_per_summer": f_t,
                    "us_per_step": f_t / roll.steps * 1e3, "policy": what, "mode": mode}
                del roll
            del venv
        except Exception as exc:  # an extra: never let it take the headline line down
            policy_info = {"error": repr(exc)[:300]}

    # ---- BASELINE configs[3] and configs[4] per GPU beside the headline (SCALE then shows them at 8 GPUs):
    # cfg4_2m = 2^21 augmented envs stepped; cfg5_8m = 2^23 envs, a full summer with the logistic policy inside
    # the launch and the statistics all-reduced (the "policy-in-the-loop rollout with NCCL return-stat reduction")
    other = {}
    if not args.no_extras and not args.no_other_configs:
        for name in ("cfg4_2m", "cfg5_8m"):
            if name == args.workload:
                continue
            try:
                wo = WORKLOADS[name]
                no = wo["envs_per_gpu"]
                so = index_streams(store, weights, no, rank, wo["augment"], "interleaved")
                io = {k: torch.from_numpy(v).to(dev) for k, v in so.items()}
                go = torch.Generator(device=dev).manual_seed(77 + rank)
                ao = (torch.rand((T, no), generator=go, device=dev) < 0.5).to(torch.uint8)
                vo = resolve(no, wo["augment"])
                eo = EnvSet(core, io, no, T, ao, vo, use_graph=not args.no_graph)
                eo.reset()
                eo.run(3)
                to, _, _, _ = timed_steps(eo, K, W, world, dist, target_ms=30.0)
                mo = all_max(statistics.median(to))
                entry = {"value": no * world * K / (mo * 1e-3), "unit": UNIT, "ms_per_step": mo / K, "envs_per_gpu": no,
                         "envs_total": no * world, "kernel": kernel_names[vo], "augment": wo["augment"],
                         "roofline_frac": BYTES_STREAMING * no / (mo * 1e-3 / K) / 1e9 / load_peaks()[0], "desc": wo["desc"]}
                if name == "cfg5_8m":
                    ret_o = torch.zeros(no, dtype=torch.float64, device=dev)
                    st_o = torch.zeros(15, dtype=torch.float64, device=dev)
                    pm = []
                    for _ in range(2):
                        eo.reset()
                        ret_o.zero_()
                        st_o.zero_()
                        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        if world > 1:
                            dist.barrier()
                        r0.record()
                        core.rollout_policy(eo.state, T, ret_o, theta, stochastic=True, seed=99, stats=st_o)
                        red = reduce_stats(st_o)  # the NCCL all-reduce of the 15 sufficient statistics, inside the timed region
                        r1.record()
                        torch.cuda.synchronize()
                        pm.append(r0.elapsed_time(r1))
                    pt = all_max(min(pm))
                    ps = summarize(red)
                    if int(round(ps["episodes"])) != no * world:
                        raise RuntimeError(f"policy rollout reduced {ps['episodes']} episodes, expected {no * world}")
                    entry["policy_rollout_with_stat_allreduce"] = {
                        "value": no * world * T / (pt * 1e-3), "unit": UNIT, "ms_per_summer": pt, "episodes": ps["episodes"],
                        "mean_return": ps["mean_return"], "alert_freq": ps.get("alert_freq"),
                        "kernel": "rollout_kernel<POLICY> (stochastic logistic policy) + all_reduce(float64[15])"}
                other[name] = entry
                del eo, io, ao
                torch.cuda.empty_cache()
            except Exception as exc:  # extras never take the headline down
                other[name] = {"error": repr(exc)[:300]}

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
        e_s = all_max(time.perf_counter() - e0)
        d2h = 29 * 4 + 4 + 1 + 4 + 1
        # the link's own ceiling, measured here: one plain pinned device->host copy of the same byte count per step,
        # all ranks at once (tools/pcie_bench.py sweeps 1..8 concurrent ranks)
        blk = torch.empty(n * d2h, dtype=torch.uint8, device=dev)
        hblk = torch.empty(n * d2h, dtype=torch.uint8).pin_memory()
        for _ in range(2):
            hblk.copy_(blk, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = time.perf_counter()
        for _ in range(8):
            hblk.copy_(blk, non_blocking=True)
        torch.cuda.synchronize()
        l_s = all_max(time.perf_counter() - l0)
        link_gbs = n * d2h * 8 / l_s / 1e9
        e2e = {"value": n * world * ke / e_s, "unit": UNIT, "steps": ke,
               "h2d_bytes_per_step": n * world * 1,
               "d2h_bytes_per_step": n * world * d2h,
               "gbs_per_gpu": n * (d2h + 1) * ke / e_s / 1e9,
               "link_gbs": link_gbs, "frac_of_link": n * (d2h + 1) * ke / e_s / 1e9 / link_gbs,
               "link_note": f"link_gbs = plain pinned cudaMemcpyAsync D2H of {n * d2h >> 20} MiB blocks per GPU with all {world} ranks copying at once",
               "api": "w2a_step_host (C ABI, pinned host buffers, 3-stream chunked pipeline)"}
        del blk, hblk

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # BASELINE configs[0] on the GPU: the single-env drop-in class, one summer, random policy
    single = None
    if not args.no_e2e and not args.no_extras:
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
    if not args.no_extras and args.workload != "cfg2_4096" and world == 1:
        n2 = WORKLOADS["cfg2_4096"]["envs_per_gpu"]
        i2 = {k: torch.from_numpy(v).to(dev) for k, v in index_streams(store, weights, n2, 0, False, "interleaved").items()}
        small = {"workload": "cfg2_4096", "unit": UNIT}
        for vname, v2 in (("table", _lib.STEP_TABLE), ("coop", _lib.STEP_COOP)):
            e2 = EnvSet(core, i2, n2, T, actions[:, :n2].contiguous(), v2)
            e2.reset()
            e2.run(3)
            tm, _, _, _ = timed_steps(e2, T, T, 1, dist, target_ms=10.0)
            small[f"ms_per_step_{vname}"] = statistics.median(tm) / T
        v_auto = resolve(n2, False)
        ms2 = small["ms_per_step_table" if v_auto == _lib.STEP_TABLE else "ms_per_step_coop"]
        st2 = core.alloc_state(n2)
        ob2 = torch.zeros((n2, 29), dtype=torch.float32, device=dev)
        act2 = actions[:, :n2].contiguous()
        ret2 = torch.zeros(n2, dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r_ms = []
        for _ in range(3):
            core.reset(st2, ob2, i2["loc"], i2["year"], i2["coef_col"], i2["coef_idx"])
            e0.record()
            core.rollout(st2, T, ret2, actions=act2)
            e1.record()
            torch.cuda.synchronize()
            r_ms.append(e0.elapsed_time(e1))
        small.update({"value": n2 / (ms2 * 1e-3), "ms_per_step": ms2, "kernel": kernel_names[v_auto],
                      "rollout_value": n2 * T / (min(r_ms[1:]) * 1e-3),
                      "note": "launch-latency bound (0.6 MB per step); the episode-fused rollout kernel is the fast path at this size"})

    peak, peak_src = load_peaks()
    sec_per_step = ms_med * 1e-3 / K
    ach = BYTES_STREAMING * n / sec_per_step / 1e9
    ach_c = BYTES_CANONICAL * n / sec_per_step / 1e9
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "envs_per_gpu": n, "envs_total": n * world,
                   "counties": len(weights.fips_list), "store_counties": store.shape[0], "years": N_YEARS, "days": T,
                   "posterior_samples": P_SAMPLES,
                   "weights": ("nn_full_medicare_all layout: the 720-county fips_list of the reference's config.yaml, synthetic "
                               "coefficients" if args.weights_layout == "nn" else
                               "linear-format synthetic (nn_* checkpoints share the layout and the path)"),
                   "augment": wl["augment"],
                   "mode": ("CUDA graphs of step segments (any --steps/--warmup), summer reset included"
                            if not args.no_graph else "per-step launches"),
                   "locality": args.locality, "variant": args.variant, "kernel": kernel_name,
                   "tables": f"whole-summer logit tables precomputed once ({core.table_bytes / 1e9:.2f} GB in HBM), no per-step build",
                   "parallelism": f"env-sharded x{world}, no step-time collective",
                   "l2": "per-step streaming traffic (151 B x envs) exceeds the 126 MB L2 for >= 2^20 envs: inputs larger than L2"},
        "gpu_launches": int(launches),
        "timing": timing,
        "wall_s": w1 - w0,
        "clocks": clocks.summary(w0, w1),
        "parity": parity,
        "e2e": e2e,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "bytes_per_env_step": BYTES_STREAMING, "peak_source": peak_src,
                     "kernel": kernel_name,
                     "note": "achieved = streaming bytes (HBM-mandatory) x envs / median step time over the timed regions"},
        "roofline_canonical_466B": {"achieved": ach_c, "frac": ach_c / peak,
                                    "note": "SURVEY 8(d) canonical 466 B/env-step; includes table gathers served from L2"},
        "sorted": sorted_info,
        "vector_env_default": venv_info,
        "return_stats": {**stats, "reduced_over": world},
        "rollout": rollout_info,
        "single_env": single,
        "configs_1_4096_envs": small,
        "policy_in_loop": policy_info,
        "other_configs": other,
    }
    traffic_file = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                tr = json.load(fh)
            if tr.get("workload") == args.workload and tr.get("envs") == n and tr.get("kernel") == kernel_name:
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
    ap.add_argument("--no-extras", action="store_true", help="headline, parity, e2e and cpu_baseline only")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the cfg4_2m / cfg5_8m extras")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate (profiling runs only)")
    ap.add_argument("--weights-layout", choices=["linear", "nn"], default="linear",
                    help="nn: the 720-county fips_list of weights/nn_full_medicare_all against the 746-county store")
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
