"""Device-side core: HBM-resident tables + the reset/step launches, on torch tensors.

PyTorch is plumbing here (device memory, the current stream); the work is done by the
sm_100a kernels behind the C ABI.  Mirrors the data-holding half of
`HeatAlertEnv.__init__` (`/root/reference/src/weather2alert/env.py:20-105`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .store import FeatureStore, RewardWeights

_ACT_DTYPES = {torch.uint8: _lib.ACT_U8, torch.bool: _lib.ACT_U8, torch.int32: _lib.ACT_I32, torch.int64: _lib.ACT_I64}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class HeatAlertCore:
    """One C context on one GPU with the feature store and coefficient tables installed."""

    def __init__(self, store: FeatureStore, weights: RewardWeights, device="cuda:0", precompute: bool = True):
        self._lib = _lib.lib()  # raises ImportError when the extension is not built
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("HeatAlertCore needs a CUDA device; there is no CPU path")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available; weather2alert_b200 has no CPU fallback")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        self.store, self.weights = store, weights
        S, Y, T = store.shape
        self.S, self.Y, self.T = S, Y, T
        self.P, self.Sw = weights.n_samples, len(weights.fips_list)
        dev = self.device
        self.x = torch.from_numpy(store.pack_x()).to(dev)
        self.n_days = torch.from_numpy(np.ascontiguousarray(store.n_days.reshape(-1))).to(dev)
        self.budget0 = torch.from_numpy(np.ascontiguousarray(store.budget0.reshape(-1))).to(dev)
        self.w = torch.from_numpy(weights.pack()).to(dev)
        ctx = C.c_void_p()
        _lib.check(self._lib.w2a_create(index, C.byref(ctx)))
        self.ctx = ctx
        _lib.check(self._lib.w2a_set_features(ctx, _ptr(self.x), _ptr(self.n_days), _ptr(self.budget0),
                                              S, Y, T), ctx)
        _lib.check(self._lib.w2a_set_coefs(ctx, _ptr(self.w), self.P, self.Sw), ctx)
        # feature-store county of every coefficient column (the reference keys both by FIPS string,
        # env.py:75,121,127); -1 = the weights know a county the store does not (e.g. a 720-county
        # `nn_*` fips_list against a bigger store, or the other way round)
        col_to_loc = np.array([store.fips_index.get(f, -1) for f in weights.fips_list], dtype=np.int32)
        self.col_to_loc = torch.from_numpy(col_to_loc).to(dev)
        _lib.check(self._lib.w2a_set_county_map(ctx, _ptr(self.col_to_loc)), ctx)
        # whole-summer logit tables (`w2a_precompute_logits`): the agent-independent part of
        # `_get_reward` (env.py:206-217) evaluated once here instead of in every step; enables STEP_TABLE
        self.has_tables = False
        self.table_bytes = int(self._lib.w2a_precompute_bytes(ctx))
        if precompute:
            self.precompute()

    def precompute(self) -> bool:
        """Build the whole-summer tables; False (and STEP_TABLE stays unavailable) when the device has no room."""
        rc = self._lib.w2a_precompute_logits(self.ctx, self._stream())
        if rc == -4:  # W2A_E_NOMEM: nothing was allocated, the other variants still work
            self.has_tables = False
            return False
        _lib.check(rc, self.ctx)
        self.has_tables = True
        return True

    def __del__(self):
        ctx = getattr(self, "ctx", None)
        if ctx:
            try:
                self._lib.w2a_destroy(ctx)
            except Exception:
                pass
            self.ctx = None

    # ---- buffers ---------------------------------------------------------------------
    def alloc_state(self, n: int) -> torch.Tensor:
        return torch.zeros(int(self._lib.w2a_state_bytes(n)), dtype=torch.uint8, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_dev(self, name, t, dtype=None, numel=None):
        if t is None:
            return
        if not t.is_cuda or t.device != self.device or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous tensor on {self.device}")
        if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
            raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
        if numel is not None and t.numel() != numel:
            raise ValueError(f"{name} must have {numel} elements, got {t.numel()}")

    # ---- launches --------------------------------------------------------------------
    def reset(self, state, obs, loc, year, coef_col, coef_idx, budget=None, mask=None, remaining=None):
        n = loc.numel()
        for name, t in (("loc", loc), ("year", year), ("coef_col", coef_col), ("coef_idx", coef_idx),
                        ("budget", budget), ("remaining", remaining)):
            self._check_dev(name, t, torch.int32, n)
        self._check_dev("mask", mask, (torch.uint8, torch.bool), n)
        self._check_dev("obs", obs, torch.float32, n * _lib.OBS_DIM)
        self._check_dev("state", state, torch.uint8, n * 16)
        _lib.check(self._lib.w2a_reset(self.ctx, self._stream(), n, _ptr(loc), _ptr(year), _ptr(coef_col),
                                       _ptr(coef_idx), _ptr(budget), _ptr(mask), _ptr(state), _ptr(obs),
                                       _ptr(remaining)), self.ctx)

    def step(self, state, action, obs, reward, terminated, remaining=None, at_budget=None,
             variant=_lib.STEP_AUTO):
        n = action.numel()
        if action.dtype not in _ACT_DTYPES:
            raise TypeError(f"action dtype {action.dtype} not supported (uint8/bool/int32/int64)")
        self._check_dev("action", action, None, n)
        self._check_dev("obs", obs, torch.float32, n * _lib.OBS_DIM)
        self._check_dev("state", state, torch.uint8, n * 16)
        self._check_dev("reward", reward, torch.float32, n)
        self._check_dev("terminated", terminated, (torch.uint8, torch.bool), n)
        self._check_dev("remaining", remaining, torch.int32, n)
        self._check_dev("at_budget", at_budget, (torch.uint8, torch.bool), n)
        _lib.check(self._lib.w2a_step(self.ctx, self._stream(), n, _ptr(action), _ACT_DTYPES[action.dtype],
                                      _ptr(state), _ptr(obs), _ptr(reward), _ptr(terminated), _ptr(remaining),
                                      _ptr(at_budget), variant), self.ctx)

    # Smallest env set stepped on the tabulated logits.  With the tables in place they win at every size (B200, graph
    # replay, 60-county tables: 1.95 / 2.37 / 2.46 us per step at 32 / 1 024 / 4 096 envs against COOP's 2.54 / 2.78 /
    # 3.22; level at 4 096 envs on the 746-county tables), so this is a knob for experiments, not a crossover.
    TABLE_MIN_ENVS = 1

    def choose_variant(self, n: int, canonical: bool) -> int:
        """Resolve STEP_AUTO above the C ABI: TABLE (tabulated logits, one record + one row per env) when
        the tables exist and every env uses its own county's coefficient column (no similar-county
        substitution -- the C library cannot know that, so its own AUTO is COOP)."""
        if canonical and getattr(self, "has_tables", False) and n >= self.TABLE_MIN_ENVS:
            return _lib.STEP_TABLE
        return _lib.STEP_AUTO

    def bind_step(self, state, obs, reward, terminated, remaining=None, at_budget=None,
                  variant=_lib.STEP_AUTO, action_dtype=torch.uint8):
        """Validate the persistent buffers once and return `f(action_tensor)` that only
        performs the C call (the per-step host cost of a rollout loop)."""
        n = reward.numel()
        self._check_dev("obs", obs, torch.float32, n * _lib.OBS_DIM)
        self._check_dev("state", state, torch.uint8, n * 16)
        self._check_dev("reward", reward, torch.float32, n)
        self._check_dev("terminated", terminated, (torch.uint8, torch.bool), n)
        self._check_dev("remaining", remaining, torch.int32, n)
        self._check_dev("at_budget", at_budget, (torch.uint8, torch.bool), n)
        fn, ctx, check = self._lib.w2a_step, self.ctx, _lib.check
        adt = _ACT_DTYPES[action_dtype]
        args = (_ptr(state), _ptr(obs), _ptr(reward), _ptr(terminated), _ptr(remaining), _ptr(at_budget), variant)
        stream = self._stream
        device = self.device

        def step(action):
            if action.dtype != action_dtype or action.numel() != n or action.device != device:
                raise ValueError(f"bound step expects {n} actions of dtype {action_dtype} on {device}")
            rc = fn(ctx, stream(), n, C.c_void_p(action.data_ptr()), adt, *args)
            if rc:
                check(rc, ctx)

        return step

    def rollout(self, state, steps: int, ep_return, actions=None, seed: int = 0, p_alert: float = 0.5, stats=None):
        """Episode-fused rollout (`w2a_rollout`): up to `steps` steps per env in one launch."""
        n = ep_return.numel()
        self._check_dev("state", state, torch.uint8, n * 16)
        self._check_dev("ep_return", ep_return, torch.float64, n)
        self._check_dev("stats", stats, torch.float64, 15)
        if actions is not None:
            self._check_dev("actions", actions, (torch.uint8, torch.bool), int(steps) * n)
        _lib.check(self._lib.w2a_rollout(self.ctx, self._stream(), n, int(steps), _ptr(actions), int(seed) & (2**64 - 1),
                                         float(p_alert), _ptr(state), _ptr(ep_return), _ptr(stats)), self.ctx)

    def rollout_policy(self, state, steps: int, ep_return, policy, stochastic: bool = False, seed: int = 0, stats=None):
        """Episode-fused rollout with a logistic policy evaluated inside the launch (`w2a_rollout_policy`):
        `policy` = float64 [30] CUDA tensor, one weight per observation slot (slot 27 ignored) + bias."""
        n = ep_return.numel()
        self._check_dev("state", state, torch.uint8, n * 16)
        self._check_dev("ep_return", ep_return, torch.float64, n)
        self._check_dev("stats", stats, torch.float64, 15)
        self._check_dev("policy", policy, torch.float64, _lib.POLICY_LEN)
        _lib.check(self._lib.w2a_rollout_policy(self.ctx, self._stream(), n, int(steps), _ptr(policy), int(bool(stochastic)),
                                                int(seed) & (2**64 - 1), _ptr(state), _ptr(ep_return), _ptr(stats)),
                   self.ctx)

    def policy_mlp(self, obs, w1, b1, w2, b2: float, action, logit=None, activation: str = "relu",
                   stochastic: bool = False, seed: int = 0, step: int = 0):
        """`action = [w2 . act(W1 . nan_to_num(obs) + b1) + b2 > 0]` for every env in one launch (`w2a_policy_mlp`; the
        contraction runs on the tensor cores, float16 inputs, float32 accumulation).  `w1` float32 [H, 29] (torch.nn.Linear layout), `b1`, `w2`
        float32 [H], H a multiple of 8 up to 64; `action` uint8 [n] out, `logit` float32 [n] out or None."""
        n = action.numel()
        H = b1.numel()
        self._check_dev("obs", obs, torch.float32, n * _lib.OBS_DIM)
        self._check_dev("w1", w1, torch.float32, H * _lib.OBS_DIM)
        self._check_dev("b1", b1, torch.float32, H)
        self._check_dev("w2", w2, torch.float32, H)
        self._check_dev("action", action, (torch.uint8, torch.bool), n)
        self._check_dev("logit", logit, torch.float32, n)
        act = {"relu": 0, "tanh": 1}[activation]
        _lib.check(self._lib.w2a_policy_mlp(self.ctx, self._stream(), n, _ptr(obs), _ptr(w1), _ptr(b1), _ptr(w2), float(b2),
                                            H, act, int(bool(stochastic)), int(seed) & (2**64 - 1), int(step) & 0xFFFFFFFF,
                                            _ptr(action), _ptr(logit)), self.ctx)

    def step_mlp(self, state, action, obs, reward, terminated, remaining, at_budget, w1, b1, w2, b2: float, next_action,
                 logit=None, ep_return=None, activation: str = "relu", stochastic: bool = False, seed: int = 0, step: int = 0,
                 variant: int | None = None):
        """One iteration of `a = policy(obs); env.step(a)` in one launch (`w2a_step_mlp`): the step of `action` and the
        MLP policy's `next_action` from the fresh observations, which are written but never read back.  Arguments as
        in `step` / `policy_mlp`; `ep_return` (float64 [n], optional) accumulates the rewards."""
        n = action.numel()
        H = b1.numel()
        if variant is None:
            variant = _lib.STEP_TABLE if self.has_tables else _lib.STEP_COOP
        self._check_dev("action", action, torch.uint8, n)
        self._check_dev("next_action", next_action, torch.uint8, n)
        if action.data_ptr() == next_action.data_ptr():
            raise ValueError("step_mlp: next_action must not alias action")
        self._check_dev("obs", obs, torch.float32, n * _lib.OBS_DIM)
        self._check_dev("reward", reward, torch.float32, n)
        self._check_dev("terminated", terminated, (torch.uint8, torch.bool), n)
        self._check_dev("remaining", remaining, torch.int32, n)
        self._check_dev("at_budget", at_budget, (torch.uint8, torch.bool), n)
        self._check_dev("w1", w1, torch.float32, H * _lib.OBS_DIM)
        self._check_dev("b1", b1, torch.float32, H)
        self._check_dev("w2", w2, torch.float32, H)
        self._check_dev("logit", logit, torch.float32, n)
        self._check_dev("ep_return", ep_return, torch.float64, n)
        act = {"relu": 0, "tanh": 1}[activation]
        _lib.check(self._lib.w2a_step_mlp(self.ctx, self._stream(), n, _ptr(action), _ptr(state), _ptr(obs), _ptr(reward),
                                          _ptr(terminated), _ptr(remaining), _ptr(at_budget), int(variant), _ptr(w1), _ptr(b1),
                                          _ptr(w2), float(b2), H, act, int(bool(stochastic)), int(seed) & (2**64 - 1),
                                          int(step) & 0xFFFFFFFF, _ptr(next_action), _ptr(logit), _ptr(ep_return)), self.ctx)

    @property
    def launch_count(self) -> int:
        return int(self._lib.w2a_launch_count(self.ctx))

    # ---- host-buffer entry points (numpy in / numpy out) --------------------------------
    def host_envs_create(self, n: int):
        _lib.check(self._lib.w2a_host_envs_create(self.ctx, n), self.ctx)
        self._host_n = n

    def reset_host(self, loc, year, coef_col, coef_idx, obs, budget=None, mask=None, remaining=None):
        def p(a):
            return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())
        _lib.check(self._lib.w2a_reset_host(self.ctx, p(loc), p(year), p(coef_col), p(coef_idx), p(budget),
                                            p(mask), p(obs), p(remaining)), self.ctx)

    def step_host(self, action, obs, reward, terminated, remaining=None, at_budget=None):
        def p(a):
            return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())
        _lib.check(self._lib.w2a_step_host(self.ctx, p(action), p(obs), p(reward), p(terminated), p(remaining),
                                           p(at_budget)), self.ctx)
