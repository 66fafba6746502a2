"""ctypes binding of `libw2a.so` (C ABI declared in `include/w2a.h`).

There is no CPU fallback: if the library has not been built, importing this module
raises, and so does every env class of the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# W2A_LIB_PATH selects another build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("W2A_LIB_PATH") or os.path.join(_HERE, "libw2a.so")

W2A_OK = 0
ACT_U8, ACT_I32, ACT_I64 = 0, 1, 2
STEP_AUTO, STEP_THREAD, STEP_COOP, STEP_TABLE = 0, 1, 2, 3
STEP_LOCAL = 0x100  # OR into a variant: env slots are in table order (tuning hint, w2a.h)
POLICY_LEN = 30  # W2A_POLICY_LEN: one weight per observation slot + bias
BUDGET_DEFAULT = -(2**31)
OBS_DIM = 29
FEATURE_WORKSPACE_BYTES = 4096

#: every symbol `include/w2a.h` declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "w2a_version": (C.c_int, []),
    "w2a_last_error": (C.c_char_p, [_P]),
    "w2a_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "w2a_destroy": (C.c_int, [_P]),
    "w2a_set_features": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "w2a_set_coefs": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "w2a_set_county_map": (C.c_int, [_P, _P]),
    "w2a_precompute_logits": (C.c_int, [_P, _P]),
    "w2a_precompute_bytes": (C.c_size_t, [_P]),
    "w2a_state_bytes": (C.c_size_t, [C.c_int64]),
    "w2a_reset": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "w2a_step": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int]),
    "w2a_rollout": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_uint64, C.c_float, _P, _P, _P]),
    "w2a_rollout_policy": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int, C.c_uint64, _P, _P, _P]),
    "w2a_host_envs_create": (C.c_int, [_P, C.c_int64]),
    "w2a_host_envs_destroy": (C.c_int, [_P]),
    "w2a_reset_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "w2a_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "w2a_engineer_features": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "w2a_policy_mlp": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, C.c_float, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                 C.c_uint32, _P, _P]),
    "w2a_step_mlp": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, C.c_float, C.c_int, C.c_int,
                               C.c_int, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "w2a_launch_count": (C.c_int64, [_P]),
}


class W2AError(RuntimeError):
    pass


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc: int, ctx=None) -> None:
    if rc != W2A_OK:
        msg = lib().w2a_last_error(ctx).decode(errors="replace")
        raise W2AError(f"libw2a error {rc}: {msg}")
