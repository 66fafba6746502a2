"""ctypes wrapper of oracle/c/heat_oracle.c (TEST INFRASTRUCTURE; see the C file header)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libheat_oracle.so")


class _Tables(C.Structure):
    _fields_ = [("S", C.c_int64), ("Y", C.c_int64), ("T", C.c_int64), ("P", C.c_int64), ("Sw", C.c_int64),
                ("feat", C.c_void_p), ("sig", C.c_void_p), ("coef", C.c_void_p), ("n_days", C.c_void_p)]


class _State(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("loc", "year", "col", "p", "budget", "t", "streak", "used", "n_steps",
                                          "last14", "obs")]


def available() -> bool:
    return os.path.exists(LIB)


class COracle:
    def __init__(self, feat, sig_code, n_days, coef, n):
        if not available():
            raise ImportError(f"{LIB} not built; run __graft_entry__.build()")
        self.lib = C.CDLL(LIB)
        self.lib.ho_max_threads.restype = C.c_int
        S, Y, T, _ = feat.shape
        P, Sw = coef.shape[:2]
        self._keep = [np.ascontiguousarray(feat, np.float64), np.ascontiguousarray(sig_code, np.float32),
                      np.ascontiguousarray(coef, np.float32), np.ascontiguousarray(n_days, np.uint8)]
        self.tb = _Tables(S, Y, T, P, Sw, *(a.ctypes.data for a in self._keep))
        self.n = n
        self.arr = {k: np.zeros(n, np.int32) for k in ("loc", "year", "col", "p", "budget", "t", "streak", "used", "n_steps")}
        self.arr["last14"] = np.zeros((n, 14), np.uint8)
        self.arr["obs"] = np.zeros((n, 29), np.float64)
        self.st = _State(*(self.arr[k].ctypes.data for k, _ in _State._fields_))
        self.reward = np.zeros(n, np.float64)
        self.done = np.zeros(n, np.uint8)
        self.rem = np.zeros(n, np.int32)
        self.atb = np.zeros(n, np.uint8)

    @property
    def threads(self) -> int:
        return int(self.lib.ho_max_threads())

    def reset(self, loc, year, col, p, budget=None):
        a = [np.ascontiguousarray(x, np.int32) for x in (loc, year, col, p)]
        b = None if budget is None else np.ascontiguousarray(budget, np.int32)
        self.lib.ho_reset(C.byref(self.tb), C.byref(self.st), C.c_int64(self.n), *(C.c_void_p(x.ctypes.data) for x in a),
                          None if b is None else C.c_void_p(b.ctypes.data), C.c_void_p(self.rem.ctypes.data))
        return self.arr["obs"]

    def step(self, action):
        act = np.ascontiguousarray(action, np.uint8)
        self.lib.ho_step(C.byref(self.tb), C.byref(self.st), C.c_int64(self.n), C.c_void_p(act.ctypes.data),
                         C.c_void_p(self.reward.ctypes.data), C.c_void_p(self.done.ctypes.data),
                         C.c_void_p(self.rem.ctypes.data), C.c_void_p(self.atb.ctypes.data))
        return self.arr["obs"], self.reward, self.done, self.rem, self.atb
