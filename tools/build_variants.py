#!/usr/bin/env python
"""Build tuning variants of libw2a.so (compile-time macros of w2a_kernels.cuh) into build/variants/.
`W2A_LIB_PATH=build/variants/libw2a_<name>.so python bench.py ...` runs one of them (weather2alert_b200/_lib.py)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from __graft_entry__ import NVCC_FLAGS  # noqa: E402

VARIANTS = {
    "base": [],
    "z128": ["-DW2A_Z_LOAD=0"],
    "zkeep": ["-DW2A_Z_POLICY=1"],
    "mb6": ["-DW2A_TABLE_MINBLOCKS=6"],
    "rows4mb6": ["-DW2A_TABLE_ROWS=4", "-DW2A_TABLE_MINBLOCKS=6"],
    "rows1": ["-DW2A_TABLE_ROWS=1"],
    "w8": ["-DW2A_TABLE_WARPS=8", "-DW2A_TABLE_MINBLOCKS=4"],
    "w2": ["-DW2A_TABLE_WARPS=2", "-DW2A_TABLE_MINBLOCKS=16"],
    "k1": ["-DW2A_KNOCK=1"], "k2": ["-DW2A_KNOCK=2"], "k4": ["-DW2A_KNOCK=4"], "k8": ["-DW2A_KNOCK=8"],
    "k16": ["-DW2A_KNOCK=16"], "k32": ["-DW2A_KNOCK=32"], "k48": ["-DW2A_KNOCK=48"],
    "k64": ["-DW2A_KNOCK=64"], "k128": ["-DW2A_KNOCK=128"], "rows4": ["-DW2A_TABLE_ROWS=4"],
    "k6": ["-DW2A_KNOCK=6"], "k9": ["-DW2A_KNOCK=9"], "k15": ["-DW2A_KNOCK=15"],
}


def build(name):
    out = os.path.join(REPO, "build", "variants", f"libw2a_{name}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *VARIANTS[name], "-I", os.path.join(REPO, "include"), "-o", out,
           os.path.join(REPO, "weather2alert_b200", "csrc", "w2a.cu")]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(8) as ex:
        for o in ex.map(build, names):
            print(o)
