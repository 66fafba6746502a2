#!/usr/bin/env python
"""Extract the county list of the reference's `nn_full_medicare_all` checkpoint.

The checkpoint blobs are not in the reference checkout (`.MISSING_LARGE_BLOBS`), but its
`config.yaml` is, and `HeatAlertEnv.__init__` takes `fips_list` from it (env.py:74-75): 720
counties against the 746 of `weights/linear`.  BASELINE configs[2] names this layout, so the bench
(`--weights-layout nn`) and the tests use the real list with synthetic coefficients.  Run here
(where /root/reference exists); the output is committed because the GPU box has no reference mount.
"""
import os
import sys

import yaml

SRC = "/root/reference/weights/nn_full_medicare_all/config.yaml"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weather2alert_b200", "data",
                   "nn_full_medicare_all_fips.txt")

with open(SRC) as fh:
    cfg = yaml.safe_load(fh)
fips = [str(x) for x in cfg["fips_list"]]  # env.py:75
assert all(len(f) == 5 and f.isdigit() for f in fips), "not 5-digit FIPS strings"
with open(DST, "w") as fh:
    fh.write("\n".join(fips) + "\n")
print(f"{len(fips)} counties -> {DST}", file=sys.stderr)
