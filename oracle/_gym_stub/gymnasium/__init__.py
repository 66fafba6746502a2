"""Minimal stand-in for `gymnasium`, used ONLY when the real package is not installed.

TEST INFRASTRUCTURE (oracle/): the reference env imports `Env` and `spaces` and nothing
else (`/root/reference/src/weather2alert/env.py:9,17,31,89,95`).  This stub lets the
unmodified reference run in a container that has no gymnasium wheel.
"""
from . import spaces  # noqa: F401


class Env:
    def __init__(self):
        pass
