"""Run the UNMODIFIED reference `HeatAlertEnv` on local fixture files.

TEST INFRASTRUCTURE -- not part of the product.  Only `tests/`, `tests/golden/make_golden.py`,
`__graft_entry__.smoke()` and `bench.py`'s reference / cpu_baseline legs may import this.

The reference env (`src/weather2alert/env.py` of NSAPH-Projects/weather2alert) downloads
its data and weights from the Hugging Face hub in its constructor (`env.py:40-47,60-67`)
and subclasses `gymnasium.Env`.  Neither the hub nor gymnasium is available here, so:

  1. if `import gymnasium` fails, `oracle/_gym_stub` (a 20-line `Env`/`Box`/`Discrete`)
     is put on `sys.path`;
  2. `huggingface_hub.hf_hub_download` is replaced, *before* the reference module is
     imported (it binds the name at import, `env.py:10`), by a function that returns
     ``<fixture_dir>/<filename>``.

No reference source is modified or copied.  The reference package is looked up in
`baseline/_ref` (a `pip install --target` of /root/reference made by
`__graft_entry__.build()`, git-ignored) and then in `/root/reference/src`.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
_FIXTURE_DIR = {"path": None}


def reference_location() -> str | None:
    """Directory that holds the `weather2alert` package of the reference, or None."""
    for cand in (os.path.join(_REPO, "baseline", "_ref"), "/root/reference/src"):
        if os.path.exists(os.path.join(cand, "weather2alert", "env.py")):
            return cand
    return None


def _fake_hf_hub_download(repo_id=None, repo_type=None, subfolder=None, filename=None, local_dir=None, **kw):
    if _FIXTURE_DIR["path"] is None:
        raise RuntimeError("oracle.ref_harness: fixture dir not set")
    return os.path.join(_FIXTURE_DIR["path"], filename)


def load_reference_class():
    """Import and return the reference's `HeatAlertEnv` class (patched I/O only)."""
    loc = reference_location()
    if loc is None:
        raise ImportError("reference package not found in baseline/_ref or /root/reference/src")
    try:
        import gymnasium  # noqa: F401
    except ImportError:
        stub = os.path.join(_HERE, "_gym_stub")
        if stub not in sys.path:
            sys.path.insert(0, stub)
    import huggingface_hub

    huggingface_hub.hf_hub_download = _fake_hf_hub_download
    if loc not in sys.path:
        sys.path.insert(0, loc)
    mod = importlib.import_module("weather2alert.env")
    mod.hf_hub_download = _fake_hf_hub_download
    return mod.HeatAlertEnv


def make_reference_env(fixture_dir: str, **ctor_kwargs):
    """Construct the reference env over the five files in `fixture_dir`."""
    cls = load_reference_class()
    _FIXTURE_DIR["path"] = fixture_dir
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return cls(**ctor_kwargs)


def encode_obs(obs, sig_labels):
    """Reference obs (object array of 29, str/NaN in slot 27) -> float64[29]."""
    import numpy as np

    out = np.empty(len(obs), dtype=np.float64)
    for i, v in enumerate(obs):
        if isinstance(v, str):
            out[i] = float(sig_labels.index(v) + 1)
        elif v is None:
            out[i] = np.nan
        else:
            out[i] = float(v)
    return out
