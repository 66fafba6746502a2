"""Locate the feature-store and checkpoint files the way the reference constructor does
(`/root/reference/src/weather2alert/env.py:40-47,60-67`), preferring local copies."""
from __future__ import annotations

import os

from .store import FeatureStore, RewardWeights

DATA_REPO = "mauriciogtec/HeatAlertsRL-Data"
MODEL_REPO = "mauriciogtec/HeatAlertsRL-Models"


def _resolve(repo_id, repo_type, subfolder, filename, data_dir):
    if data_dir is not None:
        for cand in (os.path.join(data_dir, subfolder, filename), os.path.join(data_dir, filename)):
            if os.path.exists(cand):
                return cand
    try:
        from huggingface_hub import hf_hub_download

        return hf_hub_download(repo_id=repo_id, repo_type=repo_type, subfolder=subfolder,
                               filename=filename, local_dir=data_dir)
    except Exception as exc:  # offline image: say what was looked for
        raise FileNotFoundError(
            f"{filename} not found under data_dir={data_dir!r} (tried {subfolder}/{filename} and {filename}) "
            f"and could not be fetched from {repo_id}: {exc}") from exc


def load_store(data_dir=None, split="65k") -> FeatureStore:
    import pandas as pd

    frames = {}
    for name in ("confounders", "exogenous_states", "endogenous_states_actions"):
        frames[name] = pd.read_parquet(_resolve(DATA_REPO, "dataset", "data/" + split, name + ".parquet", data_dir))
    return FeatureStore.from_frames(frames["exogenous_states"], frames["endogenous_states_actions"],
                                    frames["confounders"])


def load_weights(weights: str, data_dir=None) -> RewardWeights:
    """`posterior_samples.safetensors` + `config.yaml` of checkpoint `weights` (env.py:60-85).  A checkpoint directory
    that only holds the legacy `posterior_samples.pt` (`weights/linear_debug_sim` in the reference checkout) is read
    with `read_legacy_pt`; what happens next is what the reference's constructor would do with those tensors."""
    import yaml

    try:
        from safetensors.numpy import load_file

        tensors = load_file(_resolve(MODEL_REPO, "model", weights, "posterior_samples.safetensors", data_dir))
    except FileNotFoundError:
        legacy = None
        if data_dir is not None:
            for cand in (os.path.join(data_dir, weights, "posterior_samples.pt"), os.path.join(data_dir, "posterior_samples.pt")):
                if os.path.exists(cand):
                    legacy = cand
        if legacy is None:
            raise
        tensors = read_legacy_pt(legacy)
    with open(_resolve(MODEL_REPO, "model", weights, "config.yaml", data_dir)) as fh:
        config = yaml.safe_load(fh)
    # the reference raises KeyError when the checkpoint has no county list (env.py:75) or no
    # `baseline_bias` tensor (env.py:85) -- both are true of the legacy checkpoint generation
    return RewardWeights.from_tensors(tensors, config["fips_list"])


# ---- legacy checkpoints: a pickled tensordict.TensorDict inside a torch zip archive -------------------------
_PT_DTYPES = {"FloatStorage": "<f4", "DoubleStorage": "<f8", "HalfStorage": "<f2", "LongStorage": "<i8",
              "IntStorage": "<i4", "ShortStorage": "<i2", "CharStorage": "i1", "ByteStorage": "u1", "BoolStorage": "?"}


class _Stub(dict):
    """Stands in for `tensordict.TensorDict` / `_StringOnlyDict`: a plain dict that also takes whatever attribute
    state the pickle hands it (the pickle creates it with NEWOBJ, fills it with SETITEMS / BUILD)."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


def read_legacy_pt(path: str) -> dict:
    """Tensors of a legacy `posterior_samples.pt` (`/root/reference/weights/linear_debug_sim/posterior_samples.pt`:
    a `tensordict.TensorDict` pickled by `torch.save`, the checkpoint generation of `_deprecated/env.py`) as
    ``{name: numpy array}``, WITHOUT executing the pickle: an allow-list unpickler resolves only the six globals
    such a file contains (TensorDict, its string-keyed dict, `_rebuild_tensor_v2`, storages, OrderedDict,
    `torch.Size`) to inert stand-ins and reads the storages straight from the zip archive; anything else raises.
    Neither `tensordict` nor `torch` is imported."""
    import pickle
    import zipfile

    import numpy as np

    zf = zipfile.ZipFile(path)
    pkl = [n for n in zf.namelist() if n.endswith("/data.pkl")]
    if len(pkl) != 1:
        raise ValueError(f"{path}: not a torch zip checkpoint")
    root = pkl[0][: -len("data.pkl")]

    class _Storage:
        def __init__(self, name):
            self.dtype = np.dtype(_PT_DTYPES[name])

    def rebuild_tensor(storage, offset, size, stride, requires_grad=False, hooks=None, metadata=None):
        flat, itemsize = storage
        if not size:
            return flat[offset].copy()
        return np.lib.stride_tricks.as_strided(flat[offset:], shape=tuple(size),
                                               strides=tuple(s * itemsize for s in stride)).copy()

    allowed = {("tensordict._td", "TensorDict"): _Stub, ("tensordict.tensordict", "TensorDict"): _Stub,
               ("tensordict.utils", "_StringOnlyDict"): _Stub,
               ("torch._utils", "_rebuild_tensor_v2"): rebuild_tensor,
               ("collections", "OrderedDict"): dict, ("torch", "Size"): tuple}

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module == "torch" and name in _PT_DTYPES:
                return _Storage(name)
            if (module, name) in allowed:
                return allowed[(module, name)]
            raise pickle.UnpicklingError(f"{path}: global {module}.{name} is not on the allow-list")

        def persistent_load(self, pid):  # ('storage', <storage type>, key, location, numel)
            if not (isinstance(pid, tuple) and pid and pid[0] == "storage" and isinstance(pid[1], _Storage)):
                raise pickle.UnpicklingError(f"{path}: unexpected persistent id {pid!r}")
            raw = zf.read(f"{root}data/{pid[2]}")
            return np.frombuffer(raw, dtype=pid[1].dtype), pid[1].dtype.itemsize

    obj = Unpickler(zf.open(pkl[0])).load()
    items = getattr(obj, "_tensordict", obj)
    if not isinstance(items, dict) or not items:
        raise ValueError(f"{path}: no tensors found in the pickled object")
    return {str(k): v for k, v in items.items() if isinstance(v, np.ndarray)}
