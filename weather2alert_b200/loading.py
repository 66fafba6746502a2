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
    import yaml
    from safetensors.numpy import load_file

    tensors = load_file(_resolve(MODEL_REPO, "model", weights, "posterior_samples.safetensors", data_dir))
    with open(_resolve(MODEL_REPO, "model", weights, "config.yaml", data_dir)) as fh:
        config = yaml.safe_load(fh)
    return RewardWeights.from_tensors(tensors, config["fips_list"])  # KeyError as env.py:75
