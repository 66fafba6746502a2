"""weather2alert_b200 -- B200-native batched HeatAlertEnv reset/step hot path.

Drop-in surface of NSAPH-Projects/weather2alert's `HeatAlertEnv` plus a vector env that
returns torch CUDA tensors; the work is done by hand-written sm_100a kernels behind the
C ABI in `include/w2a.h`.  There is no CPU fallback.
"""
from . import schema  # noqa: F401
from .datautils import ClimateGroups, get_similar_counties  # noqa: F401
from .store import FeatureStore, RewardWeights  # noqa: F401

__all__ = ["HeatAlertEnv", "HeatAlertVectorEnv", "HeatAlertCore", "FeatureStore", "RewardWeights",
           "ClimateGroups", "get_similar_counties", "schema"]


def __getattr__(name):  # lazy: these need torch + the built extension
    if name == "HeatAlertEnv":
        from .env import HeatAlertEnv

        return HeatAlertEnv
    if name == "HeatAlertVectorEnv":
        from .vector_env import HeatAlertVectorEnv

        return HeatAlertVectorEnv
    if name == "HeatAlertCore":
        from .core import HeatAlertCore

        return HeatAlertCore
    raise AttributeError(name)
