import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden", "reference_trajectories.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """The committed outputs of the unmodified reference env (tests/golden/make_golden.py)."""

    def __init__(self):
        z = np.load(GOLDEN)
        self.z = z
        self.meta = json.loads(str(z["meta_json"]))
        self.similar = json.loads(str(z["similar_json"]))
        self.fips = [str(x) for x in z["fix_fips"]]
        self.years = [int(x) for x in z["fix_years"]]
        self.n_days = z["fix_n_days"]
        self.feat = z["fix_feat"]
        self.sig_code = z["fix_sig_code"]
        self.sig_labels = [str(x) for x in z["fix_sig_labels"]]
        self.conf_fips = [str(x) for x in z["fix_conf_fips"]]
        self.conf_zone = [str(x) for x in z["fix_conf_zone"]]
        self.fips_list = [str(x) for x in z["fix_fips_list"]]
        self.coef = {"std": z["fix_coef_std"], "large": z["fix_coef_large"]}

    def store(self):
        from weather2alert_b200.store import FeatureStore

        return FeatureStore(fips=self.fips, years=self.years, n_days=self.n_days, feat=self.feat,
                            sig_code=self.sig_code, sig_labels=self.sig_labels,
                            conf_fips=self.conf_fips, conf_zone=self.conf_zone)

    def weights(self, name):
        from weather2alert_b200.store import RewardWeights

        return RewardWeights(fips_list=list(self.fips_list), coef=self.coef[name])

    def episode(self, key):
        z = self.z
        return {k: z[f"{key}_{k}"] for k in ("obs0", "actions", "obs", "reward", "done", "rem", "atb")}

    def all_episodes(self):
        for sc in self.meta:
            for ep in sc["episodes"]:
                yield sc, ep


@pytest.fixture(scope="session")
def golden():
    return Golden()


def assert_obs_equal(got, want, msg=""):
    """Bit-for-bit on values (NaN == NaN), after casting the float64 golden to `got`'s dtype."""
    want = np.asarray(want).astype(got.dtype)
    np.testing.assert_array_equal(got, want, err_msg=msg)
