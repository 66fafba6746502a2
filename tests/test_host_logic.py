"""CPU tests of the host side: packing, climate groups, sampling order, ABI surface."""
import os
import re

import numpy as np
import pytest

from weather2alert_b200 import schema, synthetic
from weather2alert_b200.datautils import ClimateGroups, get_similar_counties
from weather2alert_b200.sampling import EpisodeSampler
from weather2alert_b200.store import (W_STRIDE, X_SIG, X_STRIDE, FeatureStore, RewardWeights, coef_pos_other,
                                      coef_pos_table)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schema_matches_reference_feature_names(golden):
    names = golden.meta[0]["episodes"][0]["feature_names"]
    assert list(schema.OBS_NAMES) == names
    assert len(schema.TABLE_REWARD_SLOTS) == 24 and schema.OBS_DIM == 29


def test_pack_layouts_round_trip(golden):
    store, w = golden.store(), golden.weights("std")
    S, Y, T = store.shape
    x = store.pack_x().reshape(S, Y, T, X_STRIDE)
    for u, slot in enumerate(schema.TABLE_REWARD_SLOTS):
        np.testing.assert_array_equal(x[..., u], store.feat[..., slot])
    sig = x[..., X_SIG]
    np.testing.assert_array_equal(np.where(sig == 0, np.nan, sig).astype(np.float32), store.sig_code)
    heat_qi = store.feat[..., 0]
    np.testing.assert_array_equal(x[..., 25], np.where(heat_qi > 0.5, 0.0, -1e300))
    assert (x[..., 26:] == 0).all() and np.isfinite(x).all()
    packed = w.pack().reshape(len(w.fips_list), w.n_samples, W_STRIDE)
    seen = set()
    for h in (0, 1):
        for u, slot in enumerate(schema.TABLE_REWARD_SLOTS):
            np.testing.assert_array_equal(packed[:, :, coef_pos_table(u, h)], w.coef[:, :, h, slot].T)
            seen.add(coef_pos_table(u, h))
        for k, slot in enumerate((23, 24, 25, 27)):
            np.testing.assert_array_equal(packed[:, :, coef_pos_other(k, h)], w.coef[:, :, h, slot].T)
            seen.add(coef_pos_other(k, h))
    assert (packed[..., 51] == 1.0).all()  # effectiveness coefficient of the gate column
    assert len(seen) == 56 and all((packed[..., p] == 0).all() for p in range(W_STRIDE) if p not in seen and p != 51)
    assert seen == set(range(48)) | set(range(56, 64))


def test_store_frames_round_trip(tmp_path, golden):
    store, w = golden.store(), golden.weights("std")
    synthetic.write_reference_files(store, w, str(tmp_path))
    again = FeatureStore.from_parquet_dir(str(tmp_path))
    assert again.fips == store.fips and again.years == store.years and again.sig_labels == store.sig_labels
    np.testing.assert_array_equal(again.n_days, store.n_days)
    np.testing.assert_array_equal(again.feat, store.feat)
    np.testing.assert_array_equal(again.sig_code, store.sig_code)
    assert again.conf_fips == store.conf_fips and again.conf_zone == store.conf_zone
    w2 = RewardWeights.from_dir(str(tmp_path))
    assert w2.fips_list == w.fips_list
    np.testing.assert_array_equal(w2.coef, w.coef)


def test_weights_without_fips_list_raise_keyerror(tmp_path, golden):
    import yaml

    synthetic.write_reference_files(golden.store(), golden.weights("std"), str(tmp_path))
    with open(tmp_path / "config.yaml", "w") as fh:
        yaml.safe_dump({"name": "no_list"}, fh)
    with pytest.raises(KeyError):  # env.py:75
        RewardWeights.from_dir(str(tmp_path))


def test_climate_groups_match_reference_helper(golden):
    groups = ClimateGroups(golden.conf_fips, golden.conf_zone, golden.fips_list)
    for f, want in golden.similar.items():
        assert get_similar_counties(f, golden.conf_fips, golden.conf_zone) == want
        assert groups.similar(f) == [x for x in want if x in golden.fips_list]  # env.py:116
    with pytest.raises(KeyError):
        groups.similar("99999")
    # western states collapse to one group whatever their zone (datautils.py:113-120)
    assert set(golden.similar["06037"]) >= {"04013", "06001", "06037", "06073", "53033", "20173", "08031"}


def test_sampler_reproduces_reference_rng_draws(golden):
    """Same numpy calls in the same order as env.py:143-160 -> same location/year/sample."""
    store = golden.store()
    checked = 0
    for sc in golden.meta:
        w = golden.weights(sc["weights"])
        sampler = EpisodeSampler(store, w, sc["ctor"].get("years"))
        assert [int(y) for y in sampler.valid_years] == sc["valid_years"]
        for ep in sc["episodes"]:
            kw = ep["reset"]
            rng = np.random.default_rng(kw["seed"])
            aug = kw.get("similar_climate_counties", sc["ctor"].get("similar_climate_counties", False))
            d = sampler.draw_reference(rng, kw.get("location"), aug)
            assert d["location"] + "_" + str(d["year"]) == ep["episode_index"]
            assert d["info_location"] == ep["location"]
            assert d["location_index"] == ep["location_index"]
            assert d["coef_index"] == ep["coef_index"]
            assert int(store.n_days[d["s"], d["y"]]) == ep["n_days"]
            checked += 1
    assert checked == 74


def test_sampler_errors_follow_the_reference(golden):
    sampler = EpisodeSampler(golden.store(), golden.weights("std"))
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError):  # list.index, env.py:121
        sampler.draw_reference(rng, "99999", False)
    with pytest.raises(ValueError):  # in the data, not in the weights' county list
        sampler.draw_reference(rng, "06001", False)
    with pytest.raises(KeyError):  # (fips, year) not in the data, env.py:127
        EpisodeSampler(golden.store(), golden.weights("std"), years=[1999]).draw_reference(rng, "06037", False)


def test_bulk_sampler_ranges(golden):
    sampler = EpisodeSampler(golden.store(), golden.weights("std"))
    rng = np.random.default_rng(1)
    s = sampler.draw_bulk(rng, 5000, None, augment=True)
    assert s["loc"].dtype == np.int32 and len(s["loc"]) == 5000
    assert (s["coef_col"] >= 0).all() and (s["coef_col"] < sampler.n_similar[s["list_idx"]]).all()
    assert (s["coef_idx"] >= 0).all() and (s["coef_idx"] < 12).all()
    assert set(np.unique(s["year"])) == {0, 1, 2}
    np.testing.assert_array_equal(s["loc"], sampler.list_to_data[s["list_idx"]])
    s = sampler.draw_bulk(rng, 10, "06037", augment=False)
    assert (s["coef_col"] == golden.fips_list.index("06037")).all()
    with pytest.raises(ValueError):
        sampler.draw_bulk(rng, 10, "99999")


def test_library_exports_every_declared_symbol():
    from weather2alert_b200 import _lib

    header = open(os.path.join(REPO, "include", "w2a.h")).read()
    declared = set(re.findall(r"\b(w2a_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # ImportError if the extension was not built
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.w2a_version() == 210
    assert lib.w2a_state_bytes(10) == 160
    # error paths that need no GPU
    assert lib.w2a_create(0, None) == -1
    assert b"NULL" in lib.w2a_last_error(None)
    assert lib.w2a_destroy(None) == 0


def test_auto_variant_rule_above_the_abi():
    """TABLE only for envs on their own county's coefficients, with the whole-summer tables built (the C library's
    AUTO never picks it)."""
    import types

    from weather2alert_b200 import _lib
    from weather2alert_b200.core import HeatAlertCore

    def pick(n, canonical, has_tables=True):  # no GPU needed
        stub = types.SimpleNamespace(has_tables=has_tables, TABLE_MIN_ENVS=HeatAlertCore.TABLE_MIN_ENVS)
        return HeatAlertCore.choose_variant(stub, n, canonical)

    assert pick(1 << 20, True) == _lib.STEP_TABLE and pick(1 << 23, True) == _lib.STEP_TABLE
    assert pick(1 << 20, False) == _lib.STEP_AUTO  # similar-county augmentation
    assert pick(1 << 20, True, has_tables=False) == _lib.STEP_AUTO  # no room for the tables on the device
    assert pick(HeatAlertCore.TABLE_MIN_ENVS, True) == _lib.STEP_TABLE
    assert pick(HeatAlertCore.TABLE_MIN_ENVS - 1, True) == _lib.STEP_AUTO
    assert _lib.POLICY_LEN == 30 and _lib.STEP_TABLE == 3
    # the Python constants are the header's
    import re

    hdr = open(os.path.join(REPO, "include", "w2a.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define (W2A_[A-Z0-9_]+) (0x[0-9a-fA-F]+|-?\d+)\b", hdr)}
    for name, val in (("W2A_STEP_AUTO", _lib.STEP_AUTO), ("W2A_STEP_THREAD", _lib.STEP_THREAD), ("W2A_STEP_COOP", _lib.STEP_COOP),
                      ("W2A_STEP_TABLE", _lib.STEP_TABLE), ("W2A_STEP_LOCAL", _lib.STEP_LOCAL), ("W2A_OBS_DIM", _lib.OBS_DIM),
                      ("W2A_POLICY_LEN", _lib.POLICY_LEN)):
        assert defs[name] == val, name
    assert _lib.STEP_LOCAL & 0xFF == 0  # a flag above the variant numbers


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(REPO, "weather2alert_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_intended_augmentation_moves_the_episode_to_a_similar_county(golden):
    """`augment="intended"` (not the reference's behaviour): weather AND coefficient column are those
    of a county drawn from the similar-county list; `augment=True` keeps the reference's quirk (Q6/Q7)."""
    store, weights = golden.store(), golden.weights("std")
    sp = EpisodeSampler(store, weights)
    n = 4000
    home = sp.draw_bulk(np.random.default_rng(3), n, None, False)
    ref = sp.draw_bulk(np.random.default_rng(3), n, None, True)
    new = sp.draw_bulk(np.random.default_rng(3), n, None, "intended")
    np.testing.assert_array_equal(ref["loc"], home["loc"])          # Q7: the weather stays the original county's
    assert (ref["coef_col"] < sp.n_similar[ref["list_idx"]]).all()  # Q6: an index inside the similar list
    # intended: the drawn member of the ORIGINAL county's list, used consistently
    np.testing.assert_array_equal(new["coef_col"], new["list_idx"])
    np.testing.assert_array_equal(new["loc"], sp.list_to_data[new["list_idx"]])
    for i in range(0, n, 97):
        members = sp.groups.similar(sp.fips_list[int(home["list_idx"][i])])
        assert members[int(ref["coef_col"][i])] == sp.fips_list[int(new["list_idx"][i])]
    assert (new["list_idx"] != home["list_idx"]).mean() > 0.3  # it really moves
