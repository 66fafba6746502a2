"""CPU tests: file loaders in both directory layouts, the host action stream, and bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from weather2alert_b200 import synthetic
from weather2alert_b200.distributed import rollout_action_bits
from weather2alert_b200.loading import load_store, load_weights

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_loaders_flat_and_hub_layout(tmp_path, golden):
    store, w = golden.store(), golden.weights("std")
    flat = tmp_path / "flat"
    synthetic.write_reference_files(store, w, str(flat))
    s1, w1 = load_store(str(flat)), load_weights("whatever_name", str(flat))
    np.testing.assert_array_equal(s1.feat, store.feat)
    np.testing.assert_array_equal(w1.coef, w.coef)
    # hub layout: data/<split>/*.parquet and <weights>/{posterior_samples.safetensors,config.yaml}
    hub = tmp_path / "hub"
    (hub / "data" / "65k").mkdir(parents=True)
    (hub / "nn_full_medicare_all").mkdir()
    for f in os.listdir(flat):
        dst = hub / "data" / "65k" / f if f.endswith(".parquet") else hub / "nn_full_medicare_all" / f
        os.replace(flat / f, dst)
    s2, w2 = load_store(str(hub), split="65k"), load_weights("nn_full_medicare_all", str(hub))
    np.testing.assert_array_equal(s2.feat, store.feat)
    assert s2.conf_fips == store.conf_fips and w2.fips_list == w.fips_list
    np.testing.assert_array_equal(s2.pack_x(), store.pack_x())
    np.testing.assert_array_equal(w2.pack(), w.pack())
    with pytest.raises(FileNotFoundError):
        load_weights("missing_weights", str(hub))


def test_rollout_action_bits_stream():
    a = rollout_action_bits(7, np.arange(1000), np.arange(153), 0.25)
    b = rollout_action_bits(7, np.arange(1000), np.arange(153), 0.25)
    assert a.shape == (153, 1000) and a.dtype == bool and (a == b).all()
    assert abs(a.mean() - 0.25) < 0.01
    # counter based: a sub-block equals the same block of the full stream; seeds decorrelate
    sub = rollout_action_bits(7, np.arange(100, 200), np.arange(50, 60), 0.25)
    assert (sub == a[50:60, 100:200]).all()
    assert (rollout_action_bits(8, np.arange(1000), np.arange(153), 0.25) != a).mean() > 0.3
    assert not rollout_action_bits(7, np.arange(10), np.arange(10), 0.0).any()
    assert rollout_action_bits(7, np.arange(10), np.arange(10), 1.0).all()


@pytest.mark.timeout(600)
def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the unmodified reference env, or the C port when baseline/_ref
    is absent) runs on the host cores and prints exactly one JSON line with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "4",
                          "--warmup", "3", "--ref-workers", "2"], capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
