"""bench.py on a GPU: one JSON line with every key of the measurement contract."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_bench_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--envs-per-gpu", "65536", "--steps", "306",
                          "--warmup", "153", "--cpu-seconds", "2", "--e2e-steps", "5"],
                         capture_output=True, text=True, timeout=850)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 306 and d["warmup"] >= 3 and d["scaling"] == "weak"
    assert d["value"] > 1e8 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] >= 306
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and r["achieved"] > 0
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and "traffic" in r
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "env-steps/s" and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 65536 and e["d2h_bytes_per_step"] == 65536 * 126
    assert e["value"] < d["value"]  # host copies are inside the e2e region
    k = d["clocks"]
    assert set(k) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["rollout"]["same_returns_as_stepping"] is True and d["return_stats"]["episodes"] == 65536
    assert d["single_env"]["ms_per_step"] > 0 and d["single_env"]["steps"] == 153
    # round 2: the oracle gate ran on the bench's own tables before anything was timed, the timed region is made of
    # graph-replayed K-step regions, the e2e number is read against the link's own ceiling
    pg = d["parity"]
    assert pg["ok"] is True and pg["obs_bit_exact"] is True and pg["envs"] >= 1 << 16 and pg["steps"] >= 10
    assert pg["max_reward_rel"] <= 1e-5
    assert d["timing"]["repeats"] >= 1 and "graphs" in d["config"]["mode"]
    assert e["link_gbs"] > 1 and 0 < e["frac_of_link"] < 1.5
    assert d["sorted"]["value"] > 0 and d["vector_env_default"]["sort_episodes"] is True
    assert d["policy_in_loop"]["fused_mlp_policy"]["value"] > d["policy_in_loop"]["value"]
    assert set(d["other_configs"]) == {"cfg4_2m", "cfg5_8m"} and all("error" not in v for v in d["other_configs"].values())
