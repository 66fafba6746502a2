"""1 vs N GPUs under NCCL (last row): the same global index / action streams, sharded over the
GPUs of the box, give byte-identical per-env trajectories, and the all-reduced statistics equal the single-GPU ones."""
import json
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_total", [100_003])
def test_shards_over_nccl_equal_one_gpu(tmp_path, n_total):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    world = min(world, 8)
    out = tmp_path / "multi.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "tests", "multi_gpu_worker.py"), str(out), str(n_total)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["world"] == world and res["nccl"] == "nccl" and res["episodes"] == n_total
    for key in ("state_bytes_equal", "obs_bytes_equal", "returns_equal", "remaining_equal", "rollout_returns_equal",
                "reduced_equals_sum_of_ranks", "counts_equal_single_gpu"):
        assert res[key] is True, (key, res)
    assert res["sums_rel_err"] < 1e-12, res
