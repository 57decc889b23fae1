"""Multi-GPU best-trajectory reduction behind the C ABI (include/mpcgpu_dist.h): CPU check that the library loads and
exports what the header declares; GPU check (needs 2 GPUs: `gpurun --gpus 2`) that ONE control step whose guesses are split
over two ranks ends with the same winner on both ranks and that it is the single-GPU winner, bit for bit."""
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dist_library_exports_every_declared_symbol():
    from mpc_planner_b200 import build, dist

    build.build_dist()
    header = open(os.path.join(ROOT, "include", "mpcgpu_dist.h")).read()
    declared = set(re.findall(r"\b(mpcgpu_dist_[a-z0-9_]+)\s*\(", header))
    assert declared == set(dist.EXPORTED_SYMBOLS), declared ^ set(dist.EXPORTED_SYMBOLS)
    lib = dist.load_dist_library()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.mpcgpu_dist_strerror(-1) == b"invalid argument"
    assert lib.mpcgpu_dist_create(None, 0, 1, None, None) != 0  # argument checks come before any CUDA / NCCL call


@pytest.mark.gpu
def test_sharded_control_step_agrees_with_single_gpu_winner_nccl_world_size_2():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "bench.py"), "--gpus", "2", "--mode", "sharded", "--guesses", "8,64,256", "--steps", "3", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["mode"] == "sharded" and line["n_gpus"] == 2 and len(line["rows"]) == 3
    for row in line["rows"]:
        assert row["identical_on_every_rank"] is True and row["equals_single_gpu_winner"] is True, row
        assert row["winner"] >= 0 and row["latency_ms"] > 0
