"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi`):
user-sharded training on N GPUs must reproduce the single-GPU (= oracle) result on the same global batches."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _physical_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:  # noqa: BLE001
        return torch.cuda.device_count()


@pytest.mark.parametrize("comm", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2])
def test_sharded_training_matches_oracle(world, comm):
    if _physical_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, DRB_SHARDED_COMM=comm)
    env.pop("CUDA_VISIBLE_DEVICES", None)      # an earlier test may have mirrored config['gpu'] into it
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "mp_sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "-> OK" in r.stdout and f"comm={comm}" in r.stdout
