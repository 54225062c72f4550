"""GPU (-m gpu, needs >= 2 GPUs): row-sharded multi-GPU solve equals the 1-GPU solve."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["csr", "tiled"])
def test_shard_invariance(mode):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(HERE / "_multi_worker.py"), mode]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTI_OK" in res.stdout
    assert "MULTIVIEW_OK" in res.stdout
    assert "NYSTROM_OK" in res.stdout
    assert "KNN_OK" in res.stdout
