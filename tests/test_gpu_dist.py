"""Multi-GPU parity: cells sharded over 2 GPUs (one process per GPU, NCCL all-reduce / all-gather
between the stages) must reproduce the oracle exactly like the single-GPU path.  Skipped on boxes
with fewer than 2 GPUs (the CPU gloo test covers the host logic there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_pipeline_nccl(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "helpers", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"DIST_OK world={world}" in r.stdout
