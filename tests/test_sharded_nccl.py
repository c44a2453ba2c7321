"""Fused row-sharded mixed loss on real GPUs against the fp64 oracle (a torchrun-able test against
closed_form instead of a script against the repo's own kernels).  One rank always (the shard kernels with the
exchange switched off); two ranks over NCCL + peer mailboxes when the box has two GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "sharded_nccl_worker.py")


def _run(cmd, timeout):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    return r


def test_sharded_single_rank_vs_oracle():
    r = _run([sys.executable, WORKER, "256", "4", "16", "16", "1", "2"], 300)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nccl_vs_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    port = 29611 + world
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
              "127.0.0.1", "--master-port", str(port), WORKER, "1024", "4", "16", "16", "1", "5"], 600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout
