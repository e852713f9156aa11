"""GPU, N>1: row-range sharded Q6/Q1/Q3 with NCCL merges == the oracle on the whole table.
Skipped on a single-GPU box (run it with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_queries_match_oracle(world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("multi-GPU parity OK") == world


@pytest.mark.parametrize("world", [1, 2])
def test_single_process_drives_all_devices(world):
    """pg_init_devices / pg_use_device: one process, one host thread per GPU, NCCL between the threads
    (tests/multidev_check.py).  world = 1 runs on any GPU box."""
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multidev_check.py"), str(world)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "parity OK" in r.stdout
