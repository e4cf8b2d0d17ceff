"""Sharded-state GPU tests: world size 1 in-process, world size 2 under torchrun when two GPUs exist."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def test_sharded_world_size_1():
    out = subprocess.run([sys.executable, WORKER], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]


def test_sharded_world_size_2_nccl():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29711", WORKER],
        capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
