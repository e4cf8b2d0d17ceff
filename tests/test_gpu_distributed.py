"""Sharded-state GPU tests: world size 1 in-process, world size 2 under torchrun when two GPUs exist.

Both implementations of the redistribution are covered: the peer-memory exchange kernel
(symmetric memory over NVLink, the default on one box) and NCCL all_to_all_single."""

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


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_world_size_n(exchange, world):
    """The worker's checks (LUCJ, rotations, rotated diagonal Coulomb, DC-Hamiltonian energy, a Trotter step;
    shapes with empty shards included) with the state distributed over 2, 4 and 8 ranks, both exchanges."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, FFSIM_B200_EXCHANGE=exchange, FFSIM_B200_REPORT_EXCHANGE="1", FFSIM_B200_WATCHDOG="240")
    port = 29700 + 2 * world + (exchange == "p2p")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER],
        capture_output=True, text=True, timeout=400, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world
    # the worker reports which path it actually took: no silent fallback
    assert out.stdout.count(f"exchange={exchange}") == world, out.stdout[-2000:]
