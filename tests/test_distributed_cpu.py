"""World-size-2 gloo tests (CPU): the host-side logic of the row-sharded path.

The all-to-all redistribution (row shards <-> column shards) is product code and runs
unchanged on CPU tensors; the compute that would run on the GPU between the two
redistributions is stood in for by the CPU oracle here, so the test checks that the
sharded scheme (local beta ops, transposed alpha ops, uneven splits, scalar reductions)
reproduces the single-process result.
"""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, math
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from ffsim_b200.distributed import ShardedVector, partition, all_to_all_bytes, STATS, ROWS, COLS
from oracle import gates, givens, rand, models, cistring

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for norb, nelec in [(5, (2, 3)), (6, (3, 3)), (4, (1, 4)), (3, (0, 2))]:
    dim_a, dim_b = models.dims(norb, nelec)
    rng = np.random.default_rng(norb)
    full = rand.random_state_vector(dim_a * dim_b, seed=rng)
    ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
    sv = ShardedVector.from_global(full, norb, nelec)
    assert sv.a_off == partition(dim_a, world) and sv.n_rows == sv.a_off[rank + 1] - sv.a_off[rank]
    # gather round trip and scalars
    assert np.array_equal(sv.gather().numpy(), full)
    assert abs(sv.norm() - 1.0) < 1e-13
    other = ShardedVector.from_global(full[::-1].copy(), norb, nelec)
    assert abs(sv.vdot(other) - np.vdot(full, full[::-1])) < 1e-13
    # redistribution: column shard == the global matrix restricted to this rank's columns
    b0, b1 = sv.b_off[rank], sv.b_off[rank + 1]
    assert all_to_all_bytes(sv) == 16 * sv.n_rows * (dim_b - (b1 - b0))
    n_ex = STATS["exchanges"]
    sv.set_layout(COLS)
    assert sv.layout == COLS and STATS["exchanges"] == n_ex + 1
    cols = sv.local.view(dim_a, b1 - b0)
    assert np.array_equal(cols.numpy(), full.reshape(dim_a, dim_b)[:, b0:b1])
    blk = sv.block()
    assert blk[1:] == (0, dim_a, b0, b1 - b0, b1 - b0)
    assert all_to_all_bytes(sv) == 16 * (dim_a - sv.n_rows) * (b1 - b0)
    # scalars work in either distribution (the other operand follows)
    assert abs(sv.norm() - 1.0) < 1e-13
    assert abs(sv.vdot(other) - np.vdot(full, full[::-1])) < 1e-13 and other.layout == COLS
    # alpha rotation on the column shard (oracle compute), then back
    work = np.ascontiguousarray(cols.numpy())
    gates._rotate_one_spin(work, givens.givens_decomposition(ua), norb, nelec[0])
    sv.local.copy_(torch.from_numpy(work).reshape(-1))
    sv.set_layout(COLS)  # no-op
    assert STATS["exchanges"] == n_ex + 2  # (the vdot moved `other`)
    sv.set_layout(ROWS)
    assert sv.block()[1:] == (sv.row0, sv.n_rows, 0, dim_b, dim_b)
    # beta rotation is local to the row shard
    loc = np.ascontiguousarray(sv.local.numpy().reshape(sv.n_rows, dim_b).T)
    gates._rotate_one_spin(loc, givens.givens_decomposition(ub), norb, nelec[1])
    sv.local.copy_(torch.from_numpy(np.ascontiguousarray(loc.T).reshape(-1)))
    want = gates.apply_orbital_rotation(full, (ua, ub), norb, nelec)
    got = sv.gather().numpy()
    assert np.linalg.norm(got - want) < 1e-13, (norb, nelec, np.linalg.norm(got - want))
    # a diagonal op only needs the row offset of the block
    mat = rand.random_real_symmetric_matrix(norb, seed=rng)
    want2 = gates.apply_diag_coulomb_evolution(want, mat, 0.3, norb, nelec)
    blk = want2.reshape(dim_a, dim_b)[sv.row0:sv.row0 + sv.n_rows]
    assert blk.shape[0] == sv.n_rows
hf = ShardedVector.hartree_fock(4, (2, 2), device="cpu")
g = hf.gather().numpy()
assert g[0] == 1 and np.count_nonzero(g) == 1
dist.destroy_process_group()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scheme_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    port = 29600 + world
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
        capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == world


def test_partition():
    from ffsim_b200.distributed import partition

    assert partition(10, 3) == [0, 4, 7, 10]
    assert partition(2, 4) == [0, 1, 2, 2, 2]
    offs = partition(125970, 8)
    assert offs[-1] == 125970 and max(b - a for a, b in zip(offs, offs[1:])) - min(b - a for a, b in zip(offs, offs[1:])) <= 1
