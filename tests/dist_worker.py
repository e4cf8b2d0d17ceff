"""Worker for the sharded GPU tests: run under torchrun (NCCL, one rank per GPU) or alone.

Checks the row-sharded path against the CPU oracle on a case small enough to gather:
LUCJ (apply_unitary), rotated diagonal Coulomb evolution, the DiagonalCoulombHamiltonian
energy, and a double-factorized Trotter step.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import ffsim_b200 as ffsim
from ffsim_b200.distributed import ShardedVector
from oracle import cref, models, rand


def main():
    if os.environ.get("FFSIM_B200_WATCHDOG"):  # a hung collective prints every thread's stack and exits
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["FFSIM_B200_WATCHDOG"]), exit=True)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    tol = 1e-12
    for norb, nelec in [(8, (4, 3)), (10, (5, 5)), (6, (1, 6))]:
        rng = np.random.default_rng(norb)
        full = rand.random_state_vector(models.dim(norb, nelec), seed=rng)
        # LUCJ
        op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=2, with_final_orbital_rotation=True, seed=norb)
        sv = ShardedVector.from_global(full, norb, nelec, device=dev)
        out = ffsim.apply_unitary(sv, op, norb=norb, nelec=nelec)
        assert isinstance(out, ShardedVector) and out is not sv
        want = models.ucj_spin_balanced_apply(full, op.diag_coulomb_mats, op.orbital_rotations,
                                              op.final_orbital_rotation, norb, nelec)
        got = out.gather().cpu().numpy()
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err < tol, ("lucj", norb, nelec, err)
        assert np.array_equal(sv.gather().cpu().numpy(), full)  # copy=True left the input alone
        # independent alpha / beta rotations, in place
        ua, ub = rand.random_unitary(norb, seed=rng), rand.random_unitary(norb, seed=rng)
        sv2 = ffsim.apply_orbital_rotation(sv, (ua, ub), norb, nelec, copy=False)
        assert sv2 is sv
        want = cref.apply_orbital_rotation(full, (ua, ub), norb, nelec)
        err = np.linalg.norm(sv.gather().cpu().numpy() - want) / np.linalg.norm(want)
        assert err < tol, ("rot", norb, nelec, err)
        # rotated diagonal Coulomb evolution, z representation
        mat = rand.random_real_symmetric_matrix(norb, seed=rng)
        sv = ShardedVector.from_global(full, norb, nelec, device=dev)
        out = ffsim.apply_diag_coulomb_evolution(sv, mat, 0.4, norb, nelec, orbital_rotation=ua, z_representation=True)
        from oracle import gates

        want = gates.apply_diag_coulomb_evolution(full, mat, 0.4, norb, nelec, orbital_rotation=ua, z_representation=True)
        err = np.linalg.norm(out.gather().cpu().numpy() - want) / np.linalg.norm(want)
        assert err < tol, ("dc", norb, nelec, err)
        # energy of a diagonal Coulomb Hamiltonian
        ham = ffsim.random.random_diagonal_coulomb_hamiltonian(norb, seed=norb + 1)
        linop = ffsim.linear_operator(ham, norb=norb, nelec=nelec)
        hv = linop @ sv
        energy = sv.vdot(hv).real
        want_hv = models.diagonal_coulomb_hamiltonian_matvec(full, ham.one_body_tensor, ham.diag_coulomb_mats,
                                                            ham.constant, norb, nelec)
        assert abs(energy - np.vdot(full, want_hv).real) < 1e-10, ("energy", norb, nelec)
        # Trotter step
        df = ffsim.random.random_double_factorized_hamiltonian(norb, rank=3, seed=norb + 2)
        out = ffsim.simulate_trotter_double_factorized(sv, df, 0.2, norb=norb, nelec=nelec, n_steps=1, order=1)
        want = models.simulate_trotter_double_factorized(
            full, df.one_body_tensor, df.diag_coulomb_mats, df.orbital_rotations, df.constant, False, 0.2,
            norb=norb, nelec=nelec, n_steps=1, order=1)
        err = np.linalg.norm(out.gather().cpu().numpy() - want) / np.linalg.norm(want)
        assert err < tol, ("trotter", norb, nelec, err)
        # host-resident shards: three overlapped applications of (rotation, diagonal Coulomb) on this rank's rows
        from ffsim_b200.distributed import partition

        dim_a, dim_b = models.dims(norb, nelec)
        offs = partition(dim_a, world)
        rank = dist.get_rank() if world > 1 else 0
        rows = full.reshape(-1, dim_b)[offs[rank]:offs[rank + 1]]
        steps = [("orbital_rotation", (ua, ub)), ("diag_coulomb", mat, 0.3)]
        handles = [ffsim.evolve_host_rows_async(rows, steps, norb, nelec) for _ in range(3)]
        want = gates.apply_diag_coulomb_evolution(cref.apply_orbital_rotation(full, (ua, ub), norb, nelec), mat, 0.3,
                                                  norb, nelec).reshape(-1, dim_b)[offs[rank]:offs[rank + 1]]
        for h in handles:
            got = h.result().reshape(want.shape)
            assert np.linalg.norm(got - want) <= tol * max(np.linalg.norm(want), 1e-300) + 1e-15, ("host rows", norb, nelec)
        del handles, h
    hf = ShardedVector.hartree_fock(8, (4, 4), device=dev)
    assert abs(hf.norm() - 1) < 1e-15
    if world > 1 and os.environ.get("FFSIM_B200_REPORT_EXCHANGE"):
        from ffsim_b200 import distributed

        print("exchange=" + ("p2p" if distributed.p2p_available(hf) else "nccl"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print("rank", os.environ.get("RANK", "0"), "ok")


if __name__ == "__main__":
    main()
