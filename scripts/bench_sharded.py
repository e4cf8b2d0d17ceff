"""Row-sharded LUCJ / DC-Hamiltonian-energy benchmark (torchrun, one rank per GPU).

    torchrun --nproc-per-node N scripts/bench_sharded.py --norb 20 --nelec 8 8 --n-reps 3

Times apply_unitary(UCJOpSpinBalanced) on a sharded Hartree-Fock state (BASELINE config C4) and
optionally the DiagonalCoulombHamiltonian energy (C5), device-timed, max over ranks; reports the
per-rank HBM and NVLink volumes and the parity of the first rotation against Slater minors.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import ffsim_b200 as ffsim
from ffsim_b200 import distributed
from ffsim_b200.distributed import ShardedVector
from ffsim_b200.gates.orbital_rotation import get_plan


def slater_minors(u, norb, nocc):
    """det U[I, :nocc] for every string I (ascending): the closed form of U|HF> (python/ffsim/states/slater.py:303-354)."""
    occ = ffsim.cistring.gen_occslst(norb, nocc).astype(np.int64)
    sub = u[occ[:, :, None], np.arange(nocc)[None, None, :]]
    out = np.empty(len(occ), dtype=complex)
    step = 20000
    for i in range(0, len(occ), step):
        out[i:i + step] = np.linalg.det(sub[i:i + step])
    return out


def wick_energy(ham, u, nocc_a, nocc_b):
    """<H> of the determinant U|HF> from its 1-RDM: an O(n^2) closed form."""
    pa = u[:, :nocc_a] @ u[:, :nocc_a].conj().T
    pb = u[:, :nocc_b] @ u[:, :nocc_b].conj().T
    h, (jaa, jab) = np.asarray(ham.one_body_tensor), np.asarray(ham.diag_coulomb_mats)
    e = ham.constant + np.trace(h @ pa).real + np.trace(h @ pb).real
    for p_ in (pa, pb):  # same-spin: <n_p n_q> = P_pp P_qq - |P_pq|^2 (p != q), P_pp (p == q)
        d = np.real(np.diag(p_))
        nn = np.outer(d, d) - np.abs(p_) ** 2
        np.fill_diagonal(nn, d)
        e += 0.5 * np.sum(jaa * nn)
    da, db = np.real(np.diag(pa)), np.real(np.diag(pb))
    e += 0.5 * np.sum(jab * (np.outer(da, db) + np.outer(db, da)))
    return float(e)


def run_c5(norb, nelec, dev, world, rank, sync):
    ham = ffsim.random.random_diagonal_coulomb_hamiltonian(norb, seed=2405)
    u = ffsim.random.random_unitary(norb, seed=2406)
    dim = ffsim.dim(norb, nelec)
    sync()
    t0 = time.perf_counter()
    state = ShardedVector.hartree_fock(norb, nelec, device=dev)
    state = ffsim.apply_orbital_rotation(state, u, norb, nelec, copy=False)
    sync()
    prep_ms = (time.perf_counter() - t0) * 1e3
    linop = ffsim.linear_operator(ham, norb=norb, nelec=nelec)
    times, energy = [], None
    for it in range(3):
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        hv = linop @ state
        energy = state.vdot(hv).real
        b.record()
        sync()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
        del hv
    want = wick_energy(ham, u, nelec[0], nelec[1])
    out = {"workload": f"DiagonalCoulombHamiltonian linear_operator energy after orbital rotation, norb={norb} "
                       f"nelec={list(nelec)}, row-sharded", "n_gpus": world, "dim": dim, "state_GB": dim * 16 / 1e9,
           "state_preparation_ms": prep_ms, "energy_ms": float(np.median(times[1:])), "times_ms": times,
           "energy": energy, "energy_closed_form_wick": want, "abs_err": abs(energy - want),
           "norm": state.norm(), "algorithmic_hbm_bytes_matvec": 240 * dim,
           "peak_device_GB": torch.cuda.max_memory_allocated() / 1e9}
    out["algorithmic_TBps_total"] = out["algorithmic_hbm_bytes_matvec"] / (out["energy_ms"] * 1e-3) / 1e12
    if rank == 0:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--norb", type=int, default=14)
    ap.add_argument("--nelec", type=int, nargs=2, default=[6, 6])
    ap.add_argument("--n-reps", type=int, default=3)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--energy", action="store_true")
    ap.add_argument("--c5", action="store_true",
                    help="BASELINE config C5 instead: DiagonalCoulombHamiltonian energy of a rotated determinant")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    norb, nelec = args.norb, tuple(args.nelec)
    ffsim.init_cache(norb, nelec)
    pairs_aa = [(p, p + 1) for p in range(norb - 1)]
    pairs_ab = [(p, p) for p in range(norb)]
    op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=args.n_reps, interaction_pairs=(pairs_aa, pairs_ab),
                                                  seed=2004)
    dim = ffsim.dim(norb, nelec)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.c5:
        run_c5(norb, nelec, dev, world, rank, sync)
        if world > 1:
            dist.destroy_process_group()
        return

    # parity of one sharded rotation against the closed form (outer product of Slater minors)
    hf = ShardedVector.hartree_fock(norb, nelec, device=dev)
    u = op.orbital_rotations[0]
    rot = ffsim.apply_orbital_rotation(hf, u, norb, nelec, copy=False)
    rot.set_layout(distributed.ROWS)  # the check below walks the row shard
    ma, mb = slater_minors(u, norb, nelec[0]), slater_minors(u, norb, nelec[1])
    # outer(ma[rows], mb) is built on the device in row blocks (the full product is shard-sized)
    ma_d = torch.from_numpy(ma[rot.row0:rot.row0 + rot.n_rows]).to(dev)
    mb_d = torch.from_numpy(mb).to(dev)
    local2d = rot.local.view(rot.n_rows, rot.dim_b)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    for r0 in range(0, rot.n_rows, 1024):
        want = torch.outer(ma_d[r0:r0 + 1024], mb_d)
        acc[0] += torch.linalg.vector_norm(local2d[r0:r0 + 1024] - want) ** 2
        acc[1] += torch.linalg.vector_norm(want) ** 2
        del want
    if world > 1:
        dist.all_reduce(acc)
    parity = float(torch.sqrt(acc[0] / acc[1]))
    del rot, local2d, ma_d, mb_d
    torch.cuda.empty_cache()

    times = []
    state = None
    for it in range(args.steps + 1):
        del state  # (the previous result may sit in the column distribution: start from a fresh row shard)
        state = ShardedVector.hartree_fock(norb, nelec, device=dev)
        distributed.STATS.update(exchanges=0, bytes_sent=0)
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        state = ffsim.apply_unitary(state, op, norb=norb, nelec=nelec, copy=False)
        b.record()
        sync()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(float(t))
    norm = state.norm()
    out = {"workload": f"LUCJ n_reps={args.n_reps} on Hartree-Fock, norb={norb} nelec={list(nelec)}, row-sharded",
           "n_gpus": world, "dim": dim, "state_GB": dim * 16 / 1e9, "ms": float(np.median(times)),
           "applications_per_s": 1e3 / float(np.median(times)), "norm": norm, "times_ms": times,
           "plan": get_plan(norb, nelec, op.orbital_rotations[0], op.orbital_rotations[0]).describe(),
           "peak_device_GB": torch.cuda.max_memory_allocated() / 1e9,
           "rotation_parity_vs_slater_minors": parity,
           "all_to_all_per_application": distributed.STATS["exchanges"], "exchange": distributed.STATS["mode"],
           "nvlink_bytes_sent_per_rank_per_application": distributed.STATS["bytes_sent"],
           "algorithmic_hbm_bytes_total": ((args.n_reps + 1) * 64 + args.n_reps * 32) * dim}
    # the three floors of the application, side by side (whole job over all ranks)
    import math

    from ffsim_b200 import _lib
    from ffsim_b200.linalg import givens_decomposition

    n_rot = sum(len(givens_decomposition(np.asarray(m))[0]) for m in
                [op.orbital_rotations[0].T.conj()] +
                [op.orbital_rotations[k + 1].T.conj() @ op.orbital_rotations[k] for k in range(args.n_reps - 1)] +
                [op.orbital_rotations[-1]])
    pairs = (math.comb(norb - 2, nelec[0] - 1) * math.comb(norb, nelec[1])
             + math.comb(norb - 2, nelec[1] - 1) * math.comb(norb, nelec[0]))
    fp64_peak = _lib.measure_fp64_peak()
    out["floors_ms"] = {
        "fp64": 2.0 * 12.0 * n_rot * pairs / (world * fp64_peak * 1e12) * 1e3,
        "hbm": out["algorithmic_hbm_bytes_total"] / (world * 6465.2e9) * 1e3,
        "nvlink": out["nvlink_bytes_sent_per_rank_per_application"] / 770e9 * 1e3,
        "fp64_peak_tflops_per_gpu": fp64_peak, "rotations_per_spin_total": n_rot}
    out["nvlink_GBps_if_comm_bound"] = out["nvlink_bytes_sent_per_rank_per_application"] / (out["ms"] * 1e-3) / 1e9
    out["algorithmic_TBps_total"] = out["algorithmic_hbm_bytes_total"] / (out["ms"] * 1e-3) / 1e12
    if args.energy:
        ham = ffsim.random.random_diagonal_coulomb_hamiltonian(norb, seed=2405)
        linop = ffsim.linear_operator(ham, norb=norb, nelec=nelec)
        sync()
        t0 = time.perf_counter()
        hv = linop @ state
        energy = state.vdot(hv).real
        sync()
        out["energy_ms"] = (time.perf_counter() - t0) * 1e3
        out["energy"] = energy
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
