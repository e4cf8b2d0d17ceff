"""Device-timed runs of the BASELINE.json configurations that fit one GPU (developer/report tool).

    python scripts/bench_configs.py c1      # LUCJ n_reps=2 on Hartree-Fock, norb=12 nelec=(6,6)
    python scripts/bench_configs.py c3      # DF Trotter step, norb=18 nelec=(7,7), rank 18 (16 GB state)

Prints one JSON line per configuration: ms per application, algorithmic HBM bytes
and, where the host can do it in seconds, the restated reference beside it and the parity.
"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ffsim_b200 as ffsim


def dev_time(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def c1():
    from oracle import cref

    norb, nelec, L = 12, (6, 6), 2
    pairs = ([(p, p + 1) for p in range(norb - 1)], [(p, p) for p in range(norb)])
    op = ffsim.random.random_ucj_op_spin_balanced(norb, n_reps=L, interaction_pairs=pairs, seed=1201)
    hf = ffsim.hartree_fock_state(norb, nelec)
    dim = hf.size
    dvec = torch.from_numpy(hf).cuda()
    ms = dev_time(lambda: ffsim.apply_unitary(dvec, op, norb=norb, nelec=nelec), warm=2, reps=9)
    got = ffsim.apply_unitary(hf, op, norb=norb, nelec=nelec)
    t0 = time.perf_counter()
    want = cref.ucj_spin_balanced_apply(hf, op.diag_coulomb_mats, op.orbital_rotations, op.final_orbital_rotation,
                                        norb, nelec)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    api = []
    for _ in range(7):  # steady state: the first calls pin the pooled host buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ffsim.apply_unitary(hf, op, norb=norb, nelec=nelec)
        api.append((time.perf_counter() - t0) * 1e3)
    api_ms = float(np.median(api[2:]))
    alg = ((L + 1) * 64 + L * 32) * dim
    return {"config": "C1: LUCJ n_reps=2 on Hartree-Fock, norb=12 nelec=(6,6), 853,776 amplitudes",
            "device_ms": ms, "public_api_numpy_ms": api_ms, "applications_per_s": 1e3 / ms,
            "algorithmic_bytes": alg, "algorithmic_GBps": alg / ms / 1e6,
            "cpu_restated_reference_ms": cpu_ms, "cpu_threads": cref.n_threads(),
            "rel_err_vs_cpu_oracle": float(np.linalg.norm(got - want) / np.linalg.norm(want))}


def c3(rank=18, order=0):
    norb, nelec = 18, (7, 7)
    dim = ffsim.dim(norb, nelec)
    ham = ffsim.random.random_double_factorized_hamiltonian(norb, rank=rank, seed=1803)
    g = torch.Generator(device="cuda").manual_seed(1804)
    vec = torch.randn(dim, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128)
    vec += 1j * torch.randn(dim, dtype=torch.float64, device="cuda", generator=g)
    vec /= torch.linalg.vector_norm(vec)
    ffsim.init_cache(norb, nelec)
    state = {"v": vec}

    def step():
        state["v"] = ffsim.simulate_trotter_double_factorized(state["v"], ham, 0.1, norb=norb, nelec=nelec,
                                                              n_steps=1, order=order, copy=False)

    ms = dev_time(step, warm=1, reps=2)
    norm = float(torch.linalg.vector_norm(state["v"]))
    n_terms = rank if order == 0 else 2 * rank
    alg = (n_terms * 96 + 64) * dim
    return {"config": f"C3: double-factorized Trotter step (order {order}), random DF Hamiltonian rank {rank}, "
                      f"norb=18 nelec=(7,7), {dim} amplitudes ({dim * 16 / 1e9:.1f} GB)",
            "device_ms": ms, "steps_per_s": 1e3 / ms, "algorithmic_bytes": alg,
            "algorithmic_TBps": alg / ms / 1e9, "norm_after": norm,
            "hbm_floor_ms_at_6.65TBps": alg / 6.65e9, "peak_device_GB": torch.cuda.max_memory_allocated() / 1e9}


if __name__ == "__main__":
    for name in sys.argv[1:] or ["c1"]:
        print(json.dumps({"c1": c1, "c3": c3}[name]()))
