"""Developer timing script (not the contract bench): device-timed ops at a BASELINE shape."""

import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import ffsim_b200 as ffsim
from ffsim_b200 import _lib
from ffsim_b200.gates.orbital_rotation import get_plan


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times)), float(np.min(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--norb", type=int, default=16)
    ap.add_argument("--nelec", type=int, nargs=2, default=[5, 5])
    ap.add_argument("--opts", type=str, default="")
    ap.add_argument("--only-rot", action="store_true")
    args = ap.parse_args()
    norb, nelec = args.norb, tuple(args.nelec)
    for kv in filter(None, args.opts.split(",")):
        k, v = kv.split("=")
        _lib.set_option(k, int(v))
    dim = ffsim.dim(norb, nelec)
    rng = np.random.default_rng(1)
    u = ffsim.random.random_unitary(norb, seed=rng)
    mat = ffsim.random.random_real_symmetric_matrix(norb, seed=rng)
    vec = torch.randn(dim, dtype=torch.complex128, device="cuda")
    vec /= torch.linalg.vector_norm(vec)
    plan = get_plan(norb, nelec, u, u)
    out = {"norb": norb, "nelec": nelec, "dim": dim, "opts": args.opts, "plan": plan.describe(),
           "state_passes": plan.n_state_passes()}
    gb = dim * 16 / 1e9
    med, best = timeit(lambda: ffsim.apply_orbital_rotation(vec, u, norb, nelec, copy=False))
    out["orbital_rotation_ms"] = med
    out["orbital_rotation_alg_GBps"] = 4 * gb / (best / 1e3)
    med, best = timeit(lambda: ffsim.apply_orbital_rotation(vec, (u, None), norb, nelec, copy=False))
    out["alpha_only_ms"] = med
    if args.only_rot:
        med, best = timeit(lambda: ffsim.apply_orbital_rotation(vec, (None, u), norb, nelec, copy=False))
        out["beta_only_ms"] = med
        print(json.dumps({k: out[k] for k in ("opts", "plan", "orbital_rotation_ms", "alpha_only_ms", "beta_only_ms")}))
        return
    med, best = timeit(lambda: ffsim.apply_orbital_rotation(vec, (None, u), norb, nelec, copy=False))
    out["beta_only_ms"] = med
    med, best = timeit(lambda: ffsim.apply_diag_coulomb_evolution(vec, mat, 1.0, norb, nelec, copy=False))
    out["diag_coulomb_ms"] = med
    out["diag_coulomb_GBps"] = 2 * gb / (best / 1e3)
    med, best = timeit(lambda: ffsim.apply_diag_coulomb_evolution(vec, mat, 1.0, norb, nelec, z_representation=True, copy=False))
    out["diag_coulomb_z_ms"] = med
    coeffs = rng.standard_normal(norb)
    med, best = timeit(lambda: ffsim.apply_num_op_sum_evolution(vec, coeffs, 1.0, norb, nelec, copy=False))
    out["num_op_sum_ms"] = med
    out["num_op_sum_GBps"] = 2 * gb / (best / 1e3)
    med, best = timeit(lambda: ffsim.contract_diag_coulomb(vec, mat, norb, nelec))
    out["contract_dc_ms"] = med
    w = torch.empty_like(vec)
    med, best = timeit(lambda: w.copy_(vec))
    out["torch_copy_ms"] = med
    out["torch_copy_GBps"] = 2 * gb / (best / 1e3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
