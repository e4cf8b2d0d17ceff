"""Short target for ncu: a few applications of the hot path at a BASELINE shape."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ffsim_b200 as ffsim
from ffsim_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--norb", type=int, default=16)
ap.add_argument("--nelec", type=int, nargs=2, default=[5, 5])
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--opts", type=str, default="")
args = ap.parse_args()
for kv in filter(None, args.opts.split(",")):
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
norb, nelec = args.norb, tuple(args.nelec)
rng = np.random.default_rng(1)
u = ffsim.random.random_unitary(norb, seed=rng)
mat = ffsim.random.random_real_symmetric_matrix(norb, seed=rng)
vec = torch.randn(ffsim.dim(norb, nelec), dtype=torch.complex128, device="cuda")
for _ in range(args.reps):
    ffsim.apply_orbital_rotation(vec, u, norb, nelec, copy=False)
    ffsim.apply_diag_coulomb_evolution(vec, mat, 1.0, norb, nelec, copy=False)
torch.cuda.synchronize()
print("done")
