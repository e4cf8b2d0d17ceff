"""Target for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the library on small
shapes -- the fused pass kernel with every register-block width (2..6), single- and multi-pass plans,
native and transposed beta side (transpositions folded into the passes or separate, bulk copies on and off),
the host pipeline; the diagonal kernels (evolution, contraction, z representation), the
controlled phase, the _lib-level single-rotation kernels, transpose, vdot, axpby, block copies."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ffsim_b200 as ffsim
from ffsim_b200 import _lib
from ffsim_b200.gates.orbital_rotation import apply_orbital_rotation_unfused

rng = np.random.default_rng(0)
n = 0
for norb, nelec in [(4, (2, 2)), (6, (3, 2)), (8, (4, 4)), (9, (3, 5)), (10, (5, 4))]:
    dim = ffsim.dim(norb, nelec)
    vec = ffsim.random.random_state_vector(dim, seed=rng)
    ua, ub = ffsim.random.random_unitary(norb, seed=rng), ffsim.random.random_unitary(norb, seed=rng)
    mat = ffsim.random.random_real_symmetric_matrix(norb, seed=rng)
    for opts in [{}, {"sub_window": 2}, {"sub_window": 3}, {"sub_window": 4}, {"sub_window": 5},
                 {"smem_bytes": 4096, "min_cols": 2}, {"smem_bytes": 8192, "beta_mode": 2}, {"smem_bytes": 8192, "beta_mode": 3},
                 {"smem_bytes": 4096, "min_cols": 2, "sub_window": 5, "beta_mode": 2, "bulk_copies": 0}, {"bulk_copies": 0},
                 {"beta_mode": 1, "smem_bytes": 16384},
                 {"threads": 128}]:
        saved = {k: _lib.get_option(k) for k in opts}
        for k, v in opts.items():
            _lib.set_option(k, v)
        out = ffsim.apply_orbital_rotation(vec, (ua, ub), norb, nelec)
        assert abs(np.linalg.norm(out) - 1) < 1e-10
        n += 1
        for k, v in saved.items():
            _lib.set_option(k, v)
    for z in (False, True):
        ffsim.apply_diag_coulomb_evolution(vec, (mat, mat + 0.1, mat), 0.3, norb, nelec, z_representation=z)
        ffsim.contract_diag_coulomb(vec, (mat, mat, mat), norb, nelec, z_representation=z)
    ffsim.apply_num_op_sum_evolution(vec, rng.standard_normal(norb), 0.2, norb, nelec)
    ffsim.contract_num_op_sum(vec, rng.standard_normal(norb), norb, nelec)
    ffsim.apply_num_num_interaction(vec, 0.3, (0, 1), norb, nelec)
    ffsim.apply_on_site_interaction(vec, 0.3, 1, norb, nelec)
    ffsim.apply_fsim_gate(vec, 0.3, 0.2, (1, 2), norb, nelec)
    if norb <= 6:
        apply_orbital_rotation_unfused(vec, (ua, ub), norb, nelec)
    if norb in (6, 9):  # host pipeline: strips in, blocks out, two applications in flight
        outs = ffsim.evolve_host_many([vec, vec, vec], [("orbital_rotation", (ua, ub)), ("diag_coulomb", mat, 0.2)],
                                      norb, nelec, n_chunks=2)
        assert all(abs(np.linalg.norm(o) - 1) < 1e-10 for o in outs)
    ham = ffsim.random.random_diagonal_coulomb_hamiltonian(norb, seed=rng)
    lin = ffsim.linear_operator(ham, norb=norb, nelec=nelec)
    hv = lin @ vec
    n += 8
torch.cuda.synchronize()
print("sanitize target done:", n, "calls")
