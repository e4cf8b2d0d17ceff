"""Random input generators (oracle; test infrastructure only).

Restates python/ffsim/random/random.py: random_state_vector :24-47, random_unitary
:84-106, random_hermitian :151-167, random_real_symmetric_matrix :170-189,
_random_symmetric_matrix_uniform :535-546, random_ucj_op_spin_balanced :563-665,
random_diagonal_coulomb_hamiltonian :883-908, random_double_factorized_hamiltonian
:911-956.  A ``np.random.Generator`` passed as ``seed`` is consumed in the same
order as the reference consumes it.
"""

from __future__ import annotations

import math

import numpy as np


def random_state_vector(dim, *, seed=None):
    rng = np.random.default_rng(seed)
    vec = rng.standard_normal(dim).astype(complex)
    vec += 1j * rng.standard_normal(dim)
    return vec / np.linalg.norm(vec)


def random_unitary(dim, *, seed=None):
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((dim, dim)).astype(complex)
    z += 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def random_hermitian(dim, *, seed=None):
    rng = np.random.default_rng(seed)
    mat = rng.standard_normal((dim, dim)).astype(complex)
    mat += 1j * rng.standard_normal((dim, dim))
    return mat + mat.T.conj()


def random_real_symmetric_matrix(dim, *, rank=None, seed=None):
    rng = np.random.default_rng(seed)
    mat = rng.standard_normal((dim, dim if rank is None else rank))
    return mat @ mat.T


def _symmetric_uniform(dim, scale, rng):
    vals = rng.uniform(-0.5 * scale, 0.5 * scale, size=dim * (dim + 1) // 2)
    mat = np.zeros((dim, dim))
    rows, cols = np.triu_indices(dim)
    mat[rows, cols] = vals
    mat[cols, rows] = vals
    return mat


def random_ucj_op_spin_balanced(norb, *, n_reps=1, interaction_pairs=None, with_final_orbital_rotation=False, seed=None):
    """Returns (diag_coulomb_mats[L,2,n,n], orbital_rotations[L,n,n], final|None)."""
    pairs_aa, pairs_ab = (None, None) if interaction_pairs is None else interaction_pairs
    rng = np.random.default_rng(seed)
    mats = np.stack(
        [
            np.stack([_symmetric_uniform(norb, 2 * math.pi, rng), _symmetric_uniform(norb, 2 * math.pi, rng)])
            for _ in range(n_reps)
        ]
    )
    rots = np.stack([random_unitary(norb, seed=rng) for _ in range(n_reps)])
    final = random_unitary(norb, seed=rng) if with_final_orbital_rotation else None
    for which, pairs in ((0, pairs_aa), (1, pairs_ab)):
        if pairs is not None:
            mask = np.zeros((norb, norb), dtype=bool)
            for p, q in pairs:
                mask[p, q] = mask[q, p] = True
            mats[:, which] *= mask
    return mats, rots, final


def random_diagonal_coulomb_hamiltonian(norb, *, seed=None):
    """Returns (one_body_tensor, diag_coulomb_mats[2,n,n], constant)."""
    rng = np.random.default_rng(seed)
    one_body = random_hermitian(norb, seed=rng)
    mat_aa = random_real_symmetric_matrix(norb, seed=rng)
    mat_ab = random_real_symmetric_matrix(norb, seed=rng)
    return one_body, np.stack([mat_aa, mat_ab]), rng.standard_normal()


def random_double_factorized_hamiltonian(norb, *, rank=None, z_representation=False, seed=None):
    """Returns (one_body_tensor, diag_coulomb_mats[L,n,n], orbital_rotations[L,n,n], constant, z_rep)."""
    if rank is None:
        rank = norb * (norb + 1) // 2
    rng = np.random.default_rng(seed)
    one_body = random_hermitian(norb, seed=rng)
    rots = np.stack([random_unitary(norb, seed=rng) for _ in range(rank)])
    mats = np.stack([random_real_symmetric_matrix(norb, seed=rng) for _ in range(rank)])
    return one_body, mats, rots, rng.standard_normal(), z_representation
