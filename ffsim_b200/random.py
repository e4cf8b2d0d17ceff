"""Random input generators for the hot path.

Same sampling procedures and RNG consumption order as the corresponding
functions of python/ffsim/random/random.py (random_state_vector :24-47,
random_unitary :84-106, random_hermitian :151-167, random_real_symmetric_matrix
:170-189, random_ucj_op_spin_balanced :563-665,
random_diagonal_coulomb_hamiltonian :883-908,
random_double_factorized_hamiltonian :911-956), so a given seed produces the
inputs the reference would produce.
"""

from __future__ import annotations

import math

import numpy as np

from ffsim_b200.hamiltonians import DiagonalCoulombHamiltonian, DoubleFactorizedHamiltonian
from ffsim_b200.variational import UCJOpSpinBalanced, UCJOpSpinless, UCJOpSpinUnbalanced


def random_state_vector(dim: int, *, seed=None, dtype=complex) -> np.ndarray:
    if dim < 1:
        raise ValueError("Dimension must be at least one.")
    rng = np.random.default_rng(seed)
    vec = rng.standard_normal(dim).astype(dtype, copy=False)
    if np.issubdtype(dtype, np.complexfloating):
        vec += 1j * rng.standard_normal(dim).astype(dtype, copy=False)
    vec /= np.linalg.norm(vec)
    return vec


def random_unitary(dim: int, *, seed=None, dtype=complex) -> np.ndarray:
    """Haar-distributed unitary via QR with the phase fix of arXiv:math-ph/0609050."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((dim, dim)).astype(dtype)
    z += 1j * rng.standard_normal((dim, dim)).astype(dtype)
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def random_hermitian(dim: int, *, seed=None, dtype=complex) -> np.ndarray:
    rng = np.random.default_rng(seed)
    mat = rng.standard_normal((dim, dim)).astype(dtype)
    mat += 1j * rng.standard_normal((dim, dim)).astype(dtype)
    return mat + mat.T.conj()


def random_real_symmetric_matrix(dim: int, *, rank: int | None = None, seed=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if rank is None:
        rank = dim
    mat = rng.standard_normal((dim, rank))
    return mat @ mat.T


def _random_symmetric_matrix_uniform(dim: int, *, mean: float, scale: float, seed=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    vals = mean + rng.uniform(-0.5 * scale, 0.5 * scale, size=dim * (dim + 1) // 2)
    mat = np.zeros((dim, dim))
    rows, cols = np.triu_indices(dim)
    mat[rows, cols] = vals
    mat[cols, rows] = vals
    return mat


def _random_symmetric_matrix_normal(dim: int, *, mean: float, scale: float, seed=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    vals = rng.normal(loc=mean, scale=scale, size=dim * (dim + 1) // 2)
    mat = np.zeros((dim, dim))
    rows, cols = np.triu_indices(dim)
    mat[rows, cols] = vals
    mat[cols, rows] = vals
    return mat


def random_ucj_op_spin_balanced(
    norb: int,
    *,
    n_reps: int = 1,
    interaction_pairs=None,
    with_final_orbital_rotation: bool = False,
    diag_coulomb_mean: float = 0.0,
    diag_coulomb_scale: float = 2 * math.pi,
    diag_coulomb_normal: bool = False,
    seed=None,
) -> UCJOpSpinBalanced:
    if interaction_pairs is None:
        interaction_pairs = (None, None)
    pairs_aa, pairs_ab = interaction_pairs
    rng = np.random.default_rng(seed)
    gen = _random_symmetric_matrix_normal if diag_coulomb_normal else _random_symmetric_matrix_uniform
    diag_coulomb_mats = np.stack(
        [
            np.stack(
                [
                    gen(norb, mean=diag_coulomb_mean, scale=diag_coulomb_scale, seed=rng),
                    gen(norb, mean=diag_coulomb_mean, scale=diag_coulomb_scale, seed=rng),
                ]
            )
            for _ in range(n_reps)
        ]
    )
    orbital_rotations = np.stack([random_unitary(norb, seed=rng) for _ in range(n_reps)])
    final_orbital_rotation = random_unitary(norb, seed=rng) if with_final_orbital_rotation else None
    for which, pairs in ((0, pairs_aa), (1, pairs_ab)):
        if pairs is not None:
            mask = np.zeros((norb, norb), dtype=bool)
            if pairs:
                rows, cols = zip(*pairs)
                mask[rows, cols] = True
                mask[cols, rows] = True
            diag_coulomb_mats[:, which] *= mask
    return UCJOpSpinBalanced(
        diag_coulomb_mats=diag_coulomb_mats,
        orbital_rotations=orbital_rotations,
        final_orbital_rotation=final_orbital_rotation,
    )


def _pair_mask(norb: int, pairs, symmetric: bool) -> np.ndarray:
    """Boolean mask of the matrix entries an interaction-pair list allows."""
    mask = np.zeros((norb, norb), dtype=bool)
    if pairs:
        rows, cols = zip(*pairs)
        mask[rows, cols] = True
        if symmetric:
            mask[cols, rows] = True
    return mask


def random_ucj_op_spin_unbalanced(
    norb: int,
    *,
    n_reps: int = 1,
    interaction_pairs=None,
    with_final_orbital_rotation: bool = False,
    diag_coulomb_mean: float = 0.0,
    diag_coulomb_scale: float = 2 * math.pi,
    diag_coulomb_normal: bool = False,
    seed=None,
) -> UCJOpSpinUnbalanced:
    """python/ffsim/random/random.py:668-790: same draws in the same order (per repetition J_aa, J_ab,
    J_bb; then per repetition the alpha and beta rotations; then the final pair)."""
    pairs_aa, pairs_ab, pairs_bb = (None, None, None) if interaction_pairs is None else interaction_pairs
    rng = np.random.default_rng(seed)
    same_spin = _random_symmetric_matrix_normal if diag_coulomb_normal else _random_symmetric_matrix_uniform

    def cross():
        if diag_coulomb_normal:
            return rng.normal(loc=diag_coulomb_mean, scale=diag_coulomb_scale, size=(norb, norb))
        return diag_coulomb_mean + rng.uniform(-0.5 * diag_coulomb_scale, 0.5 * diag_coulomb_scale, size=(norb, norb))

    reps = []
    for _ in range(n_reps):
        mat_aa = same_spin(norb, mean=diag_coulomb_mean, scale=diag_coulomb_scale, seed=rng)
        mat_ab = cross()
        mat_bb = same_spin(norb, mean=diag_coulomb_mean, scale=diag_coulomb_scale, seed=rng)
        reps.append(np.stack([mat_aa, mat_ab, mat_bb]))
    diag_coulomb_mats = np.stack(reps)
    rots = []
    for _ in range(n_reps):
        rot_a = random_unitary(norb, seed=rng)
        rots.append(np.stack([rot_a, random_unitary(norb, seed=rng)]))
    orbital_rotations = np.stack(rots)
    final_orbital_rotation = None
    if with_final_orbital_rotation:
        final_a = random_unitary(norb, seed=rng)
        final_orbital_rotation = np.stack([final_a, random_unitary(norb, seed=rng)])
    for which, pairs, symmetric in ((0, pairs_aa, True), (1, pairs_ab, False), (2, pairs_bb, True)):
        if pairs is not None:
            diag_coulomb_mats[:, which] *= _pair_mask(norb, pairs, symmetric)
    return UCJOpSpinUnbalanced(
        diag_coulomb_mats=diag_coulomb_mats,
        orbital_rotations=orbital_rotations,
        final_orbital_rotation=final_orbital_rotation,
    )


def random_ucj_op_spinless(
    norb: int,
    *,
    n_reps: int = 1,
    interaction_pairs=None,
    with_final_orbital_rotation: bool = False,
    diag_coulomb_mean: float = 0.0,
    diag_coulomb_scale: float = 2 * math.pi,
    diag_coulomb_normal: bool = False,
    seed=None,
) -> UCJOpSpinless:
    """python/ffsim/random/random.py:803-880."""
    rng = np.random.default_rng(seed)
    draw = _random_symmetric_matrix_normal if diag_coulomb_normal else _random_symmetric_matrix_uniform
    diag_coulomb_mats = np.stack(
        [draw(norb, mean=diag_coulomb_mean, scale=diag_coulomb_scale, seed=rng) for _ in range(n_reps)]
    )
    orbital_rotations = np.stack([random_unitary(norb, seed=rng) for _ in range(n_reps)])
    final_orbital_rotation = random_unitary(norb, seed=rng) if with_final_orbital_rotation else None
    if interaction_pairs is not None:
        diag_coulomb_mats *= _pair_mask(norb, interaction_pairs, True)
    return UCJOpSpinless(
        diag_coulomb_mats=diag_coulomb_mats,
        orbital_rotations=orbital_rotations,
        final_orbital_rotation=final_orbital_rotation,
    )


def random_diagonal_coulomb_hamiltonian(norb: int, *, real: bool = False, seed=None) -> DiagonalCoulombHamiltonian:
    rng = np.random.default_rng(seed)
    if real:
        one_body_tensor = random_real_symmetric_matrix(norb, seed=rng)
    else:
        one_body_tensor = random_hermitian(norb, seed=rng)
    diag_coulomb_mat_a = random_real_symmetric_matrix(norb, seed=rng)
    diag_coulomb_mat_b = random_real_symmetric_matrix(norb, seed=rng)
    diag_coulomb_mats = np.stack([diag_coulomb_mat_a, diag_coulomb_mat_b])
    constant = rng.standard_normal()
    return DiagonalCoulombHamiltonian(
        one_body_tensor=one_body_tensor, diag_coulomb_mats=diag_coulomb_mats, constant=constant
    )


def random_double_factorized_hamiltonian(
    norb: int, *, rank: int | None = None, z_representation: bool = False, real: bool = False, seed=None
) -> DoubleFactorizedHamiltonian:
    if rank is None:
        rank = norb * (norb + 1) // 2
    rng = np.random.default_rng(seed)
    if real:
        one_body_tensor = random_real_symmetric_matrix(norb, seed=rng)
        orbital_rotations = np.stack([_random_orthogonal(norb, rng) for _ in range(rank)])
    else:
        one_body_tensor = random_hermitian(norb, seed=rng)
        orbital_rotations = np.stack([random_unitary(norb, seed=rng) for _ in range(rank)])
    diag_coulomb_mats = np.stack([random_real_symmetric_matrix(norb, seed=rng) for _ in range(rank)])
    constant = rng.standard_normal()
    return DoubleFactorizedHamiltonian(
        one_body_tensor=one_body_tensor,
        diag_coulomb_mats=diag_coulomb_mats,
        orbital_rotations=orbital_rotations,
        constant=constant,
        z_representation=z_representation,
    )


def _random_orthogonal(dim: int, rng) -> np.ndarray:
    m = rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(m)
    d = np.diagonal(r)
    return q * (d / np.abs(d))
