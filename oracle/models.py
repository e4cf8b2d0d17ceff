"""Drivers that sequence the hot-path ops (oracle; test infrastructure only)."""

from __future__ import annotations

import cmath
import math

import numpy as np
import scipy.linalg

from oracle.contract import diag_coulomb_matvec, num_op_sum_matvec
from oracle.gates import apply_diag_coulomb_evolution, apply_num_op_sum_evolution, apply_orbital_rotation


def dims(norb, nelec):
    """python/ffsim/states/dimensions.py:18-32."""
    return math.comb(norb, nelec[0]), math.comb(norb, nelec[1])


def dim(norb, nelec):
    """python/ffsim/states/dimensions.py:35-50."""
    if isinstance(nelec, (int, np.integer)):
        return math.comb(norb, int(nelec))
    return math.comb(norb, nelec[0]) * math.comb(norb, nelec[1])


def hartree_fock_state(norb, nelec):
    """python/ffsim/states/slater.py:122-138: one-hot at address 0."""
    vec = np.zeros(dim(norb, nelec), dtype=complex)
    vec[0] = 1
    return vec


def ucj_spin_balanced_apply(vec, diag_coulomb_mats, orbital_rotations, final_orbital_rotation, norb, nelec, copy=True):
    """python/ffsim/variational/ucj_spin_balanced.py:657-696."""
    if copy:
        vec = vec.copy()
    current_basis = np.eye(norb)
    for (mat_aa, mat_ab), orbital_rotation in zip(diag_coulomb_mats, orbital_rotations):
        vec = apply_orbital_rotation(vec, orbital_rotation.T.conj() @ current_basis, norb, nelec, copy=False)
        vec = apply_diag_coulomb_evolution(vec, (mat_aa, mat_ab, mat_aa), -1.0, norb, nelec, copy=False)
        current_basis = orbital_rotation
    if final_orbital_rotation is None:
        return apply_orbital_rotation(vec, current_basis, norb, nelec, copy=False)
    return apply_orbital_rotation(vec, final_orbital_rotation @ current_basis, norb, nelec, copy=False)


def simulate_trotter_step_iterator(n_terms, time, order=0):
    """python/ffsim/trotter/_util.py:18-55."""
    if order == 0:
        for i in range(n_terms):
            yield i, time
    elif order == 1:
        for i in range(n_terms - 1):
            yield i, time / 2
        yield n_terms - 1, time
        for i in reversed(range(n_terms - 1)):
            yield i, time / 2
    else:
        split = time / (4 - 4 ** (1 / (2 * order - 1)))
        for _ in range(2):
            yield from simulate_trotter_step_iterator(n_terms, split, order - 1)
        yield from simulate_trotter_step_iterator(n_terms, time - 4 * split, order - 1)
        for _ in range(2):
            yield from simulate_trotter_step_iterator(n_terms, split, order - 1)


def simulate_trotter_double_factorized(
    vec, one_body_tensor, diag_coulomb_mats, orbital_rotations, constant, z_representation,
    time, *, norb, nelec, n_steps=1, order=0, copy=True,
):
    """python/ffsim/trotter/double_factorized.py:25-127."""
    if order < 0:
        raise ValueError(f"order must be non-negative, got {order}.")
    if n_steps < 0:
        raise ValueError(f"n_steps must be non-negative, got {n_steps}.")
    if copy:
        vec = vec.copy()
    if n_steps == 0:
        return vec
    step_time = time / n_steps
    current_basis = np.eye(norb, dtype=complex)
    for _ in range(n_steps):
        for term, t in simulate_trotter_step_iterator(1 + len(diag_coulomb_mats), step_time, order):
            if term == 0:
                current_basis = scipy.linalg.expm(-1j * t * one_body_tensor) @ current_basis
            else:
                rot = orbital_rotations[term - 1]
                vec = apply_orbital_rotation(vec, rot.T.conj() @ current_basis, norb, nelec, copy=False)
                vec = apply_diag_coulomb_evolution(
                    vec, diag_coulomb_mats[term - 1], t, norb, nelec,
                    z_representation=z_representation, copy=False,
                )
                current_basis = rot
    vec = apply_orbital_rotation(vec, current_basis, norb, nelec, copy=False)
    if constant:
        vec *= cmath.exp(-1j * time * constant)
    return vec


def simulate_trotter_diag_coulomb_split_op(
    vec, one_body_tensor, diag_coulomb_mats, constant, time, *, norb, nelec, n_steps=1, order=0, copy=True
):
    """python/ffsim/trotter/diagonal_coulomb_split_op.py:19-121."""
    if order < 0:
        raise ValueError(f"order must be non-negative, got {order}.")
    if n_steps < 0:
        raise ValueError(f"n_steps must be non-negative, got {n_steps}.")
    if copy:
        vec = vec.copy()
    if n_steps == 0:
        return vec
    step_time = time / n_steps
    mat_aa, mat_ab = diag_coulomb_mats
    current_basis = np.eye(norb, dtype=complex)
    for _ in range(n_steps):
        for term, t in simulate_trotter_step_iterator(2, step_time, order):
            if term == 0:
                current_basis = scipy.linalg.expm(-1j * t * one_body_tensor) @ current_basis
            else:
                vec = apply_orbital_rotation(vec, current_basis, norb, nelec, copy=False)
                vec = apply_diag_coulomb_evolution(vec, (mat_aa, mat_ab, mat_aa), t, norb, nelec, copy=False)
                current_basis = np.eye(norb)
    vec = apply_orbital_rotation(vec, current_basis, norb, nelec, copy=False)
    if constant:
        vec *= cmath.exp(-1j * time * constant)
    return vec


def diagonal_coulomb_hamiltonian_matvec(vec, one_body_tensor, diag_coulomb_mats, constant, norb, nelec):
    """python/ffsim/hamiltonians/diagonal_coulomb_hamiltonian.py:68-95."""
    eigs, vecs = scipy.linalg.eigh(one_body_tensor)
    vec = vec.astype(complex, copy=False)
    result = constant * vec
    result += num_op_sum_matvec(vec, eigs, norb, nelec, orbital_rotation=vecs)
    result += diag_coulomb_matvec(
        vec, (diag_coulomb_mats[0], diag_coulomb_mats[1], diag_coulomb_mats[0]), norb, nelec
    )
    return result


def double_factorized_hamiltonian_matvec(
    vec, one_body_tensor, diag_coulomb_mats, orbital_rotations, constant, z_representation, norb, nelec
):
    """python/ffsim/hamiltonians/double_factorized_hamiltonian.py:244-275."""
    eigs, vecs = scipy.linalg.eigh(one_body_tensor)
    vec = vec.astype(complex, copy=False)
    result = constant * vec
    result += num_op_sum_matvec(vec, eigs, norb, nelec, orbital_rotation=vecs)
    for mat, rot in zip(diag_coulomb_mats, orbital_rotations):
        result += diag_coulomb_matvec(
            vec, mat, norb, nelec, orbital_rotation=rot, z_representation=z_representation
        )
    return result


def ucj_spin_unbalanced_apply(vec, diag_coulomb_mats, orbital_rotations, final_orbital_rotation, norb, nelec,
                              copy=True):
    """python/ffsim/variational/ucj_spin_unbalanced.py:705-742."""
    if copy:
        vec = vec.copy()
    eye = np.eye(norb)
    current_basis = np.stack([eye, eye])
    for mats, rots in zip(diag_coulomb_mats, orbital_rotations):
        u = rots.transpose(0, 2, 1).conj() @ current_basis
        vec = apply_orbital_rotation(vec, (u[0], u[1]), norb, nelec, copy=False)
        vec = apply_diag_coulomb_evolution(vec, (mats[0], mats[1], mats[2]), -1.0, norb, nelec, copy=False)
        current_basis = rots
    last = current_basis if final_orbital_rotation is None else final_orbital_rotation @ current_basis
    return apply_orbital_rotation(vec, (last[0], last[1]), norb, nelec, copy=False)


def ucj_spinless_apply(vec, diag_coulomb_mats, orbital_rotations, final_orbital_rotation, norb, nelec, copy=True):
    """python/ffsim/variational/ucj_spinless.py:456-520 (integer and pair ``nelec``)."""
    if copy:
        vec = vec.copy()
    spinless = isinstance(nelec, (int, np.integer))
    zero = np.zeros((norb, norb))
    current_basis = np.eye(norb)
    for mat, rot in zip(diag_coulomb_mats, orbital_rotations):
        vec = apply_orbital_rotation(vec, rot.T.conj() @ current_basis, norb, nelec, copy=False)
        vec = apply_diag_coulomb_evolution(vec, mat if spinless else (mat, zero, mat), -1.0, norb, nelec, copy=False)
        current_basis = rot
    last = current_basis if final_orbital_rotation is None else final_orbital_rotation @ current_basis
    return apply_orbital_rotation(vec, last, norb, nelec, copy=False)


def qdrift_probabilities(one_body_tensor, diag_coulomb_mats, z_representation, sampling_method, nelec):
    """python/ffsim/trotter/qdrift.py:244-348 ("norm" and "uniform"), :351-455 (norm bounds)."""
    import itertools

    n_terms = 1 + len(diag_coulomb_mats)
    if sampling_method == "uniform":
        return np.ones(n_terms) / n_terms
    assert sampling_method == "norm"

    def norm_one_body(tensor, z_rep=False):
        eigs = scipy.linalg.eigh(tensor, eigvals_only=True)
        n_alpha, n_beta = nelec
        if z_rep:
            return 0.5 * max(
                abs(a + b)
                for a, b in itertools.product(
                    [sum(eigs[n_alpha:]) - sum(eigs[:n_alpha]), sum(eigs[:-n_alpha]) - sum(eigs[-n_alpha:])],
                    [sum(eigs[n_beta:]) - sum(eigs[:n_beta]), sum(eigs[:-n_beta]) - sum(eigs[-n_beta:])],
                )
            )
        return max(
            abs(a + b)
            for a, b in itertools.product(
                [sum(eigs[:n_alpha]), sum(eigs[-n_alpha:])], [sum(eigs[:n_beta]), sum(eigs[-n_beta:])]
            )
        )

    def norm_diag_coulomb(mat):
        eigs, vecs = scipy.linalg.eigh(mat)
        keep = np.abs(eigs) >= 1e-12
        tensors = np.einsum("t,it,ji,ki->tjk", np.emath.sqrt(0.5 * eigs[keep]), vecs[:, keep], np.eye(len(mat)),
                            np.eye(len(mat)))
        if len(tensors) == 1:
            if z_representation:
                bound = norm_one_body(tensors[0], z_rep=True)
                quarter_trace = 0.25 * np.trace(mat)
                return max(quarter_trace, bound**2 - quarter_trace)
            return norm_one_body(tensors[0]) ** 2
        if z_representation:
            return 0.5 * np.sum(np.abs(mat)) - 0.25 * np.sum(np.abs(np.diagonal(mat)))
        return 2 * np.sum(np.abs(mat))

    norms = np.zeros(n_terms)
    if np.all(np.linalg.matrix_rank(diag_coulomb_mats) == 1):
        norms[0] = norm_one_body(one_body_tensor)
    else:
        norms[0] = np.sum(np.abs(scipy.linalg.eigh(one_body_tensor, eigvals_only=True)))
    for i, mat in enumerate(diag_coulomb_mats):
        norms[i + 1] = norm_diag_coulomb(mat)
    return norms / np.sum(norms)


def simulate_qdrift_double_factorized(
    vec, one_body_tensor, diag_coulomb_mats, orbital_rotations, z_representation, time, *, norb, nelec,
    n_steps=1, symmetric=False, probabilities="norm", n_samples=1, seed=None,
):
    """python/ffsim/trotter/qdrift.py:23-241: one rotated evolution per sampled term."""
    initial = vec.copy()
    if n_steps == 0 or time == 0:
        return initial if n_samples == 1 else np.tile(initial, (n_samples, 1))
    if isinstance(probabilities, str):
        probabilities = qdrift_probabilities(one_body_tensor, diag_coulomb_mats, z_representation, probabilities,
                                             nelec)
    probabilities = np.array(probabilities, dtype=float)
    if symmetric:
        probabilities[0] = 0
        probabilities /= sum(probabilities)
    rng = np.random.default_rng(seed)
    energies, basis_change = scipy.linalg.eigh(one_body_tensor)
    step_time = time / n_steps

    def one_body(v, t):
        return apply_num_op_sum_evolution(v, energies, t, norb, nelec, orbital_rotation=basis_change, copy=False)

    def two_body(v, index, t):
        return apply_diag_coulomb_evolution(v, diag_coulomb_mats[index - 1], t, norb, nelec,
                                            orbital_rotation=orbital_rotations[index - 1],
                                            z_representation=z_representation, copy=False)

    results = np.empty((n_samples, initial.shape[0]), dtype=complex)
    for i in range(n_samples):
        v = initial.copy()
        term_indices = rng.choice(len(probabilities), size=n_steps, replace=True, p=probabilities)
        if symmetric:
            v = one_body(v, 0.5 * step_time)
            v = two_body(v, term_indices[0], step_time / probabilities[term_indices[0]])
            for index in term_indices[1:]:
                v = one_body(v, step_time)
                v = two_body(v, index, step_time / probabilities[index])
            v = one_body(v, 0.5 * step_time)
        else:
            for index in term_indices:
                if index == 0:
                    v = one_body(v, step_time / probabilities[0])
                else:
                    v = two_body(v, index, step_time / probabilities[index])
        results[i] = v
    return results[0] if n_samples == 1 else results
