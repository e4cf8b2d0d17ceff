"""qDRIFT simulation of a double-factorized Hamiltonian:
python/ffsim/trotter/qdrift.py:23-241 (driver), :244-348 (sampling probabilities),
:351-455 (spectral-norm bounds).

Every sampled term is one rotated number-operator-sum or diagonal-Coulomb evolution on the
device; consecutive basis changes are merged into a single orbital rotation, as in
``simulate_trotter_double_factorized``.  The state-dependent "optimal" probabilities need the
Wick-expectation machinery of python/ffsim/states/wick.py, which is outside the hot path: pass an
explicit probability array instead.
"""

from __future__ import annotations

import itertools
import numbers

import numpy as np
import scipy.linalg
import torch

from ffsim_b200 import _device
from ffsim_b200.gates.diag_coulomb import _evolve_device, _get_mat_exp
from ffsim_b200.gates.num_op_sum import _evolve_device as _evolve_num_op_sum
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device
from ffsim_b200.hamiltonians.double_factorized_hamiltonian import DoubleFactorizedHamiltonian


def spectral_norm_one_body_tensor(one_body_tensor, *, nelec, z_representation: bool = False) -> float:
    """Upper bound on the largest singular value of a one-body operator (qdrift.py:351-388)."""
    eigs = scipy.linalg.eigh(one_body_tensor, eigvals_only=True)
    n_alpha, n_beta = nelec

    def extremes(n):
        if z_representation:
            return [sum(eigs[n:]) - sum(eigs[:n]), sum(eigs[:-n]) - sum(eigs[-n:])]
        return [sum(eigs[:n]), sum(eigs[-n:])]

    bound = max(abs(a + b) for a, b in itertools.product(extremes(n_alpha), extremes(n_beta)))
    return 0.5 * bound if z_representation else bound


def one_body_square_decomposition(diag_coulomb_mat, orbital_rotation=None, truncation_threshold: float = 1e-12):
    """A two-body term as a sum of squared one-body operators (qdrift.py:429-454)."""
    if orbital_rotation is None:
        orbital_rotation = np.eye(diag_coulomb_mat.shape[0])
    eigs, vecs = scipy.linalg.eigh(diag_coulomb_mat)
    keep = np.abs(eigs) >= truncation_threshold
    eigs, vecs = eigs[keep], vecs[:, keep]
    return np.einsum("t,it,ji,ki->tjk", np.emath.sqrt(0.5 * eigs), vecs, orbital_rotation, orbital_rotation.conj())


def spectral_norm_diag_coulomb(diag_coulomb_mat, nelec, z_representation: bool = False) -> float:
    """Upper bound on the largest singular value of a diagonal Coulomb operator (qdrift.py:391-426)."""
    squares = one_body_square_decomposition(diag_coulomb_mat)
    if len(squares) == 1:  # rank one: exact in the number representation
        if z_representation:
            bound = spectral_norm_one_body_tensor(squares[0], nelec=nelec, z_representation=True)
            quarter_trace = 0.25 * np.trace(diag_coulomb_mat)
            return max(quarter_trace, bound**2 - quarter_trace)
        return spectral_norm_one_body_tensor(squares[0], nelec=nelec) ** 2
    if z_representation:
        return 0.5 * np.sum(np.abs(diag_coulomb_mat)) - 0.25 * np.sum(np.abs(np.diagonal(diag_coulomb_mat)))
    return 2 * np.sum(np.abs(diag_coulomb_mat))


def qdrift_probabilities(hamiltonian: DoubleFactorizedHamiltonian, sampling_method: str, *, nelec=None,
                         one_rdm=None) -> np.ndarray:
    """Sampling probabilities, the one-body term first (qdrift.py:244-348)."""
    n_terms = 1 + len(hamiltonian.diag_coulomb_mats)
    if sampling_method == "norm":
        if nelec is None:
            raise ValueError("The 'norm' sampling method requires nelec to be specified.")
        norms = np.zeros(n_terms)
        if np.all(np.linalg.matrix_rank(hamiltonian.diag_coulomb_mats) == 1):
            norms[0] = spectral_norm_one_body_tensor(hamiltonian.one_body_tensor, nelec=nelec)
        else:  # loose bounds on the two-body terms: use a loose one here too
            norms[0] = np.sum(np.abs(scipy.linalg.eigh(hamiltonian.one_body_tensor, eigvals_only=True)))
        for i, mat in enumerate(hamiltonian.diag_coulomb_mats):
            norms[i + 1] = spectral_norm_diag_coulomb(mat, z_representation=hamiltonian.z_representation,
                                                       nelec=nelec)
        return norms / np.sum(norms)
    if sampling_method == "uniform":
        return np.ones(n_terms) / n_terms
    if sampling_method in ("optimal", "optimal-incoherent"):
        if one_rdm is None:
            raise ValueError(f"The '{sampling_method}' sampling method requires one_rdm to be specified.")
        raise NotImplementedError(
            f"sampling method '{sampling_method}' needs the Wick expectation values of ffsim.states.wick, "
            "which are outside this package; pass the probabilities as an array"
        )
    raise ValueError(f"Unsupported sampling method: {sampling_method}.")


def simulate_qdrift_double_factorized(
    vec,
    hamiltonian: DoubleFactorizedHamiltonian,
    time: float,
    *,
    norb: int,
    nelec: tuple[int, int],
    n_steps: int = 1,
    symmetric: bool = False,
    probabilities="norm",
    one_rdm=None,
    n_samples: int = 1,
    seed=None,
):
    """Double-factorized Hamiltonian simulation via qDRIFT.

    Arguments, errors, sampling order and return shape as ``ffsim.simulate_qdrift_double_factorized``:
    a vector for ``n_samples == 1``, otherwise an array of shape ``(n_samples, dim)``.  ``vec`` may be
    a NumPy array or a CUDA tensor (the result is of the same kind); it is never modified.
    """
    if n_steps < 0:
        raise ValueError(f"n_steps must be non-negative, got {n_steps}.")
    if n_samples < 1:
        raise ValueError(f"n_samples must be positive, got {n_samples}.")
    if isinstance(nelec, numbers.Integral):
        raise TypeError("nelec must be a pair (n_alpha, n_beta)")
    nelec = (int(nelec[0]), int(nelec[1]))
    initial, kind = _device.to_device(vec, copy=True)
    if kind.sharded:
        raise NotImplementedError("qDRIFT trajectories are independent: run one per rank instead of sharding")
    _check_dim(initial, norb, nelec)

    def finish(samples):
        if n_samples == 1:
            return _device.from_device(samples[0], kind)
        stacked = torch.stack(samples)
        if kind.numpy:
            return stacked.cpu().numpy()
        return stacked.cpu() if kind.torch_cpu else stacked

    if n_steps == 0 or time == 0:
        return finish([initial.clone() for _ in range(n_samples)])

    if isinstance(probabilities, str):
        probabilities = qdrift_probabilities(hamiltonian, sampling_method=probabilities, nelec=nelec,
                                             one_rdm=one_rdm)
    probabilities = np.array(probabilities, dtype=float)
    if symmetric:  # the one-body term is applied deterministically between the sampled terms
        probabilities[0] = 0
        probabilities /= sum(probabilities)

    rng = np.random.default_rng(seed)
    energies, basis_change = scipy.linalg.eigh(hamiltonian.one_body_tensor)
    energies = np.ascontiguousarray(energies, dtype=float)
    step_time = time / n_steps
    z_rep = hamiltonian.z_representation
    eye = np.eye(norb, dtype=complex)

    samples = []
    for _ in range(n_samples):
        t = initial.clone()
        term_indices = rng.choice(len(probabilities), size=n_steps, replace=True, p=probabilities)
        basis, basis_id = eye, -1  # the orbital basis the device vector is currently expressed in

        def to_basis(new_basis, new_id):
            # bases are tracked by term index (-1 identity, 0 one-body eigenbasis, k two-body term k):
            # two consecutive samples of the same term must not cost a rotation by the identity
            nonlocal basis, basis_id
            if new_id != basis_id:
                u = new_basis.T.conj() @ basis
                _rotate_device(t, u, u, norb, nelec)
                basis, basis_id = new_basis, new_id

        def one_body(term_time):
            to_basis(basis_change, 0)
            phases = np.ascontiguousarray(np.exp(-1j * term_time * energies))
            _evolve_num_op_sum(t, phases, phases, norb, nelec)

        def two_body(index, term_time):
            to_basis(np.asarray(hamiltonian.orbital_rotations[index - 1]), int(index))
            mats = _get_mat_exp(np.asarray(hamiltonian.diag_coulomb_mats[index - 1]), term_time, norb, z_rep)
            _evolve_device(t, mats, norb, nelec, z_rep)

        if symmetric:
            one_body(0.5 * step_time)
            two_body(term_indices[0], step_time / probabilities[term_indices[0]])
            for index in term_indices[1:]:
                one_body(step_time)
                two_body(index, step_time / probabilities[index])
            one_body(0.5 * step_time)
        else:
            for index in term_indices:
                if index == 0:
                    one_body(step_time / probabilities[0])
                else:
                    two_body(index, step_time / probabilities[index])
        to_basis(eye, -1)
        samples.append(t)
    return finish(samples)
