"""qDRIFT simulation of a double-factorized Hamiltonian:
python/ffsim/trotter/qdrift.py:23-241 (driver), :244-348 (sampling probabilities),
:351-455 (spectral-norm bounds).

Every sampled term is one rotated number-operator-sum or diagonal-Coulomb evolution on the
device; consecutive basis changes are merged into a single orbital rotation, as in
``simulate_trotter_double_factorized``.  The state-dependent "optimal" / "optimal-incoherent"
probabilities (qdrift.py:304-343) use the Wick expectation values of ``ffsim_b200/states/wick.py`` (host).
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
from ffsim_b200.states.wick import expectation_one_body_power, expectation_one_body_product


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
        one_body = scipy.linalg.block_diag(hamiltonian.one_body_tensor, hamiltonian.one_body_tensor)
        incoherent = sampling_method == "optimal-incoherent"
        weights = np.zeros(n_terms)
        weights[0] = (expectation_one_body_power(one_rdm, one_body, 2).real if incoherent
                      else variance_one_body_tensor(one_rdm, one_body))
        for i, mat in enumerate(hamiltonian.diag_coulomb_mats):
            moments = _diag_coulomb_moments(one_rdm, mat, hamiltonian.orbital_rotations[i],
                                            hamiltonian.z_representation)
            weights[i + 1] = moments[1].real if incoherent else max(0, (moments[1] - abs(moments[0]) ** 2).real)
        stds = np.sqrt(weights)
        return stds / np.sum(stds)
    raise ValueError(f"Unsupported sampling method: {sampling_method}.")


def one_body_square_decomposition(diag_coulomb_mat: np.ndarray, orbital_rotation: np.ndarray | None = None,
                                  truncation_threshold: float = 1e-12) -> np.ndarray:
    """One-body matrices whose squares sum to the two-body term: sqrt(lambda_t / 2) U diag(v_t) U^dagger
    for every eigenpair of the diagonal Coulomb matrix (qdrift.py:432-460)."""
    if orbital_rotation is None:
        orbital_rotation = np.eye(diag_coulomb_mat.shape[0])
    eigs, vecs = scipy.linalg.eigh(diag_coulomb_mat)
    keep = np.abs(eigs) >= truncation_threshold
    eigs, vecs = eigs[keep], vecs[:, keep]
    scale = np.emath.sqrt(0.5 * eigs)
    return np.stack([scale[t] * (orbital_rotation * vecs[:, t]) @ orbital_rotation.T.conj() for t in range(len(eigs))]) \
        if len(eigs) else np.zeros((0,) + diag_coulomb_mat.shape, dtype=complex)


def variance_one_body_tensor(one_rdm: np.ndarray, one_body_tensor: np.ndarray) -> float:
    """Variance of a one-body operator in a Slater determinant (qdrift.py:463-480)."""
    var = (expectation_one_body_power(one_rdm, one_body_tensor, 2)
           - abs(expectation_one_body_power(one_rdm, one_body_tensor, 1)) ** 2).real
    return max(0, var)


def _diag_coulomb_moments(one_rdm, diag_coulomb_mat, orbital_rotation, z_representation):
    """(<T>, <T^2>) of a rotated diagonal Coulomb term T = sum_t O_t^2 in a Slater determinant
    (qdrift.py:483-613).  In the Z representation the term carries a one-body correction; the reference
    adds its cross terms with the LAST square root only (``one_body_op`` is the loop variable of the
    preceding loop there) -- kept, since these numbers only steer sampling probabilities and must agree."""
    if orbital_rotation is None:
        orbital_rotation = np.eye(diag_coulomb_mat.shape[0])
    ops = [scipy.linalg.block_diag(m, m) for m in one_body_square_decomposition(diag_coulomb_mat, orbital_rotation)]
    first = sum((expectation_one_body_power(one_rdm, op, 2) for op in ops), 0j)
    second = sum((expectation_one_body_power(one_rdm, op, 4) for op in ops), 0j)
    for op1, op2 in itertools.combinations(ops, 2):
        second += 2 * expectation_one_body_product(one_rdm, [op1, op1, op2, op2])
    if z_representation:
        rot, rot_c = orbital_rotation, orbital_rotation.conj()
        corr = -0.5 * (np.einsum("ij,pi,qi->pq", diag_coulomb_mat, rot, rot_c)
                       + np.einsum("ij,pj,qj->pq", diag_coulomb_mat, rot, rot_c))
        corr = scipy.linalg.block_diag(corr, corr)
        first += expectation_one_body_power(one_rdm, corr, 1)
        second += expectation_one_body_power(one_rdm, corr, 2)
        if ops:
            second += expectation_one_body_product(one_rdm, [corr, ops[-1], ops[-1]])
            second += expectation_one_body_product(one_rdm, [ops[-1], ops[-1], corr])
    return first, second


def variance_diag_coulomb(one_rdm, diag_coulomb_mat, orbital_rotation=None, z_representation: bool = False) -> float:
    first, second = _diag_coulomb_moments(one_rdm, diag_coulomb_mat, orbital_rotation, z_representation)
    return max(0, (second - abs(first) ** 2).real)


def expectation_squared_diag_coulomb(one_rdm, diag_coulomb_mat, orbital_rotation=None,
                                     z_representation: bool = False) -> float:
    return _diag_coulomb_moments(one_rdm, diag_coulomb_mat, orbital_rotation, z_representation)[1].real


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
