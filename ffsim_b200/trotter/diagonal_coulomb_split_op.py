"""Split-operator Trotter simulation of a diagonal Coulomb Hamiltonian:
python/ffsim/trotter/diagonal_coulomb_split_op.py:19-121."""

from __future__ import annotations

import cmath

import numpy as np
import scipy.linalg

from ffsim_b200 import _device
from ffsim_b200.gates.diag_coulomb import _evolve_device, _get_mat_exp
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device
from ffsim_b200.hamiltonians.diagonal_coulomb_hamiltonian import DiagonalCoulombHamiltonian, axpby
from ffsim_b200.trotter._util import simulate_trotter_step_iterator


def simulate_trotter_diag_coulomb_split_op(
    vec,
    hamiltonian: DiagonalCoulombHamiltonian,
    time: float,
    *,
    norb: int,
    nelec: tuple[int, int],
    n_steps: int = 1,
    order: int = 0,
    copy: bool = True,
):
    """Diagonal Coulomb Hamiltonian simulation using the split-operator method."""
    if order < 0:
        raise ValueError(f"order must be non-negative, got {order}.")
    if n_steps < 0:
        raise ValueError(f"n_steps must be non-negative, got {n_steps}.")
    nelec = (int(nelec[0]), int(nelec[1]))
    t, kind = _device.to_device(vec, copy=copy)
    _check_dim(t, norb, nelec)
    if n_steps == 0:
        return _device.from_device(t, kind)
    one_body_tensor = np.asarray(hamiltonian.one_body_tensor)
    mat_aa, mat_ab = hamiltonian.diag_coulomb_mats
    step_time = time / n_steps
    current_basis = np.eye(norb, dtype=complex)
    for _ in range(n_steps):
        for term_index, term_time in simulate_trotter_step_iterator(2, step_time, order):
            if term_index == 0:
                current_basis = scipy.linalg.expm(-1j * term_time * one_body_tensor) @ current_basis
            else:
                _rotate_device(t, current_basis, current_basis, norb, nelec)
                mats = _get_mat_exp((mat_aa, mat_ab, mat_aa), term_time, norb, False)
                _evolve_device(t, mats, norb, nelec, False)
                current_basis = np.eye(norb, dtype=complex)
    _rotate_device(t, current_basis, current_basis, norb, nelec)
    if hamiltonian.constant:
        axpby(cmath.exp(-1j * time * hamiltonian.constant), t, 0.0, t)
    return _device.from_device(t, kind)
