"""Trotter simulation of a double-factorized Hamiltonian:
python/ffsim/trotter/double_factorized.py:25-127."""

from __future__ import annotations

import cmath
import numbers

import numpy as np
import scipy.linalg
import torch

from ffsim_b200 import _device
from ffsim_b200.gates.diag_coulomb import _evolve_device, _get_mat_exp
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device
from ffsim_b200.hamiltonians.diagonal_coulomb_hamiltonian import axpby
from ffsim_b200.hamiltonians.double_factorized_hamiltonian import DoubleFactorizedHamiltonian
from ffsim_b200.trotter._util import simulate_trotter_step_iterator


def simulate_trotter_double_factorized(
    vec,
    hamiltonian: DoubleFactorizedHamiltonian,
    time: float,
    *,
    norb: int,
    nelec: tuple[int, int],
    n_steps: int = 1,
    order: int = 0,
    copy: bool = True,
):
    """Double-factorized Hamiltonian simulation using Trotter-Suzuki formula.

    Arguments, errors and ``copy`` semantics as ``ffsim.simulate_trotter_double_factorized``.
    """
    if order < 0:
        raise ValueError(f"order must be non-negative, got {order}.")
    if n_steps < 0:
        raise ValueError(f"n_steps must be non-negative, got {n_steps}.")
    if isinstance(nelec, numbers.Integral):
        raise TypeError("nelec must be a pair (n_alpha, n_beta)")
    nelec = (int(nelec[0]), int(nelec[1]))
    t, kind = _device.to_device(vec, copy=copy)
    _check_dim(t, norb, nelec)
    if n_steps == 0:
        return _device.from_device(t, kind)

    one_body_tensor = np.asarray(hamiltonian.one_body_tensor)
    step_time = time / n_steps
    current_basis = np.eye(norb, dtype=complex)
    n_terms = 1 + len(hamiltonian.diag_coulomb_mats)
    for _ in range(n_steps):
        for term_index, term_time in simulate_trotter_step_iterator(n_terms, step_time, order):
            if term_index == 0:
                current_basis = scipy.linalg.expm(-1j * term_time * one_body_tensor) @ current_basis
            else:
                rot = np.asarray(hamiltonian.orbital_rotations[term_index - 1])
                u = rot.T.conj() @ current_basis
                _rotate_device(t, u, u, norb, nelec)
                mats = _get_mat_exp(
                    np.asarray(hamiltonian.diag_coulomb_mats[term_index - 1]), term_time, norb,
                    hamiltonian.z_representation,
                )
                _evolve_device(t, mats, norb, nelec, hamiltonian.z_representation)
                current_basis = rot
    _rotate_device(t, current_basis, current_basis, norb, nelec)
    if hamiltonian.constant:
        phase = cmath.exp(-1j * time * hamiltonian.constant)
        axpby(phase, t, 0.0, t)
    return _device.from_device(t, kind)
