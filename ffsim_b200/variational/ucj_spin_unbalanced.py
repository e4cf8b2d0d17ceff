"""Spin-unbalanced unitary cluster Jastrow operator.

Fields, validation and ``_apply_unitary_`` follow
python/ffsim/variational/ucj_spin_unbalanced.py:66-130,705-742: separate alpha and beta
orbital rotations ``(n_reps, 2, norb, norb)`` and three diagonal Coulomb matrices per
repetition ``(n_reps, 3, norb, norb)`` (alpha-alpha, alpha-beta, beta-beta; the alpha-beta
matrix need not be symmetric).  The pyscf-backed constructors are out of scope.
"""

from __future__ import annotations

import numbers
from dataclasses import InitVar, dataclass

import numpy as np

from ffsim_b200 import _device, linalg
from ffsim_b200.variational import _packing
from ffsim_b200.gates.diag_coulomb import _evolve_device, _get_mat_exp
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device


@dataclass(frozen=True)
class UCJOpSpinUnbalanced:
    r"""A spin-unbalanced UCJ operator :math:`\prod_k \mathcal{U}_k e^{i\mathcal{J}_k}\mathcal{U}_k^\dagger`."""

    diag_coulomb_mats: np.ndarray  # shape: (n_reps, 3, norb, norb)
    orbital_rotations: np.ndarray  # shape: (n_reps, 2, norb, norb)
    final_orbital_rotation: np.ndarray | None = None  # shape: (2, norb, norb)
    validate: InitVar[bool] = True
    rtol: InitVar[float] = 1e-5
    atol: InitVar[float] = 1e-8

    def __post_init__(self, validate: bool, rtol: float, atol: float):
        if not validate:
            return
        if self.diag_coulomb_mats.ndim != 4 or self.diag_coulomb_mats.shape[1] != 3:
            raise ValueError(
                "diag_coulomb_mats should have shape (n_reps, 3, norb, norb). "
                f"Got shape {self.diag_coulomb_mats.shape}."
            )
        if self.orbital_rotations.ndim != 4 or self.orbital_rotations.shape[1] != 2:
            raise ValueError(
                "orbital_rotations should have shape (n_reps, 2, norb, norb). "
                f"Got shape {self.orbital_rotations.shape}."
            )
        if self.final_orbital_rotation is not None and self.final_orbital_rotation.ndim != 3:
            raise ValueError(
                "final_orbital_rotation should have shape (2, norb, norb). "
                f"Got shape {self.final_orbital_rotation.shape}."
            )
        if self.diag_coulomb_mats.shape[0] != self.orbital_rotations.shape[0]:
            raise ValueError(
                "diag_coulomb_mats and orbital_rotations should have the same first dimension. "
                f"Got {self.diag_coulomb_mats.shape[0]} and {self.orbital_rotations.shape[0]}."
            )
        for mats in self.diag_coulomb_mats:
            if not (linalg.is_real_symmetric(mats[0], rtol=rtol, atol=atol)
                    and linalg.is_real_symmetric(mats[2], rtol=rtol, atol=atol)):
                raise ValueError(
                    "alpha-alpha and beta-beta diagonal Coulomb matrices were not all real symmetric."
                )
        for pair in self.orbital_rotations:
            if not (linalg.is_unitary(pair[0], rtol=rtol, atol=atol)
                    and linalg.is_unitary(pair[1], rtol=rtol, atol=atol)):
                raise ValueError("Orbital rotations were not all unitary.")
        if self.final_orbital_rotation is not None and not (
            linalg.is_unitary(self.final_orbital_rotation[0], rtol=rtol, atol=atol)
            and linalg.is_unitary(self.final_orbital_rotation[1], rtol=rtol, atol=atol)
        ):
            raise ValueError("Final orbital rotation was not unitary.")

    @property
    def norb(self) -> int:
        return self.diag_coulomb_mats.shape[-1]

    @property
    def n_reps(self) -> int:
        return self.diag_coulomb_mats.shape[0]

    @staticmethod
    def n_params(norb: int, n_reps: int, *, interaction_pairs=None, with_final_orbital_rotation: bool = False) -> int:
        """Number of real parameters (interaction_pairs = (alpha-alpha, alpha-beta, beta-beta); the alpha-beta pairs are ordered)."""
        return _packing.count(norb, n_reps, (_packing.SYM, _packing.FULL, _packing.SYM), interaction_pairs, 2, with_final_orbital_rotation)

    @staticmethod
    def from_parameters(params: np.ndarray, *, norb: int, n_reps: int, interaction_pairs=None,
                        with_final_orbital_rotation: bool = False) -> "UCJOpSpinUnbalanced":
        """Build the operator from a real parameter vector (the reference's layout, see ``_packing``)."""
        mats, rots, final = _packing.unpack(params, norb, n_reps, (_packing.SYM, _packing.FULL, _packing.SYM), interaction_pairs, 2,
                                            with_final_orbital_rotation)
        return UCJOpSpinUnbalanced(diag_coulomb_mats=mats, orbital_rotations=rots, final_orbital_rotation=final)

    def to_parameters(self, *, interaction_pairs=None) -> np.ndarray:
        """The inverse of ``from_parameters`` (entries outside ``interaction_pairs`` are dropped)."""
        return _packing.pack(self.diag_coulomb_mats, self.orbital_rotations, self.final_orbital_rotation, (_packing.SYM, _packing.FULL, _packing.SYM), interaction_pairs)

    def _apply_unitary_(self, vec, norb: int, nelec, copy: bool):
        if isinstance(nelec, numbers.Integral):
            return NotImplemented
        nelec = (int(nelec[0]), int(nelec[1]))
        t, kind = _device.to_device(vec, copy=copy)
        _check_dim(t, norb, nelec)
        basis_a = basis_b = np.eye(norb)
        for (mat_aa, mat_ab, mat_bb), (rot_a, rot_b) in zip(self.diag_coulomb_mats, self.orbital_rotations):
            # leave the previous repetition's basis and enter this one's in a single rotation per spin
            _rotate_device(t, rot_a.T.conj() @ basis_a, rot_b.T.conj() @ basis_b, norb, nelec)
            _evolve_device(t, _get_mat_exp((mat_aa, mat_ab, mat_bb), -1.0, norb, False), norb, nelec, False)
            basis_a, basis_b = rot_a, rot_b
        if self.final_orbital_rotation is not None:
            basis_a = self.final_orbital_rotation[0] @ basis_a
            basis_b = self.final_orbital_rotation[1] @ basis_b
        _rotate_device(t, basis_a, basis_b, norb, nelec)
        return _device.from_device(t, kind)
