"""Spin-balanced unitary cluster Jastrow operator (LUCJ when the interaction pairs are local).

Fields, validation and ``_apply_unitary_`` follow
python/ffsim/variational/ucj_spin_balanced.py:31-132,657-696.  The pyscf-backed
constructors (``from_t_amplitudes``, ``from_cisd_vec``) are out of scope.
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
class UCJOpSpinBalanced:
    r"""A spin-balanced UCJ operator :math:`\prod_k \mathcal{U}_k e^{i\mathcal{J}_k}\mathcal{U}_k^\dagger`."""

    diag_coulomb_mats: np.ndarray  # shape: (n_reps, 2, norb, norb)
    orbital_rotations: np.ndarray  # shape: (n_reps, norb, norb)
    final_orbital_rotation: np.ndarray | None = None  # shape: (norb, norb)
    validate: InitVar[bool] = True
    rtol: InitVar[float] = 1e-5
    atol: InitVar[float] = 1e-8

    def __post_init__(self, validate: bool, rtol: float, atol: float):
        if validate:
            if self.diag_coulomb_mats.ndim != 4 or self.diag_coulomb_mats.shape[1] != 2:
                raise ValueError(
                    "diag_coulomb_mats should have shape (n_reps, 2, norb, norb). "
                    f"Got shape {self.diag_coulomb_mats.shape}."
                )
            if self.orbital_rotations.ndim != 3:
                raise ValueError(
                    "orbital_rotations should have shape (n_reps, norb, norb). "
                    f"Got shape {self.orbital_rotations.shape}."
                )
            if self.final_orbital_rotation is not None and self.final_orbital_rotation.ndim != 2:
                raise ValueError(
                    "final_orbital_rotation should have shape (norb, norb). "
                    f"Got shape {self.final_orbital_rotation.shape}."
                )
            if self.diag_coulomb_mats.shape[0] != self.orbital_rotations.shape[0]:
                raise ValueError(
                    "diag_coulomb_mats and orbital_rotations should have the same first dimension. "
                    f"Got {self.diag_coulomb_mats.shape[0]} and {self.orbital_rotations.shape[0]}."
                )
            if not all(
                linalg.is_real_symmetric(mats[0], rtol=rtol, atol=atol)
                and linalg.is_real_symmetric(mats[1], rtol=rtol, atol=atol)
                for mats in self.diag_coulomb_mats
            ):
                raise ValueError("Diagonal Coulomb matrices were not all real symmetric.")
            if not all(linalg.is_unitary(u, rtol=rtol, atol=atol) for u in self.orbital_rotations):
                raise ValueError("Orbital rotations were not all unitary.")
            if self.final_orbital_rotation is not None and not linalg.is_unitary(
                self.final_orbital_rotation, rtol=rtol, atol=atol
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
        """Number of real parameters (interaction_pairs = (alpha-alpha pairs, alpha-beta pairs), upper triangular; None = all)."""
        return _packing.count(norb, n_reps, (_packing.SYM, _packing.SYM), interaction_pairs, 1, with_final_orbital_rotation)

    @staticmethod
    def from_parameters(params: np.ndarray, *, norb: int, n_reps: int, interaction_pairs=None,
                        with_final_orbital_rotation: bool = False) -> "UCJOpSpinBalanced":
        """Build the operator from a real parameter vector (the reference's layout, see ``_packing``)."""
        mats, rots, final = _packing.unpack(params, norb, n_reps, (_packing.SYM, _packing.SYM), interaction_pairs, 1,
                                            with_final_orbital_rotation)
        return UCJOpSpinBalanced(diag_coulomb_mats=mats, orbital_rotations=rots[:, 0],
                                 final_orbital_rotation=None if final is None else final[0])

    def to_parameters(self, *, interaction_pairs=None) -> np.ndarray:
        """The inverse of ``from_parameters`` (entries outside ``interaction_pairs`` are dropped)."""
        return _packing.pack(self.diag_coulomb_mats, self.orbital_rotations[:, None], None if self.final_orbital_rotation is None else self.final_orbital_rotation[None], (_packing.SYM, _packing.SYM), interaction_pairs)

    def _apply_unitary_(self, vec, norb: int, nelec, copy: bool):
        if isinstance(nelec, numbers.Integral):
            return NotImplemented
        nelec = (int(nelec[0]), int(nelec[1]))
        t, kind = _device.to_device(vec, copy=copy)
        _check_dim(t, norb, nelec)
        current_basis = np.eye(norb)
        for (mat_aa, mat_ab), orbital_rotation in zip(self.diag_coulomb_mats, self.orbital_rotations):
            u = orbital_rotation.T.conj() @ current_basis
            _rotate_device(t, u, u, norb, nelec)
            mats = _get_mat_exp((mat_aa, mat_ab, mat_aa), -1.0, norb, False)
            _evolve_device(t, mats, norb, nelec, False)
            current_basis = orbital_rotation
        u = current_basis if self.final_orbital_rotation is None else self.final_orbital_rotation @ current_basis
        _rotate_device(t, u, u, norb, nelec)
        return _device.from_device(t, kind)
