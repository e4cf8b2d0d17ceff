"""Spinless unitary cluster Jastrow operator.

Fields, validation and ``_apply_unitary_`` follow
python/ffsim/variational/ucj_spinless.py:66-120,456-520.  With ``nelec`` an integer the
operator acts on a spinless state; with a pair it acts on both spin sectors with the same
orbital rotations, same-spin interactions only (the alpha-beta matrix is zero).
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
class UCJOpSpinless:
    r"""A spinless UCJ operator :math:`\prod_k \mathcal{U}_k e^{i\mathcal{J}_k}\mathcal{U}_k^\dagger`."""

    diag_coulomb_mats: np.ndarray  # shape: (n_reps, norb, norb)
    orbital_rotations: np.ndarray  # shape: (n_reps, norb, norb)
    final_orbital_rotation: np.ndarray | None = None  # shape: (norb, norb)
    validate: InitVar[bool] = True
    rtol: InitVar[float] = 1e-5
    atol: InitVar[float] = 1e-8

    def __post_init__(self, validate: bool, rtol: float, atol: float):
        if not validate:
            return
        if self.diag_coulomb_mats.ndim != 3:
            raise ValueError(
                "diag_coulomb_mats should have shape (n_reps, norb, norb). "
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
        if not all(linalg.is_real_symmetric(m, rtol=rtol, atol=atol) for m in self.diag_coulomb_mats):
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
        """Number of real parameters (interaction_pairs = a list of upper triangular pairs; None = all)."""
        return _packing.count(norb, n_reps, (_packing.SYM,), None if interaction_pairs is None else (interaction_pairs,), 1,
                              with_final_orbital_rotation)

    @staticmethod
    def from_parameters(params: np.ndarray, *, norb: int, n_reps: int, interaction_pairs=None,
                        with_final_orbital_rotation: bool = False) -> "UCJOpSpinless":
        """Build the operator from a real parameter vector (the reference's layout, see ``_packing``)."""
        mats, rots, final = _packing.unpack(params, norb, n_reps, (_packing.SYM,),
                                            None if interaction_pairs is None else (interaction_pairs,), 1,
                                            with_final_orbital_rotation)
        return UCJOpSpinless(diag_coulomb_mats=mats[:, 0], orbital_rotations=rots[:, 0],
                             final_orbital_rotation=None if final is None else final[0])

    def to_parameters(self, *, interaction_pairs=None) -> np.ndarray:
        """The inverse of ``from_parameters`` (entries outside ``interaction_pairs`` are dropped)."""
        return _packing.pack(self.diag_coulomb_mats[:, None], self.orbital_rotations[:, None], None if self.final_orbital_rotation is None else self.final_orbital_rotation[None], (_packing.SYM,),
                             None if interaction_pairs is None else (interaction_pairs,))

    def _apply_unitary_(self, vec, norb: int, nelec, copy: bool):
        spinless = isinstance(nelec, numbers.Integral)
        pair = (int(nelec), 0) if spinless else (int(nelec[0]), int(nelec[1]))
        t, kind = _device.to_device(vec, copy=copy)
        _check_dim(t, norb, pair)

        def rotate(u):
            _rotate_device(t, u, None if spinless else u, norb, pair)

        basis = np.eye(norb)
        for mat, rot in zip(self.diag_coulomb_mats, self.orbital_rotations):
            rotate(rot.T.conj() @ basis)
            # spinful: same-spin interactions only, i.e. (J, 0, J); spinless: a single sector
            mats = _get_mat_exp(mat if spinless else (mat, None, mat), -1.0, norb, False)
            _evolve_device(t, mats, norb, pair, False)
            basis = rot
        rotate(basis if self.final_orbital_rotation is None else self.final_orbital_rotation @ basis)
        return _device.from_device(t, kind)
