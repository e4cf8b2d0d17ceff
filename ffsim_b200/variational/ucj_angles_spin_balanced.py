"""Spin-balanced UCJ operator parameterised by gate angles
(python/ffsim/variational/ucj_angles_spin_balanced.py:26-283).

Each repetition is a ``GivensAnsatzOp`` (its orbital rotation U_k) and a ``NumNumAnsatzOpSpinBalanced``
(the Jastrow phase): the unitary is prod_k U_k exp(i J_k) U_k^dagger, applied exactly like
``UCJOpSpinBalanced`` -- consecutive rotations merged on the host, n_reps + 1 fused rotations and n_reps
diagonal sweeps on the device.
"""

from __future__ import annotations

import numbers
from dataclasses import dataclass

import numpy as np

from ffsim_b200 import _device
from ffsim_b200.gates.diag_coulomb import _evolve_device, _get_mat_exp
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device
from ffsim_b200.variational._packing import PARAM_MISMATCH
from ffsim_b200.variational.givens import GivensAnsatzOp
from ffsim_b200.variational.num_num import NumNumAnsatzOpSpinBalanced
from ffsim_b200.variational.ucj_spin_balanced import UCJOpSpinBalanced


def brickwork(norb: int, n_layers: int):
    """Pairs of a brickwork circuit: layer i covers (j, j+1) for j = i % 2, i % 2 + 2, ..."""
    for layer in range(n_layers):
        for j in range(layer % 2, norb - 1, 2):
            yield (j, j + 1)


@dataclass(frozen=True)
class UCJAnglesOpSpinBalanced:
    norb: int
    num_num_ansatz_ops: list[NumNumAnsatzOpSpinBalanced]
    givens_ansatz_ops: list[GivensAnsatzOp]
    final_givens_ansatz_op: GivensAnsatzOp | None = None

    def __post_init__(self):
        if len(self.num_num_ansatz_ops) != len(self.givens_ansatz_ops):
            raise ValueError("The number of number-number ansatz operations must equal the number of Givens ansatz "
                             f"operations. Got {len(self.num_num_ansatz_ops)} and {len(self.givens_ansatz_ops)}.")

    @property
    def n_reps(self) -> int:
        return len(self.num_num_ansatz_ops)

    @staticmethod
    def n_params(norb: int, n_reps: int, num_num_interaction_pairs, givens_interaction_pairs,
                 with_final_givens_ansatz_op: bool = False) -> int:
        per_rep = (NumNumAnsatzOpSpinBalanced.n_params(num_num_interaction_pairs)
                   + GivensAnsatzOp.n_params(norb, givens_interaction_pairs))
        return n_reps * per_rep + with_final_givens_ansatz_op * norb**2

    @staticmethod
    def from_parameters(params: np.ndarray, *, norb: int, n_reps: int, num_num_interaction_pairs,
                        givens_interaction_pairs, with_final_givens_ansatz_op: bool = False) -> "UCJAnglesOpSpinBalanced":
        """Per repetition: the Givens operator's parameters, then the number-number angles; at the end the
        final Givens operator on the full norb-layer brickwork (norb**2 parameters)."""
        expected = UCJAnglesOpSpinBalanced.n_params(norb, n_reps, num_num_interaction_pairs, givens_interaction_pairs,
                                                    with_final_givens_ansatz_op)
        if len(params) != expected:
            raise ValueError(PARAM_MISMATCH.format(expected, len(params)))
        n_givens = GivensAnsatzOp.n_params(norb, givens_interaction_pairs)
        n_num_num = NumNumAnsatzOpSpinBalanced.n_params(num_num_interaction_pairs)
        givens_ops, num_num_ops, pos = [], [], 0
        for _ in range(n_reps):
            givens_ops.append(GivensAnsatzOp.from_parameters(params[pos : pos + n_givens], norb=norb,
                                                             interaction_pairs=givens_interaction_pairs))
            pos += n_givens
            num_num_ops.append(NumNumAnsatzOpSpinBalanced.from_parameters(
                params[pos : pos + n_num_num], norb=norb, interaction_pairs=num_num_interaction_pairs))
            pos += n_num_num
        final = None
        if with_final_givens_ansatz_op:
            final = GivensAnsatzOp.from_parameters(params[pos:], norb=norb, interaction_pairs=list(brickwork(norb, norb)))
        return UCJAnglesOpSpinBalanced(norb, num_num_ansatz_ops=num_num_ops, givens_ansatz_ops=givens_ops,
                                       final_givens_ansatz_op=final)

    def to_parameters(self) -> np.ndarray:
        parts = []
        for givens_op, num_num_op in zip(self.givens_ansatz_ops, self.num_num_ansatz_ops):
            parts += [givens_op.to_parameters(), num_num_op.to_parameters()]
        if self.final_givens_ansatz_op is not None:
            parts.append(self.final_givens_ansatz_op.to_parameters())
        return np.concatenate(parts)

    @staticmethod
    def from_ucj_op(ucj_op: UCJOpSpinBalanced) -> "UCJAnglesOpSpinBalanced":
        """Angles of a matrix-based UCJ operator (Givens decomposition of every orbital rotation)."""
        final = None
        if ucj_op.final_orbital_rotation is not None:
            final = GivensAnsatzOp.from_orbital_rotation(ucj_op.final_orbital_rotation)
        return UCJAnglesOpSpinBalanced(
            norb=ucj_op.norb,
            num_num_ansatz_ops=[NumNumAnsatzOpSpinBalanced.from_diag_coulomb_mats(m) for m in ucj_op.diag_coulomb_mats],
            givens_ansatz_ops=[GivensAnsatzOp.from_orbital_rotation(u) for u in ucj_op.orbital_rotations],
            final_givens_ansatz_op=final)

    def _apply_unitary_(self, vec, norb: int, nelec, copy: bool):
        if isinstance(nelec, numbers.Integral):
            return NotImplemented
        nelec = (int(nelec[0]), int(nelec[1]))
        t, kind = _device.to_device(vec, copy=copy)
        _check_dim(t, norb, nelec)
        basis = np.eye(norb)
        for num_num_op, givens_op in zip(self.num_num_ansatz_ops, self.givens_ansatz_ops):
            rotation = givens_op.to_orbital_rotation()
            u = rotation.T.conj() @ basis
            _rotate_device(t, u, u, norb, nelec)
            mat_aa, mat_ab = num_num_op.to_diag_coulomb_mats()
            _evolve_device(t, _get_mat_exp((mat_aa, mat_ab, mat_aa), -1.0, norb, False), norb, nelec, False)
            basis = rotation
        if self.final_givens_ansatz_op is not None:
            basis = self.final_givens_ansatz_op.to_orbital_rotation() @ basis
        _rotate_device(t, basis, basis, norb, nelec)
        return _device.from_device(t, kind)

    def _approx_eq_(self, other, rtol: float, atol: float) -> bool:
        if not isinstance(other, UCJAnglesOpSpinBalanced):
            return NotImplemented
        if self.norb != other.norb or self.n_reps != other.n_reps:
            return False
        if (self.final_givens_ansatz_op is None) != (other.final_givens_ansatz_op is None):
            return False
        if self.final_givens_ansatz_op is not None and not self.final_givens_ansatz_op._approx_eq_(
                other.final_givens_ansatz_op, rtol, atol):
            return False
        return all(a._approx_eq_(b, rtol, atol) for a, b in
                   zip(self.num_num_ansatz_ops + self.givens_ansatz_ops, other.num_num_ansatz_ops + other.givens_ansatz_ops))
