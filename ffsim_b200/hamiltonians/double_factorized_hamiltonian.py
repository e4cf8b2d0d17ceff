"""Double-factorized Hamiltonian container: fields of
python/ffsim/hamiltonians/double_factorized_hamiltonian.py:90-100 and its
``_linear_operator_`` (:244-275).  The factorisation itself
(``from_molecular_hamiltonian``) is Hamiltonian pre-processing and out of scope."""

from __future__ import annotations

import dataclasses

import numpy as np
import scipy.linalg
import torch

from ffsim_b200 import _device
from ffsim_b200.contract.diag_coulomb import _contract_device as _contract_dc
from ffsim_b200.contract.diag_coulomb import _get_mats
from ffsim_b200.contract.linop import DeviceLinearOperator
from ffsim_b200.contract.num_op_sum import _contract_device as _contract_num
from ffsim_b200.gates.orbital_rotation import _rotate_device
from ffsim_b200.hamiltonians.diagonal_coulomb_hamiltonian import axpby
from ffsim_b200.states import dim


@dataclasses.dataclass(frozen=True)
class DoubleFactorizedHamiltonian:
    r""":math:`H = \sum_{pq\sigma}\kappa_{pq} a^\dagger_{p\sigma} a_{q\sigma}
    + \frac12 \sum_t \sum_{ij\sigma\tau} J^{(t)}_{ij} n^{(t)}_{i\sigma} n^{(t)}_{j\tau} + \text{constant}`."""

    one_body_tensor: np.ndarray
    diag_coulomb_mats: np.ndarray  # (L, norb, norb)
    orbital_rotations: np.ndarray  # (L, norb, norb)
    constant: float = 0.0
    z_representation: bool = False

    @property
    def norb(self) -> int:
        return self.one_body_tensor.shape[0]

    def _linear_operator_(self, norb: int, nelec) -> DeviceLinearOperator:
        assert isinstance(nelec, tuple)
        nelec = (int(nelec[0]), int(nelec[1]))
        eigs, vecs = scipy.linalg.eigh(self.one_body_tensor)
        eigs = np.ascontiguousarray(eigs, dtype=float)
        vecs_dag = vecs.T.conj()
        terms = [
            (_get_mats(np.asarray(mat), norb, self.z_representation), np.asarray(rot))
            for mat, rot in zip(self.diag_coulomb_mats, self.orbital_rotations)
        ]
        constant, z_rep = self.constant, self.z_representation

        def matvec(t):
            work = t.clone()
            _rotate_device(work, vecs_dag, vecs_dag, norb, nelec)
            result = _device.empty_like(t)
            _contract_num(work, result, eigs, norb, nelec, accumulate=False)
            _rotate_device(result, vecs, vecs, norb, nelec)
            tmp = _device.empty_like(t)
            for mats, rot in terms:
                work.copy_(t)
                rot_dag = rot.T.conj()
                _rotate_device(work, rot_dag, rot_dag, norb, nelec)
                _contract_dc(work, tmp, mats, norb, nelec, z_rep, accumulate=False)
                _rotate_device(tmp, rot, rot, norb, nelec)
                axpby(1.0, tmp, 1.0, result)
            if constant:
                axpby(constant, t, 1.0, result)
            return result

        return DeviceLinearOperator(dim(norb, nelec), matvec)
