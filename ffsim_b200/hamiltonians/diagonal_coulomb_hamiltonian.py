"""Diagonal Coulomb Hamiltonian: fields and ``_linear_operator_`` of
python/ffsim/hamiltonians/diagonal_coulomb_hamiltonian.py:37-95."""

from __future__ import annotations

import dataclasses

import numpy as np
import scipy.linalg
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.contract.diag_coulomb import _contract_device as _contract_dc
from ffsim_b200.contract.diag_coulomb import _get_mats
from ffsim_b200.contract.linop import DeviceLinearOperator
from ffsim_b200.contract.num_op_sum import _contract_device as _contract_num
from ffsim_b200.gates.orbital_rotation import _rotate_device
from ffsim_b200.states import dim


def axpby(alpha: complex, x, beta: complex, y) -> None:
    """y = alpha * x + beta * y on the device (tensors or ShardedVectors)."""
    if _device.is_sharded(x):
        _device.same_layout(x, y)  # y follows x's distribution (a redistribution only when they differ)
        x, y = x.local, y.local
    if x.numel() == 0:
        return
    with torch.cuda.device(x.device):
        _device.sync_device()
        _lib.check(
            _lib.lib.ffb_axpby(_lib.c128(alpha), x.data_ptr(), _lib.c128(beta), y.data_ptr(), x.numel(),
                               _device.stream_ptr())
        )


@dataclasses.dataclass(frozen=True)
class DiagonalCoulombHamiltonian:
    r""":math:`H = \sum_{pq\sigma} h_{pq} a^\dagger_{p\sigma} a_{q\sigma}
    + \frac12 \sum_{pq\sigma\tau} J^{\sigma\tau}_{pq} n_{p\sigma} n_{q\tau} + \text{constant}`."""

    one_body_tensor: np.ndarray
    diag_coulomb_mats: np.ndarray  # (2, norb, norb): alpha-alpha, alpha-beta
    constant: float = 0.0

    @property
    def norb(self) -> int:
        return self.one_body_tensor.shape[0]

    def _linear_operator_(self, norb: int, nelec) -> DeviceLinearOperator:
        assert isinstance(nelec, tuple)
        nelec = (int(nelec[0]), int(nelec[1]))
        eigs, vecs = scipy.linalg.eigh(self.one_body_tensor)
        eigs = np.ascontiguousarray(eigs, dtype=float)
        vecs_dag = vecs.T.conj()
        dc_mats = _get_mats(
            (self.diag_coulomb_mats[0], self.diag_coulomb_mats[1], self.diag_coulomb_mats[0]), norb, False
        )
        constant = self.constant

        def matvec(t):
            # num_linop @ vec: rotate into the eigenbasis of h, contract (in place: the operator is
            # diagonal there), rotate back -- one state-sized temporary besides the result
            result = t.clone()
            _rotate_device(result, vecs_dag, vecs_dag, norb, nelec)
            _contract_num(result, result, eigs, norb, nelec, accumulate=False)
            _rotate_device(result, vecs, vecs, norb, nelec)
            # + dc_linop @ vec (accumulate form) + constant * vec
            _contract_dc(t, result, dc_mats, norb, nelec, False, accumulate=True)
            if constant:
                axpby(constant, t, 1.0, result)
            return result

        return DeviceLinearOperator(dim(norb, nelec), matvec)
