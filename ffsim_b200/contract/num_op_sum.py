"""Number-operator-sum contraction: python/ffsim/contract/num_op_sum.py:27-29,75-81."""

from __future__ import annotations

import math

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables
from ffsim_b200.contract.linop import DeviceLinearOperator
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device


def _contract_device(t, out, coeffs: np.ndarray, norb, nelec, accumulate) -> None:
    ta, tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
    _device.same_layout(t, out)
    data, row0, n_rows, col0, n_cols, ld = _device.local_block(t, ta.dim, tb.dim)
    out_data = _device.local_block(out, ta.dim, tb.dim)[0]
    with torch.cuda.device(data.device):
        _device.sync_device()
        _lib.check(
            _lib.lib.ffb_contract_num_op_sum_block(
                ta.handle, tb.handle, _lib.ptr(coeffs), _lib.ptr(coeffs), data.data_ptr(), out_data.data_ptr(),
                int(bool(accumulate)), row0, n_rows, col0, n_cols, ld, _device.stream_ptr(),
            )
        )


def _coeffs(coeffs, norb: int) -> np.ndarray:
    c = np.ascontiguousarray(_device.as_host_matrix(coeffs), dtype=float)
    if c.shape != (norb,):
        raise ValueError(f"coeffs must have shape ({norb},), got {c.shape}")
    return c


def contract_num_op_sum(vec, coeffs, norb: int, nelec: tuple[int, int]):
    r"""Contract :math:`\sum_{i\sigma} \lambda_i n_{i\sigma}` with a vector (new vector returned).

    Both spin sectors are handled in one pass over the state; the reference makes
    two (contract/num_op_sum.py:55-70).
    """
    nelec = (int(nelec[0]), int(nelec[1]))
    c = _coeffs(coeffs, norb)
    t, kind = _device.to_device(vec, copy=False)
    _check_dim(t, norb, nelec)
    out = _device.empty_like(t)
    _contract_device(t, out, c, norb, nelec, accumulate=False)
    return _device.from_device(out, kind)


def num_op_sum_linop(coeffs, norb: int, nelec: tuple[int, int], *, orbital_rotation=None) -> DeviceLinearOperator:
    """Linear operator of a (rotated) number-operator sum (contract/num_op_sum.py:75-131)."""
    nelec = (int(nelec[0]), int(nelec[1]))
    dim = math.comb(norb, nelec[0]) * math.comb(norb, nelec[1])
    c = _coeffs(coeffs, norb)
    rot = None if orbital_rotation is None else _device.as_host_matrix(orbital_rotation)

    def matvec(t):
        out = _device.empty_like(t)
        if rot is None:
            _contract_device(t, out, c, norb, nelec, accumulate=False)
            return out
        work = t.clone()
        rot_dag = rot.T.conj()
        _rotate_device(work, rot_dag, rot_dag, norb, nelec)
        _contract_device(work, out, c, norb, nelec, accumulate=False)
        _rotate_device(out, rot, rot, norb, nelec)
        return out

    return DeviceLinearOperator(dim, matvec)
