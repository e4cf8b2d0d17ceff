"""Diagonal Coulomb contraction: python/ffsim/contract/diag_coulomb.py:41-48,195-204."""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables
from ffsim_b200.contract.linop import DeviceLinearOperator
from ffsim_b200.gates.diag_coulomb import _conjugate_orbital_rotation
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device, _split_mat


def _get_mats(mat: Any, norb: int, z_representation: bool):
    """python/ffsim/contract/diag_coulomb.py:102-130 (``None`` members stay ``None`` = zeros)."""
    if isinstance(mat, torch.Tensor):
        mat = mat.detach().cpu().numpy()

    def same_spin(m):
        if m is None:
            return None
        m = np.array(_device.as_host_matrix(m), dtype=float, copy=True)
        if not z_representation:
            m[np.diag_indices(norb)] *= 0.5
        return np.ascontiguousarray(m)

    def cross(m):
        return None if m is None else np.ascontiguousarray(_device.as_host_matrix(m), dtype=float)

    if isinstance(mat, np.ndarray) and mat.ndim == 2:
        aa = same_spin(mat)
        return aa, cross(mat), aa
    mat_aa, mat_ab, mat_bb = mat
    return same_spin(mat_aa), cross(mat_ab), same_spin(mat_bb)


def _contract_device(t, out, mats, norb, nelec, z_representation, accumulate) -> None:
    aa, ab, bb = mats
    ta, tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
    _device.same_layout(t, out)
    data, row0, n_rows, col0, n_cols, ld = _device.local_block(t, ta.dim, tb.dim)
    out_data = _device.local_block(out, ta.dim, tb.dim)[0]
    with torch.cuda.device(data.device):
        _device.sync_device()
        _lib.check(
            _lib.lib.ffb_contract_diag_coulomb_block(
                ta.handle, tb.handle, _lib.ptr(aa), _lib.ptr(ab), _lib.ptr(bb),
                int(bool(z_representation)), data.data_ptr(), out_data.data_ptr(), int(bool(accumulate)),
                row0, n_rows, col0, n_cols, ld, _device.stream_ptr(),
            )
        )


def contract_diag_coulomb(
    vec, mat, norb: int, nelec: tuple[int, int], *, z_representation: bool = False
):
    r"""Contract a diagonal Coulomb operator with a vector.

    Returns :math:`\sum_{ij,\sigma\tau} Z^{(\sigma\tau)}_{ij} n_{i\sigma} n_{j\tau}/2\,|v\rangle`
    as a new vector (``vec`` is left untouched), as ``ffsim.contract.contract_diag_coulomb``.
    """
    nelec = (int(nelec[0]), int(nelec[1]))
    mats = _get_mats(mat, norb, z_representation)
    t, kind = _device.to_device(vec, copy=False)
    _check_dim(t, norb, nelec)
    out = _device.empty_like(t)
    _contract_device(t, out, mats, norb, nelec, z_representation, accumulate=False)
    return _device.from_device(out, kind)


def diag_coulomb_linop(
    mat, norb: int, nelec: tuple[int, int], *, orbital_rotation=None, z_representation: bool = False
) -> DeviceLinearOperator:
    """Linear operator of a (rotated) diagonal Coulomb operator (contract/diag_coulomb.py:195-272)."""
    nelec = (int(nelec[0]), int(nelec[1]))
    dim = math.comb(norb, nelec[0]) * math.comb(norb, nelec[1])
    mats = _get_mats(mat, norb, z_representation)

    def matvec(t):
        out = _device.empty_like(t)
        if orbital_rotation is None:
            _contract_device(t, out, mats, norb, nelec, z_representation, accumulate=False)
            return out
        work = t.clone()
        ca, cb = _conjugate_orbital_rotation(orbital_rotation)
        _rotate_device(work, ca, cb, norb, nelec)
        _contract_device(work, out, mats, norb, nelec, z_representation, accumulate=False)
        ra, rb = _split_mat(orbital_rotation)
        _rotate_device(out, ra, rb, norb, nelec)
        return out

    return DeviceLinearOperator(dim, matvec)
