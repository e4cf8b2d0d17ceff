"""Diagonal Coulomb evolution: signature of python/ffsim/gates/diag_coulomb.py:68-80."""

from __future__ import annotations

import numbers
from typing import Any

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device, _split_mat


def _conjugate_orbital_rotation(orbital_rotation):
    """python/ffsim/gates/diag_coulomb.py:29-39."""
    mat_a, mat_b = _split_mat(orbital_rotation)
    if mat_a is mat_b:
        c = None if mat_a is None else mat_a.T.conj()
        return c, c
    return (None if mat_a is None else mat_a.T.conj(), None if mat_b is None else mat_b.T.conj())


def _get_mat_exp(mat: Any, time: float, norb: int, z_representation: bool):
    """python/ffsim/gates/diag_coulomb.py:223-275; ``None`` members stay ``None`` (= all ones)."""
    if isinstance(mat, torch.Tensor):
        mat = mat.detach().cpu().numpy()

    def same_spin(m):
        if m is None:
            return None
        m = np.array(_device.as_host_matrix(m), dtype=float, copy=True)
        m[np.diag_indices(norb)] *= 0.5
        if z_representation:
            m *= 0.25
        return np.ascontiguousarray(np.exp(-1j * time * m))

    def cross(m):
        if m is None:
            return None
        m = np.array(_device.as_host_matrix(m), dtype=float, copy=True)
        if z_representation:
            m *= 0.25
        return np.ascontiguousarray(np.exp(-1j * time * m))

    if isinstance(mat, np.ndarray) and mat.ndim == 2:
        aa = same_spin(mat)
        return aa, cross(mat), aa
    mat_aa, mat_ab, mat_bb = mat
    return same_spin(mat_aa), cross(mat_ab), same_spin(mat_bb)


def _evolve_device(t, mats, norb: int, nelec: tuple[int, int], z_representation: bool) -> None:
    aa, ab, bb = mats
    ta, tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
    data, row0, n_rows, col0, n_cols, ld = _device.local_block(t, ta.dim, tb.dim)
    with torch.cuda.device(data.device):
        _device.sync_device()
        _lib.check(
            _lib.lib.ffb_apply_diag_coulomb_evolution_block(
                ta.handle, tb.handle, _lib.ptr(aa), _lib.ptr(ab), _lib.ptr(bb),
                int(bool(z_representation)), data.data_ptr(), row0, n_rows, col0, n_cols, ld,
                _device.stream_ptr(),
            )
        )


def apply_diag_coulomb_evolution(
    vec,
    mat,
    time: float,
    norb: int,
    nelec: int | tuple[int, int],
    *,
    orbital_rotation=None,
    z_representation: bool = False,
    copy: bool = True,
):
    r"""Apply time evolution by a (rotated) diagonal Coulomb operator.

    :math:`\mathcal{U} \exp(-i t \sum_{ij,\sigma\tau} Z^{(\sigma\tau)}_{ij}
    n_{i\sigma} n_{j\tau} / 2)\, \mathcal{U}^\dagger`.  Arguments as in
    ``ffsim.apply_diag_coulomb_evolution``; ``vec`` may be a CUDA tensor.
    """
    if isinstance(nelec, numbers.Integral):
        if z_representation:
            raise NotImplementedError  # diag_coulomb.py:133-135
        nelec = (int(nelec), 0)
        if orbital_rotation is not None:
            orbital_rotation = (_device.as_host_matrix(orbital_rotation), None)
    else:
        nelec = (int(nelec[0]), int(nelec[1]))
    mats = _get_mat_exp(mat, time, norb, z_representation)
    t, kind = _device.to_device(vec, copy=copy)
    _check_dim(t, norb, nelec)
    if orbital_rotation is not None:
        ca, cb = _conjugate_orbital_rotation(orbital_rotation)
        _rotate_device(t, ca, cb, norb, nelec)
    _evolve_device(t, mats, norb, nelec, z_representation)
    if orbital_rotation is not None:
        ra, rb = _split_mat(orbital_rotation)
        _rotate_device(t, ra, rb, norb, nelec)
    return _device.from_device(t, kind)
