"""Number-operator-sum evolution: signature of python/ffsim/gates/num_op_sum.py:62-72."""

from __future__ import annotations

import numbers

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables
from ffsim_b200.gates.diag_coulomb import _conjugate_orbital_rotation
from ffsim_b200.gates.orbital_rotation import _check_dim, _rotate_device, _split_mat


def _get_phases(coeffs, time: float):
    """python/ffsim/gates/num_op_sum.py:222-236."""
    if isinstance(coeffs, torch.Tensor):
        coeffs = coeffs.detach().cpu().numpy()
    if isinstance(coeffs, np.ndarray) and coeffs.ndim == 1:
        phases = np.ascontiguousarray(np.exp(-1j * time * coeffs.astype(float)))
        return phases, phases
    coeffs_a, coeffs_b = coeffs

    def ph(c):
        if c is None:
            return None
        return np.ascontiguousarray(np.exp(-1j * time * _device.as_host_matrix(c, dtype=float)))

    return ph(coeffs_a), ph(coeffs_b)


def _evolve_device(t, phases_a, phases_b, norb: int, nelec: tuple[int, int]) -> None:
    ta, tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
    data, row0, n_rows, col0, n_cols, ld = _device.local_block(t, ta.dim, tb.dim)
    with torch.cuda.device(data.device):
        _device.sync_device()
        _lib.check(
            _lib.lib.ffb_apply_num_op_sum_evolution_block(
                ta.handle, tb.handle, _lib.ptr(phases_a), _lib.ptr(phases_b), data.data_ptr(), row0, n_rows,
                col0, n_cols, ld, _device.stream_ptr(),
            )
        )


def apply_num_op_sum_evolution(
    vec,
    coeffs,
    time: float,
    norb: int,
    nelec: int | tuple[int, int],
    *,
    orbital_rotation=None,
    copy: bool = True,
):
    r"""Apply time evolution by a (rotated) linear combination of number operators.

    :math:`\mathcal{U} \exp(-i t \sum_{i\sigma} \lambda^{(\sigma)}_i n_{i\sigma})
    \mathcal{U}^\dagger`.  Arguments as in ``ffsim.apply_num_op_sum_evolution``.
    Both spin sectors are handled in one pass over the state.
    """
    if isinstance(nelec, numbers.Integral):
        nelec = (int(nelec), 0)
        coeffs = (_device.as_host_matrix(coeffs, dtype=float), None)
        if orbital_rotation is not None:
            orbital_rotation = (_device.as_host_matrix(orbital_rotation), None)
    else:
        nelec = (int(nelec[0]), int(nelec[1]))
    phases_a, phases_b = _get_phases(coeffs, time)
    for p in (phases_a, phases_b):
        if p is not None and p.shape != (norb,):
            raise ValueError(f"coeffs must have shape ({norb},), got {p.shape}")
    t, kind = _device.to_device(vec, copy=copy)
    _check_dim(t, norb, nelec)
    if orbital_rotation is not None:
        ca, cb = _conjugate_orbital_rotation(orbital_rotation)
        _rotate_device(t, ca, cb, norb, nelec)
    _evolve_device(t, phases_a, phases_b, norb, nelec)
    if orbital_rotation is not None:
        ra, rb = _split_mat(orbital_rotation)
        _rotate_device(t, ra, rb, norb, nelec)
    return _device.from_device(t, kind)
