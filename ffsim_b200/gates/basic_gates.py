"""Named two-orbital gates: signatures of python/ffsim/gates/basic_gates.py:54-629.

Every gate is one of three device operations, never a loop over amplitudes on the host:

* a rotation of two orbitals (Givens, fSWAP)  -> the fused orbital-rotation kernel with a 2x2 block;
* exp(i theta n_p)                             -> the number-operator-sum kernel;
* a phase controlled on occupied orbitals      -> ``ffb_apply_num_op_prod_phase`` (one sparse sweep).

The compound gates (tunneling, hop, fSim) chain these with ``copy=False`` exactly as the reference does.
"""

from __future__ import annotations

import cmath
import math
import numbers
from collections.abc import Sequence

import numpy as np
import torch

from ffsim_b200 import _device, _lib
from ffsim_b200.cistring import get_tables
from ffsim_b200.gates.num_op_sum import apply_num_op_sum_evolution
from ffsim_b200.gates.orbital_rotation import _check_dim, apply_orbital_rotation
from ffsim_b200.states.spin import Spin, pair_for_spin


def _two_orbital_matrix(norb: int, target_orbs, block) -> np.ndarray:
    mat = np.eye(norb, dtype=complex)
    mat[np.ix_(target_orbs, target_orbs)] = block
    return mat


def _rotate_two_orbitals(vec, mat, norb, nelec, spin, copy):
    if isinstance(nelec, numbers.Integral):
        return apply_orbital_rotation(vec, mat, norb=norb, nelec=nelec, copy=copy)
    return apply_orbital_rotation(vec, pair_for_spin(mat, spin=spin), norb=norb, nelec=nelec, copy=copy)


def apply_givens_rotation(vec, theta: float, target_orbs: tuple[int, int], norb: int, nelec,
                          spin: Spin = Spin.ALPHA_AND_BETA, *, phi: float = 0.0, copy: bool = True):
    r"""Givens rotation gate :math:`\prod_\sigma e^{i\varphi n_{p\sigma}} e^{\theta(a^\dagger_{p\sigma}a_{q\sigma}
    - a^\dagger_{q\sigma}a_{p\sigma})} e^{-i\varphi n_{p\sigma}}` (basic_gates.py:54-123)."""
    if len(set(target_orbs)) == 1:
        raise ValueError(f"The orbitals to rotate must be distinct. Got {target_orbs}.")
    c, s = math.cos(theta), cmath.exp(1j * phi) * math.sin(theta)
    mat = _two_orbital_matrix(norb, target_orbs, [[c, s], [-s.conjugate(), c]])
    return _rotate_two_orbitals(vec, mat, norb, nelec, spin, copy)


def apply_num_interaction(vec, theta: float, target_orb: int, norb: int, nelec,
                          spin: Spin = Spin.ALPHA_AND_BETA, *, copy: bool = True):
    r"""Number interaction :math:`\prod_\sigma e^{i\theta n_{p\sigma}}` (basic_gates.py:195-246)."""
    coeffs = np.zeros(norb)
    coeffs[target_orb] = 1.0
    if isinstance(nelec, numbers.Integral):
        return apply_num_op_sum_evolution(vec, coeffs, -theta, norb=norb, nelec=nelec, copy=copy)
    return apply_num_op_sum_evolution(vec, pair_for_spin(coeffs, spin), -theta, norb=norb, nelec=nelec, copy=copy)


def apply_tunneling_interaction(vec, theta: float, target_orbs: tuple[int, int], norb: int, nelec,
                                spin: Spin = Spin.ALPHA_AND_BETA, *, copy: bool = True):
    r"""Tunneling interaction :math:`\prod_\sigma e^{i\theta(a^\dagger_{p\sigma}a_{q\sigma} + \mathrm{h.c.})}`
    as number interaction, Givens rotation, inverse number interaction (basic_gates.py:126-192)."""
    if len(set(target_orbs)) == 1:
        raise ValueError(f"The orbitals to rotate must be distinct. Got {target_orbs}.")
    vec = apply_num_interaction(vec, -math.pi / 2, target_orbs[0], norb=norb, nelec=nelec, spin=spin, copy=copy)
    vec = apply_givens_rotation(vec, theta, target_orbs, norb=norb, nelec=nelec, spin=spin, copy=False)
    return apply_num_interaction(vec, math.pi / 2, target_orbs[0], norb=norb, nelec=nelec, spin=spin, copy=False)


def _mask(orbs: Sequence[int], norb: int) -> int:
    m = 0
    for p in orbs:
        if not 0 <= int(p) < norb:
            raise IndexError(f"orbital index {p} out of range for norb={norb}")
        m |= 1 << int(p)
    return m


def apply_num_op_prod_interaction(vec, theta: float, target_orbs: tuple[Sequence[int], Sequence[int]], norb: int,
                                  nelec: tuple[int, int], *, copy: bool = True):
    r"""exp(i theta prod n) over the listed alpha and beta orbitals: a phase on every amplitude whose strings
    have all of them occupied (basic_gates.py:375-424)."""
    nelec = (int(nelec[0]), int(nelec[1]))
    alpha_orbs, beta_orbs = target_orbs
    t, kind = _device.to_device(vec, copy=copy)
    _check_dim(t, norb, nelec)
    ta, tb = get_tables(norb, nelec[0]), get_tables(norb, nelec[1])
    data, row0, n_rows, col0, n_cols, ld = _device.local_block(t, ta.dim, tb.dim)
    with torch.cuda.device(data.device):
        _device.sync_device()
        _lib.check(_lib.lib.ffb_apply_num_op_prod_phase(
            ta.handle, tb.handle, _mask(alpha_orbs, norb), _mask(beta_orbs, norb), _lib.c128(cmath.exp(1j * theta)),
            data.data_ptr(), row0, n_rows, col0, n_cols, ld, _device.stream_ptr()))
    return _device.from_device(t, kind)


def apply_num_num_interaction(vec, theta: float, target_orbs: tuple[int, int], norb: int, nelec,
                              spin: Spin = Spin.ALPHA_AND_BETA, *, copy: bool = True):
    r"""Number-number interaction :math:`\prod_\sigma e^{i\theta n_{p\sigma} n_{q\sigma}}` (basic_gates.py:249-325)."""
    if len(set(target_orbs)) == 1:
        raise ValueError(f"The orbitals to interact must be distinct. Got {target_orbs}.")
    if isinstance(nelec, numbers.Integral):
        return apply_num_op_prod_interaction(vec, theta, (target_orbs, []), norb=norb, nelec=(int(nelec), 0), copy=copy)
    t, kind = _device.to_device(vec, copy=copy)
    if spin & Spin.ALPHA:
        t = apply_num_op_prod_interaction(t, theta, (target_orbs, []), norb=norb, nelec=nelec, copy=False)
    if spin & Spin.BETA:
        t = apply_num_op_prod_interaction(t, theta, ([], target_orbs), norb=norb, nelec=nelec, copy=False)
    return _device.from_device(t, kind)


def apply_on_site_interaction(vec, theta: float, target_orb: int, norb: int, nelec: tuple[int, int], *,
                              copy: bool = True):
    r"""On-site interaction :math:`e^{i\theta n_{p\alpha} n_{p\beta}}` (basic_gates.py:328-372)."""
    return apply_num_op_prod_interaction(vec, theta, ([target_orb], [target_orb]), norb=norb, nelec=nelec, copy=copy)


def apply_hop_gate(vec, theta: float, target_orbs: tuple[int, int], norb: int, nelec,
                   spin: Spin = Spin.ALPHA_AND_BETA, *, copy: bool = True):
    r"""Hop gate: Givens rotation followed by :math:`e^{i\pi n_p n_q}` (basic_gates.py:427-497)."""
    vec = apply_givens_rotation(vec, theta, target_orbs, norb=norb, nelec=nelec, spin=spin, copy=copy)
    return apply_num_num_interaction(vec, math.pi, target_orbs, norb=norb, nelec=nelec, spin=spin, copy=False)


def apply_fsim_gate(vec, theta: float, phi: float, target_orbs: tuple[int, int], norb: int, nelec,
                    spin: Spin = Spin.ALPHA_AND_BETA, *, copy: bool = True):
    r"""fSim gate: tunneling interaction by -theta, then number-number interaction by -phi (basic_gates.py:500-572)."""
    vec = apply_tunneling_interaction(vec, -theta, target_orbs, norb=norb, nelec=nelec, spin=spin, copy=copy)
    return apply_num_num_interaction(vec, -phi, target_orbs, norb=norb, nelec=nelec, spin=spin, copy=False)


def apply_fswap_gate(vec, target_orbs: tuple[int, int], norb: int, nelec,
                     spin: Spin = Spin.ALPHA_AND_BETA, *, copy: bool = True):
    r"""Fermionic swap: the orbital rotation by the 2x2 exchange matrix (basic_gates.py:575-629)."""
    mat = _two_orbital_matrix(norb, target_orbs, [[0, 1], [1, 0]])
    return _rotate_two_orbitals(vec, mat, norb, nelec, spin, copy)
