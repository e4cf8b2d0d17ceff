"""Givens-rotation ansatz operator (python/ffsim/variational/givens.py:27-296).

A list of Givens rotations G(theta, phi) on orbital pairs followed by one layer of single-orbital
phases.  On the state it IS an orbital rotation: ``to_orbital_rotation`` multiplies the norb x norb
matrices on the host and ``_apply_unitary_`` hands the product to the fused rotation kernel (one
plan, a few sweeps), instead of one gate per pair.
"""

from __future__ import annotations

import cmath
import itertools
import math
from dataclasses import dataclass

import numpy as np

from ffsim_b200 import linalg
from ffsim_b200.gates.orbital_rotation import apply_orbital_rotation
from ffsim_b200.variational._packing import PARAM_MISMATCH


def _brickwork_slots(norb: int):
    """Empty brickwork: alternating layers of pairs (0,1),(2,3).. and (1,2),(3,4).., ceil(n/2) + floor(n/2) of them."""
    q, r = divmod(norb, 2)
    even = [[((i, i + 1), 0.0, 0.0) for i in range(0, norb - 1, 2)] for _ in range(q + r)]
    odd = [[((i, i + 1), 0.0, 0.0) for i in range(1, norb - 1, 2)] for _ in range(q)]
    return even, odd


def _into_brickwork(pairs, thetas, phis, norb: int):
    """Place a sparse list of adjacent-pair rotations into the full brickwork pattern, every rotation in
    the earliest layer after the last layer that touched one of its orbitals (givens.py:226-296)."""
    even, odd = _brickwork_slots(norb)
    last_even, last_odd = [-1] * norb, [-1] * norb
    for (i, j), theta, phi in zip(pairs, thetas, phis):
        if i > j:
            i, j, theta, phi = j, i, -theta, -phi
        if i % 2 == 0:
            layer = max(last_even[i], last_even[j], last_odd[i], last_odd[j]) + 1
            even[layer][i // 2] = ((i, j), theta, phi)
            last_even[i] = last_even[j] = layer
        else:
            layer = max(last_odd[i] + 1, last_odd[j] + 1, last_even[i], last_even[j])
            odd[layer][i // 2] = ((i, j), theta, phi)
            last_odd[i] = last_odd[j] = layer
    flat = [g for e, o in itertools.zip_longest(even, odd, fillvalue=()) for g in (*e, *o)]
    return [g[0] for g in flat], [g[1] for g in flat], [g[2] for g in flat]


@dataclass(frozen=True)
class GivensAnsatzOp:
    """Givens rotations on ``interaction_pairs`` (angles ``thetas``, optional phases ``phis``) followed by
    optional single-orbital phases ``phase_angles``."""

    norb: int
    interaction_pairs: list[tuple[int, int]]
    thetas: np.ndarray
    phis: np.ndarray | None
    phase_angles: np.ndarray | None

    def __post_init__(self):
        n = len(self.interaction_pairs)
        if len(self.thetas) != n:
            raise ValueError("The number of thetas must equal the number of interaction pairs. "
                             f"Got {len(self.thetas)} and {n}.")
        if self.phis is not None and len(self.phis) != n:
            raise ValueError("The number of phis must equal the number of interaction pairs. "
                             f"Got {len(self.phis)} and {n}.")
        if self.phase_angles is not None and len(self.phase_angles) != self.norb:
            raise ValueError("The number of phase angles must equal the number of orbitals. "
                             f"Got {len(self.phase_angles)} and {self.norb}.")

    def _apply_unitary_(self, vec, norb: int, nelec, copy: bool):
        return apply_orbital_rotation(vec, self.to_orbital_rotation(), norb=norb, nelec=nelec, copy=copy)

    @staticmethod
    def n_params(norb: int, interaction_pairs, with_phis: bool = True, with_phase_angles: bool = True) -> int:
        return (1 + with_phis) * len(interaction_pairs) + with_phase_angles * norb

    def to_parameters(self) -> np.ndarray:
        parts = [np.asarray(self.thetas, dtype=float)]
        if self.phis is not None:
            parts.append(np.asarray(self.phis, dtype=float))
        if self.phase_angles is not None:
            parts.append(np.asarray(self.phase_angles, dtype=float))
        return np.concatenate(parts)

    @staticmethod
    def from_parameters(params: np.ndarray, norb: int, interaction_pairs, with_phis: bool = True,
                        with_phase_angles: bool = True) -> "GivensAnsatzOp":
        n = len(interaction_pairs)
        expected = GivensAnsatzOp.n_params(norb, interaction_pairs, with_phis, with_phase_angles)
        if len(params) != expected:
            raise ValueError(PARAM_MISMATCH.format(expected, len(params)))
        thetas, rest = params[:n], params[n:]
        phis = rest[:n] if with_phis else None
        phase_angles = (rest[n:] if with_phis else rest) if with_phase_angles else None
        return GivensAnsatzOp(norb=norb, interaction_pairs=interaction_pairs, thetas=thetas, phis=phis,
                              phase_angles=phase_angles)

    @staticmethod
    def from_orbital_rotation(orbital_rotation: np.ndarray) -> "GivensAnsatzOp":
        """Givens decomposition of a unitary, expanded to the full brickwork pattern."""
        norb = orbital_rotation.shape[0]
        rotations, phases = linalg.givens_decomposition(orbital_rotation)
        pairs, thetas, phis = [], [], []
        for c, s, i, j in rotations:
            r, phi = cmath.polar(s)
            pairs.append((i, j))
            thetas.append(math.atan2(r, c))
            phis.append(phi)
        pairs, thetas, phis = _into_brickwork(pairs, thetas, phis, norb)
        return GivensAnsatzOp(norb=norb, interaction_pairs=pairs, thetas=np.array(thetas), phis=np.array(phis),
                              phase_angles=np.angle(phases))

    def to_orbital_rotation(self) -> np.ndarray:
        """The norb x norb unitary the gate sequence implements: D * G_1 * ... * G_L with the last gate
        applied to the columns first, each a plane rotation of columns (j, i) by (cos t, sin t e^{-i phi})."""
        n = len(self.interaction_pairs)
        phis = np.zeros(n) if self.phis is None else self.phis
        phase_angles = np.zeros(self.norb) if self.phase_angles is None else self.phase_angles
        u = np.diag(np.exp(1j * np.asarray(phase_angles, dtype=float))).astype(complex)
        for (i, j), theta, phi in zip(self.interaction_pairs[::-1], self.thetas[::-1], phis[::-1]):
            c, s = math.cos(theta), cmath.rect(math.sin(theta), -phi)
            col_j, col_i = u[:, j].copy(), u[:, i].copy()
            u[:, j] = c * col_j + s * col_i              # LAPACK zrot(x, y, c, s): x <- c x + s y,
            u[:, i] = c * col_i - np.conj(s) * col_j     #                          y <- c y - conj(s) x
        return u

    def _approx_eq_(self, other, rtol: float, atol: float) -> bool:
        if not isinstance(other, GivensAnsatzOp):
            return NotImplemented
        if self.norb != other.norb or self.interaction_pairs != other.interaction_pairs:
            return False
        for a, b in ((self.thetas, other.thetas), (self.phis, other.phis), (self.phase_angles, other.phase_angles)):
            if (a is None) != (b is None):
                return False
            if a is not None and not np.allclose(a, b, rtol=rtol, atol=atol):
                return False
        return True
