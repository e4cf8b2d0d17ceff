"""Number-number interaction ansatz operator (python/ffsim/variational/num_num.py:41-193).

exp(i sum theta_pq n_p n_q) on same-spin and opposite-spin orbital pairs.  On the state it is ONE
sweep of the diagonal Coulomb kernel with matrices assembled on the host.
"""

from __future__ import annotations

import itertools
import numbers
from dataclasses import dataclass

import numpy as np

from ffsim_b200.gates.diag_coulomb import apply_diag_coulomb_evolution
from ffsim_b200.variational._packing import PARAM_MISMATCH


def _check_pairs(pairs) -> None:
    if pairs is None:
        return
    if len(set(pairs)) != len(pairs):
        raise ValueError(f"Duplicate interaction pairs encountered: {pairs}.")
    for i, j in pairs:
        if i > j:
            raise ValueError("You must provide only upper triangular interaction pairs. "
                             f"Got {(i, j)}, which is a lower triangular pair.")


@dataclass(frozen=True)
class NumNumAnsatzOpSpinBalanced:
    """``interaction_pairs`` = (same-spin pairs, opposite-spin pairs), ``thetas`` the matching angle arrays."""

    norb: int
    interaction_pairs: tuple[list[tuple[int, int]], list[tuple[int, int]]]
    thetas: tuple[np.ndarray, np.ndarray]

    def __post_init__(self):
        for pairs, angles in zip(self.interaction_pairs, self.thetas):
            _check_pairs(pairs)
            if len(pairs) != len(angles):
                raise ValueError("The number of interaction pairs must be equal to the number of rotation angles."
                                 f"Got {len(pairs)}, and {len(angles)}.")

    def _apply_unitary_(self, vec, norb: int, nelec, copy: bool):
        if isinstance(nelec, numbers.Integral):
            return NotImplemented
        mat_aa, mat_ab = self.to_diag_coulomb_mats()
        return apply_diag_coulomb_evolution(vec, (mat_aa, mat_ab, mat_aa), time=-1.0, norb=norb, nelec=nelec, copy=copy)

    @staticmethod
    def n_params(interaction_pairs) -> int:
        return sum(len(pairs) for pairs in interaction_pairs)

    def to_parameters(self) -> np.ndarray:
        return np.concatenate([np.asarray(t, dtype=float) for t in self.thetas])

    @staticmethod
    def from_parameters(params: np.ndarray, norb: int, interaction_pairs) -> "NumNumAnsatzOpSpinBalanced":
        pairs_aa, pairs_ab = interaction_pairs
        expected = len(pairs_aa) + len(pairs_ab)
        if len(params) != expected:
            raise ValueError(PARAM_MISMATCH.format(expected, len(params)))
        return NumNumAnsatzOpSpinBalanced(norb=norb, interaction_pairs=interaction_pairs,
                                          thetas=(params[: len(pairs_aa)], params[len(pairs_aa) :]))

    @staticmethod
    def from_diag_coulomb_mats(diag_coulomb_mats) -> "NumNumAnsatzOpSpinBalanced":
        """The non-zero upper-triangular entries of (same-spin matrix, opposite-spin matrix)."""
        mat_aa, mat_ab = diag_coulomb_mats
        norb = mat_aa.shape[0]
        pairs, thetas = ([], []), ([], [])
        for k, mat in enumerate((mat_aa, mat_ab)):
            for pair in itertools.combinations_with_replacement(range(norb), 2):
                if mat[pair]:
                    pairs[k].append(pair)
                    thetas[k].append(mat[pair])
        return NumNumAnsatzOpSpinBalanced(norb=norb, interaction_pairs=pairs,
                                          thetas=(np.array(thetas[0]), np.array(thetas[1])))

    def to_diag_coulomb_mats(self) -> np.ndarray:
        """(2, norb, norb): symmetric same-spin and opposite-spin interaction matrices."""
        mats = np.zeros((2, self.norb, self.norb))
        for mat, pairs, angles in zip(mats, self.interaction_pairs, self.thetas):
            for (i, j), theta in zip(pairs, angles):
                mat[i, j] = mat[j, i] = theta
        return mats

    def _approx_eq_(self, other, rtol: float, atol: float) -> bool:
        if not isinstance(other, NumNumAnsatzOpSpinBalanced):
            return NotImplemented
        return (self.norb == other.norb and self.interaction_pairs == other.interaction_pairs
                and all(np.allclose(a, b, rtol=rtol, atol=atol) for a, b in zip(self.thetas, other.thetas)))
