"""Expectation values of products of one-body operators in a Slater determinant (Wick's theorem).

Counterpart of python/ffsim/states/wick.py:23-132; host-side (spin-orbital matrices of size 2 norb), used
by the state-dependent qDRIFT probabilities.  The sum over complete contractions is organised by the
cycles of the permutation that pairs creators with annihilators: every cycle is the trace of a chain of
small matrices, so no general tensor-contraction engine is needed.
"""

from __future__ import annotations

import itertools
from collections.abc import Sequence

import numpy as np


def _contraction_sign(perm: Sequence[int]) -> int:
    """Sign of the complete contraction that pairs creator i (slot 2i of c0 a0 c1 a1 ...) with annihilator
    perm[i] (slot 2 perm[i] + 1): minus one for every operator standing between the two partners when
    the pair is taken out, pairs taken in creator order (= parity of the crossings of the pairing)."""
    slots = list(range(2 * len(perm)))
    sign = 1
    for i, j in enumerate(perm):
        pc, pa = slots.index(2 * i), slots.index(2 * j + 1)
        if (abs(pc - pa) - 1) % 2:
            sign = -sign
        slots.remove(2 * i)
        slots.remove(2 * j + 1)
    return sign


def expectation_one_body_product(one_rdm: np.ndarray, one_body_tensors: Sequence[np.ndarray]) -> complex:
    r""":math:`\langle\psi| O_1 O_2 \cdots O_k |\psi\rangle` for :math:`O_i = \sum_{pq} M^{(i)}_{pq} a^\dagger_p a_q`
    and a Slater determinant with one-body reduced density matrix ``one_rdm`` (spin-orbital basis, not
    spin-summed; the matrices must have the same shape)."""
    k = len(one_body_tensors)
    if k == 0:
        return 1.0
    one_rdm = np.asarray(one_rdm)
    hole = np.eye(one_rdm.shape[0]) - one_rdm
    mats_t = [np.asarray(m).T for m in one_body_tensors]
    total = 0.0
    for perm in itertools.permutations(range(k)):
        term = float(_contraction_sign(perm))
        seen = [False] * k
        for start in range(k):
            if seen[start]:
                continue
            # cycle start -> perm[start] -> ...: trace of  M_i^T C_{i, perm[i]}  along the cycle, where
            # C = <a^+ a> (the 1-RDM) when the creator stands left of its partner, <a a^+> otherwise
            chain = None
            i = start
            while not seen[i]:
                seen[i] = True
                j = perm[i]
                link = mats_t[i] @ (one_rdm if i <= j else hole)
                chain = link if chain is None else chain @ link
                i = j
            term = term * np.trace(chain)
        total = total + term
    return total


def expectation_one_body_power(one_rdm: np.ndarray, one_body_tensor: np.ndarray, power: int = 1) -> complex:
    r""":math:`\langle\psi| O^k |\psi\rangle` (wick.py:98-132)."""
    return expectation_one_body_product(one_rdm, [one_body_tensor] * power)
