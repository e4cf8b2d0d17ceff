"""Occupation-string tables (oracle; test infrastructure only).

Restates the published algorithm of ``pyscf.fci.cistring`` (pyscf 2.14.0, not
vendored in /root/reference) at the call sites the hot path uses:

* ``make_strings`` / ``gen_occslst``  <- python/ffsim/_cistring.py:21-31
* permuted ``make_strings`` + argsort <- python/ffsim/gates/orbital_rotation.py:203-236

Bit ``i`` of a string is orbital ``i``; for ``orb_list == range(norb)`` strings
come out in ascending integer order (tests/python/states/bitstring_test.py:24-97).
"""

from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


def make_strings(orb_list, nelec: int) -> np.ndarray:
    """All strings with ``nelec`` of the orbitals in ``orb_list`` occupied.

    pyscf order: recurse on the LAST entry of ``orb_list`` -- first every
    string that leaves it empty, then every string that fills it.
    """
    orbs = [int(o) for o in orb_list]
    if nelec < 0:
        raise ValueError("nelec must be non-negative")
    if nelec == 0:
        return np.zeros(1, dtype=np.int64)
    if nelec > len(orbs):
        return np.zeros(0, dtype=np.int64)

    # table[m][e]: strings over the first m orbitals with e electrons, built
    # bottom-up instead of by recursion (same order as the recursive definition).
    prev = [np.zeros(1, dtype=np.int64)] + [None] * nelec  # m = 0
    for m in range(1, len(orbs) + 1):
        bit = np.int64(1) << np.int64(orbs[m - 1])
        cur = [np.zeros(1, dtype=np.int64)] + [None] * nelec
        for e in range(1, min(m, nelec) + 1):
            without = prev[e] if (e <= m - 1 and prev[e] is not None) else np.zeros(0, np.int64)
            with_ = prev[e - 1] | bit
            cur[e] = np.concatenate([without, with_])
        prev = cur
    return prev[nelec]


def gen_occslst(orb_list, nelec: int) -> np.ndarray:
    """Occupied-orbital lists, one row per string, ascending within a row.

    python/ffsim/_cistring.py:27-31 casts the pyscf int32 result to ``np.uint``.
    """
    orbs = [int(o) for o in orb_list]
    strings = make_strings(orbs, nelec)
    out = np.zeros((len(strings), nelec), dtype=np.uint64)
    if nelec == 0 or len(strings) == 0:
        return out
    # pyscf lists, for each string, the entries of orb_list that are occupied in
    # list order; for range(norb) that is ascending orbital index.
    col = np.zeros(len(strings), dtype=np.int64)
    for o in orbs:
        occ = ((strings >> np.int64(o)) & 1).astype(bool)
        out[occ, col[occ]] = o
        col += occ
    return out


def strs2addr(norb: int, nelec: int, strings) -> np.ndarray:
    """Colexicographic rank of each string = its address (pyscf ``strs2addr``)."""
    strings = np.atleast_1d(np.asarray(strings, dtype=np.int64))
    addr = np.zeros(strings.shape, dtype=np.int64)
    seen = np.zeros(strings.shape, dtype=np.int64)
    for pos in range(norb):
        occ = (strings >> np.int64(pos)) & 1
        seen += occ
        binom = np.array([math.comb(pos, int(m)) for m in range(nelec + 2)], dtype=np.int64)
        addr += occ * binom[np.minimum(seen, nelec + 1)]
    return addr


def shifted_orbitals(norb: int, target_orbs: tuple[int, ...]) -> np.ndarray:
    """python/ffsim/gates/orbital_rotation.py:230-236."""
    n_rest = norb - len(target_orbs)
    orbitals = list(range(n_rest))
    for index, val in sorted(zip(target_orbs, range(n_rest, norb))):
        orbitals.insert(index, val)
    return np.array(orbitals, dtype=np.int64)


@lru_cache(maxsize=None)
def zero_one_subspace_indices(norb: int, nocc: int, target_orbs: tuple[int, int]) -> np.ndarray:
    """python/ffsim/gates/orbital_rotation.py:203-213 (argsort construction)."""
    strings = make_strings(shifted_orbitals(norb, target_orbs), nocc)
    indices = np.argsort(strings, kind="stable")
    n00 = math.comb(norb - 2, nocc)
    n11 = math.comb(norb - 2, nocc - 2) if nocc >= 2 else 0
    return indices[n00 : len(indices) - n11].astype(np.uint64)


@lru_cache(maxsize=None)
def one_subspace_indices(norb: int, nocc: int, target_orbs: tuple[int, ...]) -> np.ndarray:
    """python/ffsim/gates/orbital_rotation.py:216-226."""
    strings = make_strings(shifted_orbitals(norb, target_orbs), nocc)
    indices = np.argsort(strings, kind="stable")
    n0 = math.comb(norb, nocc)
    if nocc >= len(target_orbs):
        n0 -= math.comb(norb - len(target_orbs), nocc - len(target_orbs))
    return indices[n0:].astype(np.uint64)
