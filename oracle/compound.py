"""Independent closed-form oracle for orbital rotations (test infrastructure only).

Not Givens based.  An orbital rotation U acts on the k-electron string space by
the k-th compound matrix, C_k(U)[I, J] = det U[I, J] over occupied sets in string
order, so ``new = C_ka(Ua) @ M @ C_kb(Ub).T``.  For a rotated Slater determinant
this collapses to an outer product of minors, the formula of
python/ffsim/states/slater.py:303-354 (tested tests/python/states/slater_test.py:242-288).
"""

from __future__ import annotations

import numpy as np

from oracle.cistring import gen_occslst


def compound_matrix(mat, norb, nocc):
    occ = gen_occslst(range(norb), nocc).astype(np.int64)
    if nocc == 0:
        return np.ones((1, 1), dtype=complex)
    sub = mat[occ[:, None, :, None], occ[None, :, None, :]]  # [I, J, k, k]
    return np.linalg.det(sub)


def apply_orbital_rotation_compound(vec, mat, norb, nelec):
    mat_a, mat_b = (mat, mat) if isinstance(mat, np.ndarray) and mat.ndim == 2 else mat
    n_alpha, n_beta = nelec
    eye = np.eye(norb, dtype=complex)
    ca = compound_matrix(eye if mat_a is None else mat_a, norb, n_alpha)
    cb = compound_matrix(eye if mat_b is None else mat_b, norb, n_beta)
    m = vec.reshape(ca.shape[0], cb.shape[0])
    return (ca @ m @ cb.T).reshape(-1)


def slater_minors(mat, norb, nocc, occupied=None):
    """Amplitudes of U|occupied> on one spin sector: det U[I, occupied] for each string I."""
    occupied = list(range(nocc)) if occupied is None else list(occupied)
    occ = gen_occslst(range(norb), nocc).astype(np.int64)
    if nocc == 0:
        return np.ones(1, dtype=complex)
    sub = mat[occ[:, :, None], np.array(occupied)[None, None, :]]
    return np.linalg.det(sub)
