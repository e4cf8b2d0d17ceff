"""Diagonal contractions (oracle; test infrastructure only)."""

from __future__ import annotations

import math

import numpy as np

from oracle.cistring import gen_occslst, make_strings
from oracle.gates import _conjugate_orbital_rotation, apply_orbital_rotation


def contract_num_op_sum_spin_into_buffer(vec, coeffs, occupations, out) -> None:
    """src/contract/num_op_sum.rs:20-39: out[row] += (sum coeffs[occ[row]]) * vec[row]."""
    occ = np.asarray(occupations, dtype=np.int64)
    row_coeff = np.asarray(coeffs, dtype=float)[occ].sum(axis=1) if occ.shape[1] else np.zeros(len(occ))
    out += row_coeff[:, None] * vec


def _pair_sums(mat, occ):
    out = np.zeros(len(occ))
    nocc = occ.shape[1]
    for j in range(nocc):
        for k in range(j, nocc):
            out += mat[occ[:, j], occ[:, k]]
    return out


def contract_diag_coulomb_into_buffer_num_rep(
    vec, mat_aa, mat_ab, mat_bb, norb, occupations_a, occupations_b, out
) -> None:
    """src/contract/diag_coulomb.rs:22-95."""
    occ_a = np.asarray(occupations_a, dtype=np.int64)
    occ_b = np.asarray(occupations_b, dtype=np.int64)
    beta = _pair_sums(mat_bb, occ_b)
    alpha = _pair_sums(mat_aa, occ_a)
    coeff_map = np.zeros((len(occ_a), norb))
    for j in range(occ_a.shape[1]):
        coeff_map += mat_ab[occ_a[:, j]]
    coeff = alpha[:, None] + beta[None, :]
    for j in range(occ_b.shape[1]):
        coeff = coeff + coeff_map[:, occ_b[:, j]]
    out += coeff * vec


def contract_diag_coulomb_into_buffer_z_rep(
    vec, mat_aa, mat_ab, mat_bb, norb, strings_a, strings_b, out
) -> None:
    """src/contract/diag_coulomb.rs:100-174."""
    sa = np.asarray(strings_a, dtype=np.int64)
    sb = np.asarray(strings_b, dtype=np.int64)

    def sign(strs, j):
        return 1.0 - 2.0 * ((strs >> np.int64(j)) & 1)

    def same_spin(strs, m):
        acc = np.zeros(len(strs))
        for j in range(norb):
            for k in range(j + 1, norb):
                acc += sign(strs, j) * sign(strs, k) * m[j, k]
        return acc

    beta = same_spin(sb, mat_bb)
    alpha = same_spin(sa, mat_aa)
    coeff_map = np.zeros((len(sa), norb))
    for j in range(norb):
        coeff_map += sign(sa, j)[:, None] * mat_ab[j][None, :]
    coeff = alpha[:, None] + beta[None, :]
    for j in range(norb):
        coeff = coeff + sign(sb, j)[None, :] * coeff_map[:, j][:, None]
    out += 0.25 * coeff * vec


def get_mats(mat, norb, z_representation):
    """python/ffsim/contract/diag_coulomb.py:102-130 (tuple semantics; 2-D ndarray case)."""
    def same_spin(m):
        if m is None:
            return np.zeros((norb, norb))
        m = np.array(m, dtype=float, copy=True)
        if not z_representation:
            m[np.diag_indices(norb)] *= 0.5
        return m

    if isinstance(mat, np.ndarray) and mat.ndim == 2:
        aa = same_spin(mat)
        return aa, np.asarray(mat, dtype=float), aa
    mat_aa, mat_ab, mat_bb = mat
    ab = np.zeros((norb, norb)) if mat_ab is None else np.asarray(mat_ab, dtype=float)
    return same_spin(mat_aa), ab, same_spin(mat_bb)


def contract_diag_coulomb(vec, mat, norb, nelec, *, z_representation=False):
    """python/ffsim/contract/diag_coulomb.py:41-192."""
    mat_aa, mat_ab, mat_bb = get_mats(mat, norb, z_representation)
    vec = vec.astype(complex, copy=False)
    n_alpha, n_beta = nelec
    dim_a, dim_b = math.comb(norb, n_alpha), math.comb(norb, n_beta)
    vec = vec.reshape((dim_a, dim_b))
    out = np.zeros_like(vec)
    if z_representation:
        contract_diag_coulomb_into_buffer_z_rep(
            vec, mat_aa, mat_ab, mat_bb, norb,
            make_strings(range(norb), n_alpha), make_strings(range(norb), n_beta), out,
        )
    else:
        contract_diag_coulomb_into_buffer_num_rep(
            vec, mat_aa, mat_ab, mat_bb, norb,
            gen_occslst(range(norb), n_alpha), gen_occslst(range(norb), n_beta), out,
        )
    return out.reshape(-1)


def contract_num_op_sum(vec, coeffs, norb, nelec):
    """python/ffsim/contract/num_op_sum.py:27-72."""
    vec = vec.astype(complex, copy=False)
    n_alpha, n_beta = nelec
    dim_a, dim_b = math.comb(norb, n_alpha), math.comb(norb, n_beta)
    vec = vec.reshape((dim_a, dim_b))
    out = np.zeros_like(vec)
    contract_num_op_sum_spin_into_buffer(vec, coeffs, gen_occslst(range(norb), n_alpha), out)
    contract_num_op_sum_spin_into_buffer(vec.T, coeffs, gen_occslst(range(norb), n_beta), out.T)
    return out.reshape(-1)


def diag_coulomb_matvec(vec, mat, norb, nelec, *, orbital_rotation=None, z_representation=False):
    """matvec of python/ffsim/contract/diag_coulomb.py:195-272."""
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, _conjugate_orbital_rotation(orbital_rotation), norb, nelec)
    vec = contract_diag_coulomb(vec, mat, norb, nelec, z_representation=z_representation)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, orbital_rotation, norb, nelec, copy=False)
    return vec


def num_op_sum_matvec(vec, coeffs, norb, nelec, *, orbital_rotation=None):
    """matvec of python/ffsim/contract/num_op_sum.py:75-131."""
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, orbital_rotation.T.conj(), norb, nelec)
    vec = contract_num_op_sum(vec, coeffs, norb, nelec)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, orbital_rotation, norb, nelec, copy=False)
    return vec
