"""Gate kernels and drivers (oracle; test infrastructure only).

Kernels restate the Rust loops; drivers restate the Python that sequences them.
Vectorised over whole rows with numpy, same per-element arithmetic.
"""

from __future__ import annotations

import math

import numpy as np

from oracle.cistring import gen_occslst, make_strings, one_subspace_indices, zero_one_subspace_indices
from oracle.givens import givens_decomposition

# --------------------------------------------------------------------------- kernels


def apply_givens_rotation_in_place(vec, c: float, s: complex, slice1, slice2) -> None:
    """src/gates/orbital_rotation.rs:20-104: rows slice1[k], slice2[k] <- zrot."""
    if len(slice1) == 0:
        return
    i = np.asarray(slice1, dtype=np.int64)
    j = np.asarray(slice2, dtype=np.int64)
    x = vec[i]
    y = vec[j]
    vec[i] = c * x + s * y
    vec[j] = c * y - np.conj(s) * x


def apply_phase_shift_in_place(vec, phase: complex, indices) -> None:
    """src/gates/phase_shift.rs:18-30."""
    vec[np.asarray(indices, dtype=np.int64)] *= phase


def apply_num_op_sum_evolution_in_place(vec, phases, occupations) -> None:
    """src/gates/num_op_sum.rs:20-36: row *= prod(phases[orb] for orb in occ[row])."""
    occ = np.asarray(occupations, dtype=np.int64)
    row_phase = np.prod(np.asarray(phases)[occ], axis=1) if occ.shape[1] else np.ones(len(occ), complex)
    vec *= row_phase[:, None]


def _pair_products(mat_exp, occ):
    """prod_{j<=k} M[o_j, o_k] per row of occ (src/gates/diag_coulomb.rs:47-59)."""
    out = np.ones(len(occ), dtype=complex)
    nocc = occ.shape[1]
    for j in range(nocc):
        for k in range(j, nocc):
            out *= mat_exp[occ[:, j], occ[:, k]]
    return out


def apply_diag_coulomb_evolution_in_place_num_rep(
    vec, mat_exp_aa, mat_exp_ab, mat_exp_bb, norb, occupations_a, occupations_b
) -> None:
    """src/gates/diag_coulomb.rs:21-90."""
    occ_a = np.asarray(occupations_a, dtype=np.int64)
    occ_b = np.asarray(occupations_b, dtype=np.int64)
    beta_phases = _pair_products(mat_exp_bb, occ_b)
    alpha_phases = _pair_products(mat_exp_aa, occ_a)
    phase_map = np.ones((len(occ_a), norb), dtype=complex)
    for j in range(occ_a.shape[1]):
        phase_map *= mat_exp_ab[occ_a[:, j]]
    phase = alpha_phases[:, None] * beta_phases[None, :]
    for j in range(occ_b.shape[1]):
        phase = phase * phase_map[:, occ_b[:, j]]
    vec *= phase


def apply_diag_coulomb_evolution_in_place_z_rep(
    vec, mat_exp_aa, mat_exp_ab, mat_exp_bb, mat_exp_aa_conj, mat_exp_ab_conj, mat_exp_bb_conj,
    norb, strings_a, strings_b,
) -> None:
    """src/gates/diag_coulomb.rs:95-190."""
    sa = np.asarray(strings_a, dtype=np.int64)
    sb = np.asarray(strings_b, dtype=np.int64)

    def bits(strs, j):
        return ((strs >> np.int64(j)) & 1).astype(bool)

    def same_spin(strs, m, m_conj):
        out = np.ones(len(strs), dtype=complex)
        for j in range(norb):
            for k in range(j + 1, norb):
                differ = bits(strs, j) ^ bits(strs, k)
                out *= np.where(differ, m_conj[j, k], m[j, k])
        return out

    beta_phases = same_spin(sb, mat_exp_bb, mat_exp_bb_conj)
    alpha_phases = same_spin(sa, mat_exp_aa, mat_exp_aa_conj)
    phase_map = np.ones((len(sa), norb), dtype=complex)
    for j in range(norb):
        phase_map *= np.where(bits(sa, j)[:, None], mat_exp_ab_conj[j][None, :], mat_exp_ab[j][None, :])
    phase = alpha_phases[:, None] * beta_phases[None, :]
    for j in range(norb):
        col = phase_map[:, j][:, None]
        phase = phase * np.where(bits(sb, j)[None, :], np.conj(col), col)
    vec *= phase


# --------------------------------------------------------------------------- drivers


def _givens_decompositions(mat):
    """python/ffsim/gates/orbital_rotation.py:157-174."""
    if isinstance(mat, np.ndarray) and mat.ndim == 2:
        d = givens_decomposition(mat)
        return d, d
    mat_a, mat_b = mat
    return (
        None if mat_a is None else givens_decomposition(mat_a),
        None if mat_b is None else givens_decomposition(mat_b),
    )


def _rotate_one_spin(vec, decomp, norb, nocc) -> None:
    """Loops of python/ffsim/gates/orbital_rotation.py:130-138 (one spin sector)."""
    rotations, phase_shifts = decomp
    for c, s, i, j in rotations:
        assert abs(i - j) == 1
        idx = zero_one_subspace_indices(norb, nocc, (i, j))
        half = len(idx) // 2
        apply_givens_rotation_in_place(vec, c, np.conj(s), idx[:half], idx[half:])
    for i, phase in enumerate(phase_shifts):
        apply_phase_shift_in_place(vec, phase, one_subspace_indices(norb, nocc, (i,)))


def apply_orbital_rotation(vec, mat, norb, nelec, *, copy=True):
    """python/ffsim/gates/orbital_rotation.py:44-154."""
    if copy:
        vec = vec.copy()
    if isinstance(nelec, (int, np.integer)):
        decomp = givens_decomposition(mat)
        vec = np.ascontiguousarray(vec.reshape((-1, 1)))
        _rotate_one_spin(vec, decomp, norb, int(nelec))
        return vec.reshape(-1)
    decomp_a, decomp_b = _givens_decompositions(mat)
    n_alpha, n_beta = nelec
    dim_a, dim_b = math.comb(norb, n_alpha), math.comb(norb, n_beta)
    vec = np.ascontiguousarray(vec.reshape((dim_a, dim_b)))
    if decomp_a is not None:
        _rotate_one_spin(vec, decomp_a, norb, n_alpha)
    if decomp_b is not None:
        vec = np.ascontiguousarray(vec.T)
        _rotate_one_spin(vec, decomp_b, norb, n_beta)
        vec = vec.T
    return vec.reshape(-1)


def _conjugate_orbital_rotation(orbital_rotation):
    """python/ffsim/gates/diag_coulomb.py:29-39."""
    if isinstance(orbital_rotation, np.ndarray) and orbital_rotation.ndim == 2:
        return orbital_rotation.T.conj()
    a, b = orbital_rotation
    return (None if a is None else a.T.conj(), None if b is None else b.T.conj())


def get_mat_exp(mat, time, norb, z_representation):
    """python/ffsim/gates/diag_coulomb.py:223-275."""
    def same_spin(m):
        if m is None:
            return np.ones((norb, norb), dtype=complex)
        m = np.array(m, dtype=float, copy=True)
        m[np.diag_indices(norb)] *= 0.5
        if z_representation:
            m *= 0.25
        return np.exp(-1j * time * m)

    if isinstance(mat, np.ndarray) and mat.ndim == 2:
        aa = same_spin(mat)
        ab = np.exp(-1j * time * (mat * 0.25 if z_representation else mat))
        return aa, ab, aa
    mat_aa, mat_ab, mat_bb = mat
    if mat_ab is None:
        ab = np.ones((norb, norb), dtype=complex)
    else:
        ab = np.exp(-1j * time * (np.asarray(mat_ab) * 0.25 if z_representation else np.asarray(mat_ab)))
    return same_spin(mat_aa), ab, same_spin(mat_bb)


def apply_diag_coulomb_evolution(
    vec, mat, time, norb, nelec, *, orbital_rotation=None, z_representation=False, copy=True
):
    """python/ffsim/gates/diag_coulomb.py:68-220."""
    if copy:
        vec = vec.copy()
    if isinstance(nelec, (int, np.integer)):
        if z_representation:
            raise NotImplementedError
        nelec = (int(nelec), 0)
    aa, ab, bb = get_mat_exp(mat, time, norb, z_representation)
    n_alpha, n_beta = nelec
    dim_a, dim_b = math.comb(norb, n_alpha), math.comb(norb, n_beta)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(
            vec, _conjugate_orbital_rotation(orbital_rotation), norb, nelec, copy=False
        )
    vec = vec.reshape((dim_a, dim_b))
    if z_representation:
        apply_diag_coulomb_evolution_in_place_z_rep(
            vec, aa, ab, bb, aa.conj(), ab.conj(), bb.conj(), norb,
            make_strings(range(norb), n_alpha), make_strings(range(norb), n_beta),
        )
    else:
        apply_diag_coulomb_evolution_in_place_num_rep(
            vec, aa, ab, bb, norb,
            gen_occslst(range(norb), n_alpha), gen_occslst(range(norb), n_beta),
        )
    vec = vec.reshape(-1)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, orbital_rotation, norb, nelec, copy=False)
    return vec


def apply_num_op_sum_evolution(vec, coeffs, time, norb, nelec, *, orbital_rotation=None, copy=True):
    """python/ffsim/gates/num_op_sum.py:62-236."""
    if copy:
        vec = vec.copy()
    if isinstance(nelec, (int, np.integer)):
        nelec = int(nelec)
        phases = np.exp(-1j * time * np.asarray(coeffs))
        if orbital_rotation is not None:
            vec = apply_orbital_rotation(vec, orbital_rotation.T.conj(), norb, nelec, copy=False)
        vec = vec.reshape((-1, 1))
        apply_num_op_sum_evolution_in_place(vec, phases, gen_occslst(range(norb), nelec))
        vec = vec.reshape(-1)
        if orbital_rotation is not None:
            vec = apply_orbital_rotation(vec, orbital_rotation, norb, nelec, copy=False)
        return vec
    if isinstance(coeffs, np.ndarray):
        phases_a = phases_b = np.exp(-1j * time * coeffs)
    else:
        ca, cb = coeffs
        phases_a = None if ca is None else np.exp(-1j * time * np.asarray(ca))
        phases_b = None if cb is None else np.exp(-1j * time * np.asarray(cb))
    n_alpha, n_beta = nelec
    dim_a, dim_b = math.comb(norb, n_alpha), math.comb(norb, n_beta)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(
            vec, _conjugate_orbital_rotation(orbital_rotation), norb, nelec, copy=False
        )
    vec = vec.reshape((dim_a, dim_b))
    if phases_a is not None:
        apply_num_op_sum_evolution_in_place(vec, phases_a, gen_occslst(range(norb), n_alpha))
    if phases_b is not None:
        vec = vec.T
        apply_num_op_sum_evolution_in_place(vec, phases_b, gen_occslst(range(norb), n_beta))
        vec = vec.T
    vec = vec.reshape(-1)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, orbital_rotation, norb, nelec, copy=False)
    return vec
