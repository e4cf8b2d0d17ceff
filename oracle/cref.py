"""Fast CPU oracle: the reference's Python drivers over the C restatement of its Rust kernels.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Same call structure as the
reference -- one native call per Givens rotation and per phase shift, physical
transposes around the beta sector (python/ffsim/gates/orbital_rotation.py:117-154)
-- so that timing it is a fair stand-in for ``ffsim`` on the host cores
(``bench.py`` cpu_baseline, kind "port") and so that parity at the BASELINE
shapes (C1, C2) can be checked in seconds.  The kernels live in
``oracle/c/ref_kernels.c``; they are validated against the numpy oracle in
``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

from oracle.cistring import gen_occslst, make_strings, one_subspace_indices, zero_one_subspace_indices
from oracle.gates import _conjugate_orbital_rotation, _givens_decompositions, get_mat_exp
from oracle.givens import givens_decomposition

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "libffsim_ref_kernels.so")
_lib = None

_VP, _I64, _DBL, _INT = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_int


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _DIR], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.ref_max_threads.restype = _INT
        _lib.ref_apply_givens_rotation_in_place.argtypes = [_VP, _I64, _DBL, _DBL, _DBL, _VP, _VP, _I64, _INT]
        _lib.ref_apply_phase_shift_in_place.argtypes = [_VP, _I64, _DBL, _DBL, _VP, _I64]
        _lib.ref_apply_num_op_sum_evolution_in_place.argtypes = [_VP, _I64, _I64, _I64, _I64, _VP, _VP, _I64]
        _lib.ref_apply_diag_coulomb_evolution_in_place_num_rep.argtypes = [
            _VP, _I64, _I64, _VP, _VP, _VP, _I64, _VP, _I64, _VP, _I64]
        _lib.ref_apply_diag_coulomb_evolution_in_place_z_rep.argtypes = [
            _VP, _I64, _I64, _VP, _VP, _VP, _I64, _VP, _VP]
        _lib.ref_contract_num_op_sum_spin_into_buffer.argtypes = [_VP, _I64, _I64, _I64, _I64, _VP, _VP, _I64, _VP]
        _lib.ref_contract_diag_coulomb_into_buffer_num_rep.argtypes = [
            _VP, _I64, _I64, _VP, _VP, _VP, _I64, _VP, _I64, _VP, _I64, _VP]
        _lib.ref_contract_diag_coulomb_into_buffer_z_rep.argtypes = [
            _VP, _I64, _I64, _VP, _VP, _VP, _I64, _VP, _VP, _VP]
        _lib.ref_transpose.argtypes = [_VP, _VP, _I64, _I64]
    return _lib


def n_threads() -> int:
    """Thread count the kernels use: RAYON_NUM_THREADS if set (as the reference reads it,
    src/gates/orbital_rotation.rs:37-45), else all cores OpenMP sees."""
    env = os.environ.get("RAYON_NUM_THREADS")
    return int(env) if env else int(lib().ref_max_threads())


def _p(a):
    return a.ctypes.data_as(_VP)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _rotate_one_spin(vec2d, decomp, norb, nocc) -> None:
    rotations, phase_shifts = decomp
    L, nt = lib(), n_threads()
    dim_b = vec2d.shape[1]
    for c, s, i, j in rotations:
        idx = zero_one_subspace_indices(norb, nocc, (i, j))
        half = len(idx) // 2
        s1, s2 = _c(idx[:half], np.uint64), _c(idx[half:], np.uint64)
        sc = np.conj(s)
        L.ref_apply_givens_rotation_in_place(_p(vec2d), dim_b, float(c), sc.real, sc.imag, _p(s1), _p(s2), half, nt)
    for i, phase in enumerate(phase_shifts):
        idx = _c(one_subspace_indices(norb, nocc, (i,)), np.uint64)
        L.ref_apply_phase_shift_in_place(_p(vec2d), dim_b, phase.real, phase.imag, _p(idx), len(idx))


def _transpose(mat2d):
    out = np.empty((mat2d.shape[1], mat2d.shape[0]), dtype=complex)
    lib().ref_transpose(_p(mat2d), _p(out), mat2d.shape[0], mat2d.shape[1])
    return out


def apply_orbital_rotation(vec, mat, norb, nelec, *, copy=True):
    """python/ffsim/gates/orbital_rotation.py:44-154 over the C kernels."""
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
        vec = _transpose(vec)
        _rotate_one_spin(vec, decomp_b, norb, n_beta)
        vec = _transpose(vec)
    return vec.reshape(-1)


def apply_diag_coulomb_evolution(
    vec, mat, time, norb, nelec, *, orbital_rotation=None, z_representation=False, copy=True
):
    """python/ffsim/gates/diag_coulomb.py:68-220 over the C kernels."""
    if copy:
        vec = vec.copy()
    aa, ab, bb = (_c(m, complex) for m in get_mat_exp(mat, time, norb, z_representation))
    n_alpha, n_beta = nelec
    dim_a, dim_b = math.comb(norb, n_alpha), math.comb(norb, n_beta)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, _conjugate_orbital_rotation(orbital_rotation), norb, nelec, copy=False)
    vec = np.ascontiguousarray(vec.reshape((dim_a, dim_b)))
    L = lib()
    if z_representation:
        sa, sb = _c(make_strings(range(norb), n_alpha), np.int64), _c(make_strings(range(norb), n_beta), np.int64)
        L.ref_apply_diag_coulomb_evolution_in_place_z_rep(_p(vec), dim_a, dim_b, _p(aa), _p(ab), _p(bb), norb, _p(sa), _p(sb))
    else:
        oa, ob = _c(gen_occslst(range(norb), n_alpha), np.uint64), _c(gen_occslst(range(norb), n_beta), np.uint64)
        L.ref_apply_diag_coulomb_evolution_in_place_num_rep(
            _p(vec), dim_a, dim_b, _p(aa), _p(ab), _p(bb), norb, _p(oa), n_alpha, _p(ob), n_beta)
    vec = vec.reshape(-1)
    if orbital_rotation is not None:
        vec = apply_orbital_rotation(vec, orbital_rotation, norb, nelec, copy=False)
    return vec


def ucj_spin_balanced_apply(vec, diag_coulomb_mats, orbital_rotations, final_orbital_rotation, norb, nelec, copy=True):
    """python/ffsim/variational/ucj_spin_balanced.py:657-696 over the C kernels."""
    if copy:
        vec = vec.copy()
    current_basis = np.eye(norb)
    for (mat_aa, mat_ab), orbital_rotation in zip(diag_coulomb_mats, orbital_rotations):
        vec = apply_orbital_rotation(vec, orbital_rotation.T.conj() @ current_basis, norb, nelec, copy=False)
        vec = apply_diag_coulomb_evolution(vec, (mat_aa, mat_ab, mat_aa), -1.0, norb, nelec, copy=False)
        current_basis = orbital_rotation
    if final_orbital_rotation is not None:
        current_basis = final_orbital_rotation @ current_basis
    return apply_orbital_rotation(vec, current_basis, norb, nelec, copy=False)
