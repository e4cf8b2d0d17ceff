"""Givens decomposition of a unitary (exact path of python/ffsim/linalg/givens.py:59-156)."""

from __future__ import annotations

import ctypes
from typing import NamedTuple

import numpy as np

from ffsim_b200 import _lib


class GivensRotation(NamedTuple):
    """python/ffsim/linalg/givens.py:26-56: the rotation [[c, s], [-conj(s), c]] on orbitals (i, j)."""

    c: float
    s: complex
    i: int
    j: int


def _decompose_raw(mat: np.ndarray, tol: float = 1e-12) -> tuple[np.ndarray, np.ndarray]:
    """Structured array of rotations (``_lib.GIVENS_DTYPE``) and the diagonal phases."""
    mat = np.asarray(mat)
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1]:
        raise ValueError("mat must be a square matrix")  # src/linalg/givens.rs:78-80
    n = mat.shape[0]
    m = np.ascontiguousarray(mat.astype(complex))
    rots = np.zeros(max(1, n * (n - 1) // 2), dtype=_lib.GIVENS_DTYPE)
    phases = np.zeros(n, dtype=complex)
    n_rot = ctypes.c_int(0)
    _lib.check(
        _lib.lib.ffb_givens_decomposition(
            _lib.ptr(m), n, float(tol), _lib.ptr(rots), ctypes.byref(n_rot), _lib.ptr(phases)
        )
    )
    return rots[: n_rot.value], phases


def givens_decomposition(mat: np.ndarray) -> tuple[list[GivensRotation], np.ndarray]:
    r"""Givens rotation decomposition of a unitary matrix.

    Same contract as ``ffsim.linalg.givens_decomposition``: returns the rotations
    and the diagonal phases with :math:`U = D G_L^* \cdots G_1^*`; every rotation
    acts on adjacent rows/columns; entries below 1e-12 are skipped.
    """
    rots, phases = _decompose_raw(mat)
    return (
        [GivensRotation(float(r["c"]), complex(r["s"]), int(r["i"]), int(r["j"])) for r in rots],
        phases,
    )
