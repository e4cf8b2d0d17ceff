"""Matrix predicates used for validation (python/ffsim/linalg/predicates.py:26-83)."""

from __future__ import annotations

import numpy as np


def is_hermitian(mat: np.ndarray, *, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    m, n = mat.shape
    return bool(m == n and np.allclose(mat, mat.T.conj(), rtol=rtol, atol=atol))


def is_real_symmetric(mat: np.ndarray, *, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    m, n = mat.shape
    return bool(m == n and np.all(np.isreal(mat)) and np.allclose(mat, mat.T, rtol=rtol, atol=atol))


def is_unitary(mat: np.ndarray, *, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    m, n = mat.shape
    return bool(m == n and np.allclose(mat @ mat.T.conj(), np.eye(m), rtol=rtol, atol=atol))
