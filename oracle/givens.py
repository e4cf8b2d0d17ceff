"""Givens decomposition of a unitary (oracle; test infrastructure only).

Follows src/linalg/givens.rs:20-149 (``zrotg_safe`` :20-34, column/row rotation
helpers :36-68, elimination sweeps :86-124, left->right conversion :126-146).
Returns ``([(c, s, i, j), ...], phases)`` with ``U = D * G_L^* ... G_1^*`` in the
convention of python/ffsim/linalg/givens.py:59-156.
"""

from __future__ import annotations

import math

import numpy as np


def zrotg_safe(a: complex, b: complex, tol: float) -> tuple[float, complex]:
    """src/linalg/givens.rs:20-34."""
    abs_a, abs_b = abs(a), abs(b)
    if abs_b <= tol:
        return 1.0, 0j
    if abs_a <= tol:
        return 0.0, 1 + 0j
    r = math.hypot(abs_a, abs_b)
    c = abs_a / r
    s = (a / abs_a) * b.conjugate() / r
    return min(max(c, -1.0), 1.0), s


def givens_decomposition(mat, tol: float = 1e-12):
    mat = np.asarray(mat)
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1]:
        raise ValueError("mat must be a square matrix")
    n = mat.shape[0]
    cur = mat.astype(complex, copy=True)
    left: list[tuple[float, complex, int, int]] = []
    right: list[tuple[float, complex, int, int]] = []

    def rot(x, y, c, s):
        return c * x + s * y, c * y - np.conj(s) * x

    for i in range(max(n - 1, 0)):
        if i % 2 == 0:
            # zero out an anti-diagonal from the right: column operations
            for j in range(i + 1):
                t = i - j
                row = n - j - 1
                if abs(cur[row, t]) > tol:
                    c, s = zrotg_safe(complex(cur[row, t + 1]), complex(cur[row, t]), tol)
                    right.append((c, s, t + 1, t))
                    cur[:, t + 1], cur[:, t] = rot(cur[:, t + 1].copy(), cur[:, t].copy(), c, s)
        else:
            # zero out an anti-diagonal from the left: row operations
            for j in range(i + 1):
                t = n - i + j - 1
                col = j
                if abs(cur[t, col]) > tol:
                    c, s = zrotg_safe(complex(cur[t - 1, col]), complex(cur[t, col]), tol)
                    left.append((c, s, t - 1, t))
                    cur[t - 1, :], cur[t, :] = rot(cur[t - 1, :].copy(), cur[t, :].copy(), c, s)

    # commute the left rotations through the diagonal (givens.rs:126-146)
    for c_l, s_l, i, j in reversed(left):
        c, s = zrotg_safe(c_l * complex(cur[j, j]), s_l.conjugate() * complex(cur[i, i]), tol)
        right.append((c, -s.conjugate(), i, j))
        di, dj = complex(cur[i, i]), complex(cur[j, j])
        g00, g01 = c * di, -s * dj
        g10, g11 = s.conjugate() * di, c * dj
        c2, s2 = zrotg_safe(g11, g10, tol)
        cur[i, i] = g00 * c2 + g01 * (-s2.conjugate())
        cur[j, j] = g10 * s2 + g11 * c2

    return right, np.diagonal(cur).copy()
