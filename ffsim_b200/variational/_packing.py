"""Real parameter vectors of the UCJ operators.

One generic packer serves the three operator classes.  The layout of the vector is the reference's
(python/ffsim/variational/ucj_spin_balanced.py:144-295, ucj_spin_unbalanced.py, ucj_spinless.py and
python/ffsim/linalg/util.py:25-185), so that parameter vectors are interchangeable with ffsim:

  per repetition:  [ unitary ] x n_rot   then one block per diagonal Coulomb matrix
  at the end:      [ unitary ] x n_rot   when there is a final orbital rotation

with a unitary stored as the parameters of its anti-Hermitian logarithm (strict upper triangle of the
real part, then upper triangle including the diagonal of the imaginary part) and a diagonal Coulomb
matrix as its entries on a list of index pairs (default: the upper triangle for symmetric blocks, the
whole matrix for the alpha-beta block of the spin-unbalanced operator).
"""

from __future__ import annotations

import itertools

import numpy as np
import scipy.linalg

PARAM_MISMATCH = ("The number of parameters passed did not match the number expected based on the function inputs. "
                  "Expected {} but got {}.")


def validate_interaction_pairs(pairs, ordered: bool) -> None:
    """python/ffsim/variational/util.py:19-35."""
    if pairs is None:
        return
    if len(set(pairs)) != len(pairs):
        raise ValueError(f"Duplicate interaction pairs encountered: {pairs}.")
    if not ordered:
        for i, j in pairs:
            if i > j:
                raise ValueError("When specifying spinless, alpha-alpha or beta-beta interaction pairs, "
                                 "you must provide only upper triangular pairs. "
                                 f"Got {(i, j)}, which is a lower triangular pair.")


def unitary_from_parameters(params: np.ndarray, dim: int) -> np.ndarray:
    """exp of the anti-Hermitian matrix with the given dim**2 real parameters."""
    params = np.asarray(params, dtype=float)
    n_triu = dim * (dim - 1) // 2
    gen = np.zeros((dim, dim), dtype=complex)
    rows, cols = np.triu_indices(dim)
    gen[rows, cols] = 1j * params[n_triu:]
    gen[cols, rows] = 1j * params[n_triu:]
    rows, cols = np.triu_indices(dim, k=1)
    gen[rows, cols] += params[:n_triu]
    gen[cols, rows] -= params[:n_triu]
    return scipy.linalg.expm(gen)


def unitary_to_parameters(mat: np.ndarray) -> np.ndarray:
    dim = mat.shape[0]
    gen = scipy.linalg.logm(mat)
    n_triu = dim * (dim - 1) // 2
    out = np.zeros(dim * dim)
    rows, cols = np.triu_indices(dim, k=1)
    out[:n_triu] = gen[rows, cols].real
    rows, cols = np.triu_indices(dim)
    out[n_triu:] = gen[rows, cols].imag
    return out


class Block:
    """One diagonal Coulomb matrix of a repetition: its default index pairs and whether it is symmetric."""

    def __init__(self, symmetric: bool):
        self.symmetric = symmetric

    def default_pairs(self, norb: int):
        if self.symmetric:
            return list(itertools.combinations_with_replacement(range(norb), 2))
        return list(itertools.product(range(norb), repeat=2))


SYM, FULL = Block(True), Block(False)


def resolve_pairs(norb: int, blocks, pairs):
    """The index pairs of every block (validated; ``None`` means the block's default)."""
    if pairs is None:
        pairs = (None,) * len(blocks)
    out = []
    for block, p in zip(blocks, pairs):
        validate_interaction_pairs(p, ordered=not block.symmetric)
        out.append(block.default_pairs(norb) if p is None else list(p))
    return out


def count(norb: int, n_reps: int, blocks, pairs, n_rot: int, with_final: bool) -> int:
    per_rep = sum(len(p) for p in resolve_pairs(norb, blocks, pairs)) + n_rot * norb**2
    return n_reps * per_rep + (n_rot * norb**2 if with_final else 0)


def unpack(params, norb: int, n_reps: int, blocks, pairs, n_rot: int, with_final: bool):
    """-> (diag_coulomb_mats [n_reps, n_blocks, norb, norb], rotations [n_reps, n_rot, norb, norb], final or None)."""
    n_expected = count(norb, n_reps, blocks, pairs, n_rot, with_final)
    if len(params) != n_expected:
        raise ValueError(PARAM_MISMATCH.format(n_expected, len(params)))
    params = np.asarray(params, dtype=float)
    resolved = resolve_pairs(norb, blocks, pairs)
    mats = np.zeros((n_reps, len(blocks), norb, norb))
    rots = np.zeros((n_reps, n_rot, norb, norb), dtype=complex)
    pos = 0

    def take(n):
        nonlocal pos
        chunk = params[pos : pos + n]
        pos += n
        return chunk

    for rep in range(n_reps):
        for k in range(n_rot):
            rots[rep, k] = unitary_from_parameters(take(norb**2), norb)
        for b, (block, idx) in enumerate(zip(blocks, resolved)):
            if not idx:
                continue
            r, c = (list(t) for t in zip(*idx))
            vals = take(len(idx))
            mats[rep, b, r, c] = vals
            if block.symmetric:
                mats[rep, b, c, r] = vals
    final = None
    if with_final:
        final = np.stack([unitary_from_parameters(take(norb**2), norb) for _ in range(n_rot)])
    return mats, rots, final


def pack(mats, rots, final, blocks, pairs) -> np.ndarray:
    n_reps, n_rot, norb, _ = rots.shape
    resolved = resolve_pairs(norb, blocks, pairs)
    out = []
    for rep in range(n_reps):
        for k in range(n_rot):
            out.append(unitary_to_parameters(rots[rep, k]))
        for b, idx in enumerate(resolved):
            if idx:
                r, c = (list(t) for t in zip(*idx))
                out.append(np.asarray(mats[rep, b, r, c], dtype=float))
    if final is not None:
        out.extend(unitary_to_parameters(u) for u in final)
    return np.concatenate(out) if out else np.zeros(0)
