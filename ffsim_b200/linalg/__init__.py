"""Linear-algebra helpers on the hot path (host side, norb x norb)."""

from ffsim_b200.linalg.givens import GivensRotation, givens_decomposition
from ffsim_b200.linalg.predicates import is_hermitian, is_real_symmetric, is_unitary

__all__ = ["GivensRotation", "givens_decomposition", "is_hermitian", "is_real_symmetric", "is_unitary"]
