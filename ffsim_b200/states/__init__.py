"""State-vector helpers needed to build inputs for the hot path."""

from __future__ import annotations

import math
import numbers
from typing import Any

import numpy as np


def dims(norb: int, nelec: tuple[int, int]) -> tuple[int, int]:
    """python/ffsim/states/dimensions.py:18-32."""
    n_alpha, n_beta = nelec
    return math.comb(norb, n_alpha), math.comb(norb, n_beta)


def dim(norb: int, nelec: int | tuple[int, int]) -> int:
    """python/ffsim/states/dimensions.py:35-50."""
    if isinstance(nelec, numbers.Integral):
        return math.comb(norb, int(nelec))
    n_alpha, n_beta = nelec
    return math.comb(norb, n_alpha) * math.comb(norb, n_beta)


def hartree_fock_state(norb: int, nelec: int | tuple[int, int], *, device: Any = None):
    """python/ffsim/states/slater.py:122-138: one-hot at address 0.

    With ``device`` set (e.g. ``"cuda"``) the state is created as a torch tensor on
    that device instead of as a NumPy array.
    """
    d = dim(norb, nelec)
    if device is None:
        vec = np.zeros(d, dtype=complex)
        vec[0] = 1
        return vec
    import torch

    vec = torch.zeros(d, dtype=torch.complex128, device=device)
    vec[0] = 1
    return vec


from ffsim_b200.states.spin import Spin, pair_for_spin  # noqa: E402

__all__ = ["Spin", "dim", "dims", "hartree_fock_state", "pair_for_spin"]
