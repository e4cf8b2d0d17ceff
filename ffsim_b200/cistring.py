"""Occupation-string tables (host copies for inspection, handles for the kernels).

Counterpart of python/ffsim/_cistring.py:21-42 and of the cached address tables
in python/ffsim/gates/orbital_rotation.py:203-236; the arithmetic is in
csrc/tables.cpp.
"""

from __future__ import annotations

import ctypes
from functools import lru_cache

import numpy as np

from ffsim_b200 import _lib


class Tables:
    """Owns an ``ffb_tables`` handle for one (norb, nocc) sector."""

    def __init__(self, norb: int, nocc: int):
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib.ffb_tables_create(int(norb), int(nocc), ctypes.byref(handle)))
        self.handle = handle
        self.norb = int(norb)
        self.nocc = int(nocc)
        self.dim = int(_lib.lib.ffb_tables_dim(handle))


@lru_cache(maxsize=None)
def _tables_for_device(norb: int, nocc: int, device: int) -> Tables:
    return Tables(norb, nocc)


def get_tables(norb: int, nocc: int) -> Tables:
    """The (norb, nocc) sector's handle for the CURRENT CUDA device.

    A handle owns device buffers (the uploaded strings, per-call scratch), so one handle per
    device: the same sector used on cuda:0 and then on cuda:1 must not share them.
    """
    import torch

    device = torch.cuda.current_device() if torch.cuda.is_available() else -1
    return _tables_for_device(int(norb), int(nocc), device)


@lru_cache(maxsize=None)
def make_strings(norb: int, nocc: int) -> np.ndarray:
    """``cistring.make_strings(range(norb), nocc)``: int64, ascending."""
    t = get_tables(norb, nocc)
    out = np.zeros(t.dim, dtype=np.int64)
    _lib.check(_lib.lib.ffb_tables_strings(t.handle, _lib.ptr(out)))
    out.setflags(write=False)
    return out


@lru_cache(maxsize=None)
def gen_occslst(norb: int, nocc: int) -> np.ndarray:
    """``cistring.gen_occslst(range(norb), nocc)`` cast to ``np.uint`` (``_cistring.py:27-31``)."""
    t = get_tables(norb, nocc)
    out = np.zeros((t.dim, nocc), dtype=np.uint64)
    _lib.check(_lib.lib.ffb_tables_occupations(t.handle, _lib.ptr(out)))
    out.setflags(write=False)
    return out


def strs2addr(norb: int, nocc: int, strings) -> np.ndarray:
    t = get_tables(norb, nocc)
    s = np.ascontiguousarray(np.atleast_1d(strings), dtype=np.int64)
    out = np.zeros(len(s), dtype=np.int64)
    _lib.check(_lib.lib.ffb_tables_strs2addr(t.handle, _lib.ptr(s), len(s), _lib.ptr(out)))
    return out


@lru_cache(maxsize=None)
def zero_one_subspace_indices(norb: int, nocc: int, target_orbs: tuple[int, int]) -> np.ndarray:
    """``_zero_one_subspace_indices`` (``gates/orbital_rotation.py:203-213``)."""
    t = get_tables(norb, nocc)
    n_pairs = int(_lib.lib.ffb_tables_n_pairs(t.handle))
    out = np.zeros(2 * n_pairs, dtype=np.uint64)
    n = ctypes.c_int64()
    i, j = target_orbs
    _lib.check(
        _lib.lib.ffb_tables_zero_one_subspace(t.handle, int(i), int(j), _lib.ptr(out), ctypes.byref(n))
    )
    out.setflags(write=False)
    return out


@lru_cache(maxsize=None)
def one_subspace_indices(norb: int, nocc: int, target_orbs: tuple[int]) -> np.ndarray:
    """``_one_subspace_indices`` for a single target orbital (``orbital_rotation.py:216-226``)."""
    (i,) = target_orbs
    t = get_tables(norb, nocc)
    n_one = int(_lib.lib.ffb_tables_n_one(t.handle))
    out = np.zeros(n_one, dtype=np.uint64)
    n = ctypes.c_int64()
    _lib.check(_lib.lib.ffb_tables_one_subspace(t.handle, int(i), _lib.ptr(out), ctypes.byref(n)))
    out.setflags(write=False)
    return out
