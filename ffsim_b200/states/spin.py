"""Spin selector of the gate functions (python/ffsim/states/spin.py:19-50)."""

from __future__ import annotations

from enum import Flag, auto


class Spin(Flag):
    """Which spin sector(s) a gate acts on."""

    ALPHA = auto()
    BETA = auto()
    ALPHA_AND_BETA = ALPHA | BETA


def pair_for_spin(obj, spin: Spin):
    """``(obj or None, obj or None)`` for (alpha, beta) according to ``spin``."""
    return (obj if spin & Spin.ALPHA else None, obj if spin & Spin.BETA else None)
