"""Cache initialisation: python/ffsim/init_cache.py:25-41."""

from __future__ import annotations

from ffsim_b200.cistring import gen_occslst, get_tables, make_strings


def init_cache(norb: int, nelec: tuple[int, int]) -> None:
    """Build the string tables for ``(norb, nelec)`` ahead of time.

    Call before benchmarking so that table construction is not counted.  The
    per-unitary plan structure (which depends on the rotation pattern) is cached
    on first use by ``apply_orbital_rotation``.
    """
    for nocc in nelec:
        get_tables(norb, nocc)
        make_strings(norb, nocc)
        gen_occslst(norb, nocc)
