"""Suzuki-Trotter term iterator: python/ffsim/trotter/_util.py:18-55."""

from __future__ import annotations

from collections.abc import Iterator


def simulate_trotter_step_iterator(n_terms: int, time: float, order: int = 0) -> Iterator[tuple[int, float]]:
    if order == 0:
        for i in range(n_terms):
            yield i, time
    elif order == 1:
        for i in range(n_terms - 1):
            yield i, time / 2
        yield n_terms - 1, time
        for i in reversed(range(n_terms - 1)):
            yield i, time / 2
    else:
        split_time = time / (4 - 4 ** (1 / (2 * order - 1)))
        for _ in range(2):
            yield from simulate_trotter_step_iterator(n_terms, split_time, order - 1)
        yield from simulate_trotter_step_iterator(n_terms, time - 4 * split_time, order - 1)
        for _ in range(2):
            yield from simulate_trotter_step_iterator(n_terms, split_time, order - 1)
