"""Duck-typed dispatch: python/ffsim/protocols/apply_unitary_protocol.py:48-88 and
linear_operator_protocol.py:40-60."""

from __future__ import annotations

from typing import Any


def apply_unitary(vec, obj: Any, norb: int, nelec, copy: bool = True):
    """Apply a unitary transformation to a vector (``obj._apply_unitary_``)."""
    method = getattr(obj, "_apply_unitary_", None)
    if method is not None:
        result = method(vec, norb=norb, nelec=nelec, copy=copy)
        if result is not NotImplemented:
            return result
    raise TypeError(
        "ffsim.apply_unitary failed. "
        "Object doesn't have a unitary effect.\n"
        f"type: {type(obj)}\n"
        f"object: {obj!r}\n"
        "The object did not have an _apply_unitary_ method that returned "
        "a value besides NotImplemented."
    )


def linear_operator(obj: Any, norb: int, nelec):
    """Return a SciPy LinearOperator representing the object (``obj._linear_operator_``)."""
    method = getattr(obj, "_linear_operator_", None)
    if method is not None:
        return method(norb=norb, nelec=nelec)
    raise TypeError(f"Object of type {type(obj)} has no _linear_operator_ method.")
